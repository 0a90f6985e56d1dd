"""Worker of tests/test_multigpu_gpu.py (launched by torchrun, one rank per GPU, NCCL).  Every sharded code path of
SURVEY.md 8(e) against its single-GPU result on the SAME device kernels: top-k (bit-exact), rank-of-target /
eval_retrieval (bit-exact), CCA.fit (1e-9) and the refine_cca script (pickle entries, 1e-6 after the float32 cast)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_dir):
    from audio_sheet_retrieval_b200.dist import ShardedDB, init_from_env, shard_bounds
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    from audio_sheet_retrieval_b200.utils.cca import CCA
    from audio_sheet_retrieval_b200.utils.train_dcca_pool import eval_retrieval, retrieval_ranks
    rank, world, local = init_from_env("nccl")
    dev = torch.device("cuda", local)
    group = dist.group.WORLD
    res = {"rank": rank, "world": world}
    rng = np.random.RandomState(5)

    # 1. sharded top-k == single-GPU top-k, indices and scores, incl. a cross-shard tie and the tf32 pre-filter path
    for name, n_db, nq, k in (("exact", 30011, 37, 25), ("prefilter", 400000, 512, 25)):
        D = rng.normal(size=(n_db, 32)).astype(np.float32)
        D[n_db - 7] = D[11]
        Q = (D[rng.randint(0, n_db, nq)] + 0.3 * rng.normal(size=(nq, 32))).astype(np.float32)
        Dg, Qg = torch.as_tensor(D).to(dev), torch.as_tensor(Q).to(dev)
        s_ref, i_ref = EmbeddingDB(Dg).topk_device(Qg, k)
        lo, hi = shard_bounds(n_db, rank, world)
        sdb = ShardedDB(Dg[lo:hi].contiguous(), lo, group=group)
        s, i = sdb.topk_device(Qg, k)
        res["topk_" + name] = bool(torch.equal(i, i_ref) and torch.equal(s, s_ref))
        # donated buffer (rows normalised in place): same answer, no second copy
        sdb2 = ShardedDB(Dg[lo:hi].clone(), lo, group=group, normalise_in_place=True)
        s2, i2 = sdb2.topk_device(Qg, k)
        res["topk_inplace_" + name] = bool(torch.equal(i2, i_ref) and torch.equal(s2, s_ref))

    # 1b. piece identification with the recordings split over the ranks (whole DB on every rank) == single GPU == sharded DB
    from audio_sheet_retrieval_b200.dist import ReplicatedDB
    from audio_sheet_retrieval_b200.retrieval import vote_device
    n_db, n_rec, win = 60000, 7, 30
    D = rng.normal(size=(n_db, 32)).astype(np.float32)
    ids = (np.arange(n_db) // 50).astype(np.int32)
    rows_q = (rng.randint(0, n_db // 50, n_rec)[:, None] * 50 + rng.randint(0, 50, (n_rec, win))).reshape(-1)
    Q = (D[rows_q] + 0.3 * rng.normal(size=(n_rec * win, 32))).astype(np.float32)
    Dg, Qg = torch.as_tensor(D).to(dev), torch.as_tensor(Q).to(dev)
    _, i_ref = EmbeddingDB(Dg).topk_device(Qg, 10)
    p_ref, c_ref = vote_device(i_ref.view(n_rec, -1), torch.as_tensor(ids).to(dev), 4)
    p_rep, c_rep = ReplicatedDB(Dg, ids, group=group).identify(Qg, n_rec, 4, 10)
    lo, hi = shard_bounds(n_db, rank, world)
    p_sh, c_sh = ShardedDB(Dg[lo:hi].contiguous(), lo, row_ids_global=ids, group=group).identify(Qg, n_rec, 4, 10)
    res["identify_replicated"] = bool(torch.equal(p_rep, p_ref) and torch.equal(c_rep, c_ref))
    res["identify_sharded"] = bool(torch.equal(p_sh, p_ref) and torch.equal(c_sh, c_ref))

    # 2. eval_retrieval with the view-2 rows sharded == unsharded (ranks and target scores bit-exact), grouped views too
    for name, n1, n2 in (("square", 1501, 1501), ("grouped", 500, 1500)):
        base = rng.normal(size=(max(n1, n2), 32))
        a = (base[:n1] + 0.8 * rng.normal(size=(n1, 32))).astype(np.float32)
        b = ((np.repeat(base[:n1], n2 // n1, axis=0) if n2 > n1 else base[:n2]) + 0.8 * rng.normal(size=(n2, 32))).astype(np.float32)
        r_ref, t_ref = retrieval_ranks(a, b)
        r, t = retrieval_ranks(a, b, group=group)
        res["ranks_" + name] = bool((r == r_ref).all() and (t == t_ref).all())
        res["eval_" + name] = bool(eval_retrieval(a, b, group=group)[4] == eval_retrieval(a, b)[4])

    # 3. CCA.fit with rows sharded + one all-reduce == single-GPU fit
    Z = rng.normal(size=(25000, 32))
    H1 = (Z @ rng.normal(size=(32, 32)) * 0.05 + 0.3).astype(np.float32)
    H2 = (Z @ rng.normal(size=(32, 32)) * 0.05 + 0.02 * rng.normal(size=(25000, 32)) - 0.1).astype(np.float32)
    ref = CCA()
    sig_ref = ref.fit(H1, H2)
    lo, hi = shard_bounds(len(H1), rank, world)
    c = CCA()
    sig = c.fit(H1[lo:hi], H2[lo:hi], group=group)
    res["cca_sigma"] = float(np.abs(sig - sig_ref).max())
    res["cca_UV"] = float(max(np.abs(c.U - ref.U).max() / np.abs(ref.U).max(), np.abs(c.V - ref.V).max() / np.abs(ref.V).max()))
    res["cca_means"] = float(max(np.abs(c.m1 - ref.m1).max(), np.abs(c.m2 - ref.m2).max()))

    # 4. the refine_cca script under torchrun == its single-process run (rank 0 writes the pickle)
    from audio_sheet_retrieval_b200 import refine_cca
    from audio_sheet_retrieval_b200.params import load_params
    pkl = os.path.join(ROOT, "tests", "golden", "params_all_split_mutopia_full_aug.pkl")
    out = os.path.join(out_dir, "sharded", "params.pkl")
    refine_cca.main(["--model", "mutopia_ccal_cont_rsz", "--data", "synthetic", "--n_train", "600", "--param_file", pkl,
                     "--out_file", out])
    dist.barrier()
    if rank == 0:
        os.environ["WORLD_SIZE"], os.environ["RANK"] = "1", "0"     # single-process refit of the same 600 rows
        out1 = os.path.join(out_dir, "single", "params.pkl")
        refine_cca.main(["--model", "mutopia_ccal_cont_rsz", "--data", "synthetic", "--n_train", "600", "--param_file", pkl,
                         "--out_file", out1])
        a, b = load_params(out), load_params(out1)
        res["refine_unchanged"] = bool(all((a[i] == b[i]).all() for i in range(97) if i not in (90, 91, 92, 93)))
        res["refine_diff"] = float(max(np.abs(a[i] - b[i]).max() / max(1.0, np.abs(b[i]).max()) for i in (90, 91, 92, 93)))
        os.environ["WORLD_SIZE"], os.environ["RANK"] = str(world), "0"
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as fp:
        json.dump(res, fp)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
