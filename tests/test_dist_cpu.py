"""world_size-2 gloo tests of the multi-GPU host logic (sharding, gather layout, merge order).
The per-shard kernels are replaced by the oracle here; the GPU versions are in test_retrieval_gpu."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import search


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audio_sheet_retrieval_b200.dist import gather_topk, shard_bounds
    rng = np.random.RandomState(0)
    D = rng.normal(size=(1001, 32)).astype(np.float32)
    D[900] = D[3]                                           # tie across shards -> lower global index wins
    Q = rng.normal(size=(7, 32)).astype(np.float32)
    k = 5
    lo, hi = shard_bounds(len(D), rank, world)
    s, i = search.pinned_topk(Q, D[lo:hi], k, idx_base=lo)
    gs, gi = gather_topk(torch.as_tensor(s), torch.as_tensor(i))
    assert gs.shape == (7, world * k)
    ms, mi = search.merge_topk([gs.numpy()], [gi.numpy()], k)
    s_ref, i_ref = search.pinned_topk(Q, D, k)
    ok = bool((mi == i_ref).all() and (ms == s_ref).all())
    # CCA sums all-reduce: sum of per-shard second moments == global
    H = rng.normal(size=(501, 4))
    lo2, hi2 = shard_bounds(len(H), rank, world)
    part = torch.as_tensor(H[lo2:hi2].T @ H[lo2:hi2])
    dist.all_reduce(part)
    ok = ok and bool(np.allclose(part.numpy(), H.T @ H))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_sharded_topk_gather_merge_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_bounds_cover_everything():
    from audio_sheet_retrieval_b200.dist import shard_bounds
    for n in (1, 7, 1000, 10 ** 8 + 3):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
