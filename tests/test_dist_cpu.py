"""world_size-2 gloo tests of the multi-GPU host logic (sharding, the single-buffer chunk layout of the
all-gathers, merge order, the counted CCA all-reduce).  The per-shard kernels are replaced by the oracle here;
the same code paths run on NCCL with the real kernels in tests/test_multigpu_gpu.py."""
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
    from audio_sheet_retrieval_b200.retrieval import chunk_layout, chunk_views
    rng = np.random.RandomState(0)
    D = rng.normal(size=(1001, 32)).astype(np.float32)
    D[900] = D[3]                                           # tie across shards -> lower global index wins
    Q = rng.normal(size=(7, 32)).astype(np.float32)
    k = 5
    lo, hi = shard_bounds(len(D), rank, world)
    s, i = search.pinned_topk(Q, D[lo:hi], k, idx_base=lo)
    gathered = gather_topk(torch.as_tensor(s), torch.as_tensor(i))
    chunk, idx_off = chunk_layout(7, k)
    ok = gathered.dtype == torch.uint8 and gathered.numel() == world * chunk and chunk % 8 == 0 and idx_off % 8 == 0
    lists = [chunk_views(gathered, 7, k, r) for r in range(world)]
    ok = ok and bool((lists[rank][0].numpy() == s).all() and (lists[rank][1].numpy() == i).all())
    gs = np.concatenate([l[0].numpy() for l in lists], axis=1)
    gi = np.concatenate([l[1].numpy() for l in lists], axis=1)
    ms, mi = search.merge_topk([gs], [gi], k)
    s_ref, i_ref = search.pinned_topk(Q, D, k)
    ok = ok and bool((mi == i_ref).all() and (ms == s_ref).all())
    # rank-of-target exchange: chunks [tscore (nq) | pad | tidx (nq)] with k = 1; odd nq exercises the padding
    nq = 7
    ts = np.full(nq, -np.inf, np.float32)
    ti = np.full(nq, -1, np.int64)
    for j in range(nq):
        if (j % world) == rank:                             # this shard owns query j's correct item
            ts[j], ti[j] = 0.25 * j, 100 * rank + j
    mine = torch.empty(chunk_layout(nq, 1)[0], dtype=torch.uint8)
    cs, ci = chunk_views(mine, nq, 1)
    cs.view(-1).copy_(torch.as_tensor(ts)); ci.view(-1).copy_(torch.as_tensor(ti))
    g2 = torch.empty(world * mine.numel(), dtype=torch.uint8)
    dist.all_gather_into_tensor(g2, mine)
    best = [max(((chunk_views(g2, nq, 1, r)[0][j, 0].item(), -chunk_views(g2, nq, 1, r)[1][j, 0].item())
                 for r in range(world) if chunk_views(g2, nq, 1, r)[1][j, 0].item() >= 0)) for j in range(nq)]
    ok = ok and all(best[j] == (0.25 * j, -(100 * (j % world) + j)) for j in range(nq))
    # counted CCA sums: [second moments | row count] all-reduced as one buffer == global
    H = rng.normal(size=(501, 4))
    lo2, hi2 = shard_bounds(len(H), rank, world)
    part = torch.cat([torch.as_tensor(H[lo2:hi2].T @ H[lo2:hi2]).view(-1), torch.tensor([float(hi2 - lo2)], dtype=torch.float64)])
    dist.all_reduce(part)
    ok = ok and bool(np.allclose(part[:-1].numpy().reshape(4, 4), H.T @ H)) and part[-1].item() == 501.0
    # recordings split over the ranks (ReplicatedDB): uneven blocks, one all-gather, padding rows dropped
    from audio_sheet_retrieval_b200.dist import gather_recording_results, recording_layout
    n_rec, top_k = 7, 3
    per, bounds, rowmap = recording_layout(n_rec, world)
    expect_ids = np.arange(n_rec * top_k, dtype=np.int32).reshape(n_rec, top_k)
    lo3, hi3 = bounds[rank]
    mine = torch.full((2, per, top_k), -1, dtype=torch.int32)
    mine[0, :hi3 - lo3] = torch.as_tensor(expect_ids[lo3:hi3])
    mine[1, :hi3 - lo3] = torch.as_tensor(expect_ids[lo3:hi3] + 1000)
    gathered = torch.empty((world, 2, per, top_k), dtype=torch.int32)
    gi3, gc3 = gather_recording_results(mine, gathered, torch.as_tensor(rowmap, dtype=torch.int64))
    ok = ok and per == 4 and sum(h - l for l, h in bounds) == n_rec
    ok = ok and bool((gi3.numpy() == expect_ids).all() and (gc3.numpy() == expect_ids + 1000).all())
    q.put((rank, bool(ok)))
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


def test_chunk_layout_alignment():
    from audio_sheet_retrieval_b200.retrieval import chunk_layout, chunk_views
    for nq, k in ((1, 1), (7, 5), (10000, 25), (3, 128)):
        chunk, off = chunk_layout(nq, k)
        assert chunk % 8 == 0 and off % 8 == 0 and off >= nq * k * 4 and chunk == off + nq * k * 8
        buf = torch.zeros(3 * chunk, dtype=torch.uint8)
        s, i = chunk_views(buf, nq, k, rank=2)
        assert s.shape == (nq, k) and i.shape == (nq, k) and s.dtype == torch.float32 and i.dtype == torch.int64
        i.fill_(-1)
        assert buf[:2 * chunk].sum() == 0                   # views alias exactly the third chunk


def test_shard_bounds_cover_everything():
    from audio_sheet_retrieval_b200.dist import shard_bounds
    for n in (1, 7, 1000, 10 ** 8 + 3):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
