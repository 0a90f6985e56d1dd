#!/usr/bin/env python
"""bench.py - headline benchmark of the retrieval hot path (one JSON line on rank 0).

Headline workload (BASELINE.json configs[1]): the full-resolution encoder pair `mutopia_ccal_cont`
(12/24/48/48 filters; synthetic weights in the reference's pickle format, no weights are shipped
for this model) batch-embedding `--pairs` synthetic snippet pairs (sheet 160x200 uint8 +
spectrogram 92x42 float32) + CCA projection + length norm on every GPU.  One step = one pass
over all pairs.  Weak scaling: every rank embeds its own `--pairs` pairs, no data-path collective.

  value      pairs/s with inputs resident in HBM (asr_encoder_embed, chunks of --max-batch)
  e2e        pairs/s through the host-buffer C-ABI entry RetrievalWrapper uses
             (asr_encoder_embed_host): pinned host inputs, H2D + D2H inside the timed region
  roofline   tcgen05 kernels (fused layer 0+1, layers 1..7, both views): algorithmic TFLOP/s vs measured bf16
             peak, kernel time measured with CUDA events on the launching stream inside the timed region

The other BASELINE configs ride along as extra keys of the same line (their sizes are BASELINE's):
  config1_rsz_eval       configs[0]: rsz model + tutorial pickle, 2000 pairs, R@k / MRR / MR with parity vs the oracle
  config3_refit          configs[2]: CCA refit on 25 000 latent pairs, rows sharded over the ranks, ONE all-reduce
  piece_identification   configs[3]: 10 000 queries vs a 10^6-row DB sharded over the ranks, k = 25, all-gather +
                         merge + vote; per-phase times; bit_exact = a 64-query sample equals the pinned-order oracle
  retrieval_sweep        configs[4]: 10^5 .. 10^8 rows x Q in {1, 16, 100}, DB sharded over the ranks, fraction of the
                         measured HBM bandwidth per cell
  cpu_baseline(s)        the oracle port of the reference's CPU path on a bounded sample (rank 0, N = 1 only)

`--impl reference` times the CPU port of the encoder path alone (rank 0 only), same generator and chunking.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL = "mutopia_ccal_cont"
PKL = os.path.join(ROOT, "tests", "golden", "params_synth_mutopia_ccal_cont.pkl")
PKL_RSZ = os.path.join(ROOT, "tests", "golden", "params_all_split_mutopia_full_aug.pkl")
METRIC = "snippet pairs embedded per second (both encoder branches + CCA projection)"
UNIT = "pairs/s"
L2_BYTES = 126e6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def make_inputs(torch, n, device, seed):
    """The synthetic workload of BOTH arms: sheet-like uint8 (mostly white, ~19 % dark pixels) and spectrogram-like
    float32 (sparse, mean ~0.1)."""
    g = torch.Generator(device=device).manual_seed(seed)
    X1 = torch.where(torch.rand((n, 1, 160, 200), generator=g, device=device) < 0.19,
                     torch.randint(0, 120, (n, 1, 160, 200), generator=g, device=device, dtype=torch.uint8),
                     torch.full((1,), 255, device=device, dtype=torch.uint8))
    X2 = torch.relu(torch.randn((n, 1, 92, 42), generator=g, device=device) - 1.2) * 0.8
    return X1, X2


# ----------------------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference's CPU path (bounded samples; reported baselines, never the target)
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_pairs_per_s(n_sample, steps=1, warmup=0, chunk=4096):
    """The reference's CPU path (oracle port: torch-CPU fp32 NCHW conv/BN/ELU/pool in batches of 100 as
    retrieval_wrapper.py:60, CCA projection, length norm) on all host threads, over the same generator and the same
    chunking (chunks of --max-batch pairs) as the GPU arm."""
    import torch
    from oracle.encoders import OracleNet, load_param_list
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1; use every host core
    net = OracleNet(MODEL, load_param_list(PKL))
    X1, X2 = make_inputs(torch, n_sample, "cpu", 23)
    X1, X2 = X1.numpy().astype(np.float32), X2.numpy()          # the reference pools hand out float32 holding 0..255
    times = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        for s in range(0, n_sample, chunk):
            net.compute_view_1(X1[s:s + chunk])
            net.compute_view_2(X2[s:s + chunk])
        dt = time.perf_counter() - t
        if i >= warmup:
            times.append(dt)
    return n_sample / float(np.median(times)), float(np.median(times)), torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    n_sample = args.cpu_sample or 1024
    v, dt, cores = cpu_reference_pairs_per_s(n_sample, steps=args.steps, warmup=args.warmup, chunk=args.max_batch)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: %s full-res encoders (12/24/48/48), synthetic pairs + CCA projection + length norm"
                               % MODEL,
                   "pairs_per_step": n_sample, "max_batch": args.max_batch,
                   "generator": "bench.make_inputs (the generator of the GPU arm)",
                   "note": "bounded sample of the %d-pair job: a rate, so it compares with the GPU arm's pairs/s" % args.pairs},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d pairs per step, torch-CPU fp32 oracle port (the reference is Python 2 + Theano/Lasagne: "
                                   "not installable here)" % n_sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_retrieve_loop(db, queries, n_candidates):
    """The reference's per-query loop (audio_sheet_server.py:230-234 -> :530-551): cdist(DB, q, 'cosine') over the whole
    DB + a FULL argsort, one query at a time.  Returns (seconds per query, indices (nq, n_candidates))."""
    from oracle.search import retrieve_ids_ref
    ids = np.zeros(len(db), np.int64)
    out = np.empty((len(queries), n_candidates), np.int64)
    t = time.perf_counter()
    for i in range(len(queries)):
        _, out[i] = retrieve_ids_ref(db, ids, queries[i:i + 1], n_candidates)
    return (time.perf_counter() - t) / len(queries), out


# ----------------------------------------------------------------------------------------------------------------
# GPU legs
# ----------------------------------------------------------------------------------------------------------------
def timed_ms(torch, fn, iters, flush=None):
    """Mean device time of fn() over `iters` calls, each timed by its own event pair; `flush` (a big buffer) is
    rewritten before every call so that nothing of the previous call is left in L2."""
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return float(np.mean(times))


def unit_rows(torch, n, dev, seed):
    """(n,32) fp32 unit-norm Gaussian directions, generated in slabs (a 1e8-row DB is 12.8 GB)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    D = torch.empty((n, 32), device=dev)
    slab = 1 << 24
    for s in range(0, n, slab):
        d = torch.randn((min(slab, n - s), 32), generator=g, device=dev)
        D[s:s + slab] = d / d.norm(dim=1, keepdim=True)
    return D


def retrieval_sweep_leg(torch, dist, dev, pk, rank, world, rows_list, quick=False):
    """Config 5: rows x Q sweep, the DB sharded over the ranks (contiguous shards, queries replicated), local top-k +
    ONE all-gather + merge inside the timed region, max over ranks.  hbm_frac = whole-DB bytes / time / (N x peak)."""
    from audio_sheet_retrieval_b200.dist import ShardedDB, shard_bounds
    flush = torch.empty(int(2 * L2_BYTES), dtype=torch.uint8, device=dev)
    cells = []
    for rows in rows_list:
        lo, hi = shard_bounds(rows, rank, world)
        shard = unit_rows(torch, hi - lo, dev, 1000 + rank)
        sdb = ShardedDB(shard, lo, group=None, normalise_in_place=True)      # donated buffer: no second copy
        g = torch.Generator(device=dev).manual_seed(7)
        for nq in (1, 16, 100):
            q = torch.randn((nq, 32), generator=g, device=dev)
            for _ in range(3):
                sdb.topk_device(q, 25)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ms = timed_ms(torch, lambda: sdb.topk_device(q, 25), 3 if quick else 8, flush)
            t = torch.tensor([ms], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            gbs = rows * 128 / (ms * 1e-3) / 1e9
            cells.append({"rows": rows, "q": nq, "ms": ms, "queries_per_s": nq / (ms * 1e-3), "algorithmic_gbs": gbs,
                          "hbm_frac": gbs / (world * pk["hbm_gbs"]),
                          "shard_fits_l2": bool((hi - lo) * 128 <= L2_BYTES)})
            if nq > 1 and rows == rows_list[-1]:
                # Full-size parity without a host copy of the DB: the default dispatch (the tensor-core pre-filter at this
                # size) must return, index for index and bit for bit, what the exact streaming kernel returns -- two
                # independent kernels, the exact one pinned against the oracle at the sizes the oracle can do.
                s_a, i_a = sdb.topk_device(q, 25)
                os.environ["ASR_TOPK_PATH"] = "exact"
                try:
                    s_b, i_b = sdb.topk_device(q[:4].contiguous(), 25)
                finally:
                    del os.environ["ASR_TOPK_PATH"]
                same = torch.tensor([int(torch.equal(i_a[:4], i_b) and torch.equal(s_a[:4], s_b))], device=dev)
                if world > 1:
                    dist.all_reduce(same, op=dist.ReduceOp.MIN)
                cells[-1]["default_path_equals_exact_kernel_on_4_queries"] = bool(same.item())
        sdb.local.close()
        del sdb, shard
        torch.cuda.empty_cache()
    return {"k": 25, "n_gpus": world, "cells": cells,
            "timing": "CUDA events per call, L2 flushed (252 MB rewritten) before every call, max over ranks; "
                      "N > 1: local top-k + one NCCL all-gather + merge kernel inside the timed region",
            "directions": "audio->sheet and sheet->audio are the same computation (which codes are the DB)",
            "scaling_note": "strong scaling over the ranks (DB rows sharded): every call pays the local top-k launch, ONE all-gather "
                            "of the (Q, k) lists and the merge launch, ~0.1 ms together whatever the shard size, so a cell scales "
                            "until its shard's stream time approaches that (10^8 rows, Q = 1: 1.90 ms on one GPU, 0.24 ms of "
                            "streaming per shard on eight)",
            "peak_gbs_per_gpu": pk["hbm_gbs"]}


def piece_id_leg(torch, dist, dev, rank, world, with_cpu, quick=False):
    """Config 4: 10k queries (100 recordings x 100 windows) vs a 1M-row DB sharded over the ranks, k=25, all-gather +
    merge + vote; per-phase device times; a 64-query sample checked against the pinned-order oracle."""
    from audio_sheet_retrieval_b200.dist import ShardedDB, shard_bounds
    n_db, n_rec, win = 1000000, (100 if not quick else 10), 100
    g = torch.Generator(device=dev).manual_seed(1)
    D = torch.randn((n_db, 32), generator=g, device=dev)
    D = D / D.norm(dim=1, keepdim=True)
    ids = (torch.arange(n_db, device=dev) // 100).to(torch.int32)
    g2 = torch.Generator(device=dev).manual_seed(2)
    true_piece = torch.randint(0, n_db // 100, (n_rec,), generator=g2, device=dev)
    rows = (true_piece[:, None] * 100 + torch.randint(0, 100, (n_rec, win), generator=g2, device=dev)).view(-1)
    Q = (D[rows] + 0.12 * torch.randn((n_rec * win, 32), generator=g2, device=dev)).contiguous()
    lo, hi = shard_bounds(n_db, rank, world)
    sdb = ShardedDB(D[lo:hi].contiguous(), lo, row_ids_global=ids, group=None)
    for _ in range(2):
        pid, cnt = sdb.identify(Q, n_rec, 5, 25)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    iters = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        pid, cnt = sdb.identify(Q, n_rec, 5, 25)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # per-phase breakdown (one more call with events between the phases)
    ev = []
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    v0.record()
    pid, cnt = sdb.identify(Q, n_rec, 5, 25, events=ev)
    v1.record()
    torch.cuda.synchronize()
    if ev:
        ph = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]),
                           ev[3].elapsed_time(v1)], device=dev)
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
        phases = dict(zip(("local_topk_ms", "allgather_ms", "merge_ms", "vote_ms"), [float(x) for x in ph.tolist()]))
    else:
        phases = {"local_topk_and_vote_ms": v0.elapsed_time(v1)}
    acc = float((pid[:, 0].long() == true_piece).float().mean().item())
    ms_sharded = float(ms.item())
    out = {"db_rows": n_db, "queries": n_rec * win, "k": 25, "ms": ms_sharded,
           "queries_per_s": n_rec * win / (ms_sharded * 1e-3), "top1_piece_accuracy": acc, "phases": phases,
           "strategy": "DB rows sharded over the ranks",
           "regime": "tcgen05 tf32 pre-filter + exact fp32 re-scoring; DB sharded over %d GPU(s); "
                     "one all-gather of [scores | indices] chunks, merge kernel reads the rank-major chunks" % world}
    if world > 1:
        # The other way to use N GPUs when the DB fits on each of them (128 MB here): every rank holds the whole DB and
        # identifies its share of the RECORDINGS (retrieval + vote), one all-gather of (recordings x top_k) results.
        from audio_sheet_retrieval_b200.dist import ReplicatedDB
        rdb = ReplicatedDB(D, ids, group=None)
        for _ in range(2):
            pid_r, cnt_r = rdb.identify(Q, n_rec, 5, 25)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(iters):
            pid_r, cnt_r = rdb.identify(Q, n_rec, 5, 25)
        e1.record()
        torch.cuda.synchronize()
        ms_r = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(ms_r, op=dist.ReduceOp.MAX)
        evr = []
        dist.barrier()
        rdb.identify(Q, n_rec, 5, 25, events=evr)
        torch.cuda.synchronize()
        phr = torch.tensor([evr[0].elapsed_time(evr[1]), evr[1].elapsed_time(evr[2])], device=dev)
        dist.all_reduce(phr, op=dist.ReduceOp.MAX)
        ms_rep = float(ms_r.item())
        out["by_strategy"] = {
            "db_sharded": {"ms": ms_sharded, "phases": phases},
            "recordings_sharded_db_replicated": {
                "ms": ms_rep, "phases": {"local_topk_and_vote_ms": float(phr[0].item()), "allgather_results_ms": float(phr[1].item())},
                "same_result_as_db_sharded": bool(torch.equal(pid_r, pid) and torch.equal(cnt_r, cnt))}}
        if ms_rep < ms_sharded:
            out.update({"ms": ms_rep, "queries_per_s": n_rec * win / (ms_rep * 1e-3),
                        "phases": out["by_strategy"]["recordings_sharded_db_replicated"]["phases"],
                        "regime": "tcgen05 tf32 pre-filter + exact fp32 re-scoring; every rank holds the whole DB and identifies "
                                  "its share of the %d recordings; one all-gather of (recordings x top_k) results" % n_rec,
                        "strategy": "recordings sharded over the ranks, DB (128 MB) replicated: no candidate exchange, no merge"})
        rdb.local.close()
    # parity at full size: a 64-query sample of the merged result (indices AND scores) vs the pinned-order C oracle
    s_all, i_all = sdb.topk_device(Q, 25)
    if rank == 0:
        from oracle import clib
        sample = np.linspace(0, Q.shape[0] - 1, 64).astype(int)
        Dh = D.cpu().numpy()
        s_ref, i_ref = clib.topk(Q[sample].cpu().numpy(), Dh, 25)
        out["bit_exact"] = bool((i_all[sample].cpu().numpy() == i_ref).all() and (s_all[sample].cpu().numpy() == s_ref).all())
        out["bit_exact_sample"] = "64 of %d queries, indices and scores after the merge, vs oracle/c pinned-order top-k" % Q.shape[0]
        if with_cpu:
            nq_cpu = 100 if not quick else 5
            qs = np.linspace(0, Q.shape[0] - 1, nq_cpu).astype(int)
            spq, idx_cpu = cpu_retrieve_loop(Dh, Q[qs].cpu().numpy(), 25)
            same = float((idx_cpu == i_all[qs].cpu().numpy()).mean())
            out["cpu_baseline"] = {"value": 1.0 / spq, "unit": "queries/s", "cores": 1, "kind": "port",
                                   "sample": "%d of the %d queries (%.1f s): scipy cdist(DB, q, 'cosine') over the 10^6 rows + "
                                             "full argsort per query, the loop of audio_sheet_server.py:230-240"
                                             % (nq_cpu, Q.shape[0], spq * nq_cpu),
                                   "index_agreement_with_gpu": same}
    sdb.local.close()
    return out


def streaming_leg(torch, dev):
    """The streaming loop of the reference's server (audio_sheet_server.py:83-211 without GUI / microphone): one
    spectrogram column per frame, embed the running 92 x 42 window, retrieve 25 candidates from a 10^6-row sheet DB,
    vote over the last 100 frames.  Reports the frame rate the reference prints ("Server is running at ... fps")."""
    from audio_sheet_retrieval_b200.audio_sheet_server import AudioSheetServer
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    import contextlib
    import io
    rng = np.random.RandomState(3)
    srv = AudioSheetServer()
    with contextlib.redirect_stdout(io.StringIO()):
        srv.initialize_embedding_network(model, PKL_RSZ)
    n_db = 1000000
    g = torch.Generator(device=dev).manual_seed(4)
    D = torch.randn((n_db, 32), generator=g, device=dev)
    srv.set_sheet_db(D.cpu().numpy(), np.arange(n_db) // 100, dict((i, "piece%d" % i) for i in range(n_db // 100)))
    spec = np.abs(rng.normal(0, 0.3, (92, 42 + 300))).astype(np.float32)
    srv.run(spec[:, :42 + 20], top_k=5, n_candidates=25, running_frames=100, verbose=False)      # warm-up
    names, probs, fps = srv.run(spec, top_k=5, n_candidates=25, running_frames=100, verbose=False)
    return {"frames": 300, "db_rows": n_db, "n_candidates": 25, "running_frames": 100, "frames_per_s": fps,
            "ms_per_frame": 1e3 / fps if fps else None,
            "api": "AudioSheetServer.run -> process_frame (NumPy window in: compute_view_2, Q = 1 top-k, vote kernel); the "
                   "reference's tutorial audio runs at 20 frames per second of signal"}


def config1_leg(torch, dev, with_cpu):
    """Config 1: rsz model + tutorial pickle, 2000 synthetic pairs through RetrievalWrapper + eval_retrieval;
    R@k / MRR / median rank next to the oracle's on the same inputs (tolerance of the north star: 0.5 % absolute)."""
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    from audio_sheet_retrieval_b200.retrieval_wrapper import RetrievalWrapper
    from audio_sheet_retrieval_b200.utils.mutopia_data import SyntheticPairPool
    from audio_sheet_retrieval_b200.utils.train_dcca_pool import eval_retrieval
    import contextlib
    import io
    n = 2000
    pool = SyntheticPairPool(2000, seed=25)
    X1, X2 = pool[np.linspace(0, 1999, n).astype(int)]
    with contextlib.redirect_stdout(io.StringIO()):
        w = RetrievalWrapper(model, PKL_RSZ, prepare_view_1=model.prepare, prepare_view_2=None)
    w.compute_view_1(X1[:200]); w.compute_view_2(X2[:200])
    torch.cuda.synchronize()
    t_embed = t_eval = float("inf")
    for _ in range(3):                      # 20 + 20 latency-bound calls of 100 rows each (the reference's batching): best of 3
        t = time.perf_counter()
        c1, c2 = w.compute_view_1(X1), w.compute_view_2(X2)
        t_embed = min(t_embed, time.perf_counter() - t)
    eval_retrieval(c1[:100], c2[:100])
    for _ in range(3):
        t = time.perf_counter()
        mr, med, md, hr, mrr = eval_retrieval(c1, c2)
        t_eval = min(t_eval, time.perf_counter() - t)
    out = {"model": "mutopia_ccal_cont_rsz (tutorial pickle)", "pairs": n, "embed_pairs_per_s": n / t_embed,
           "embed_ms": t_embed * 1e3, "eval_retrieval_ms": t_eval * 1e3,
           "api": "RetrievalWrapper.compute_view_1/2 (NumPy in, NumPy out, 100 rows per call like the reference) + eval_retrieval; best of 3",
           "metrics": {"mrr": float(mrr), "median_rank": float(med), "mean_rank": float(mr),
                       "recall_at_k": dict((str(k), 100.0 * hr[k] / n) for k in (1, 5, 10, 25))}}
    if with_cpu:
        from oracle import metrics
        from oracle.encoders import OracleNet, load_param_list
        torch.set_num_threads(os.cpu_count() or 1)
        onet = OracleNet("mutopia_ccal_cont_rsz", load_param_list(PKL_RSZ))
        t = time.perf_counter()
        r1, r2 = onet.compute_view_1(X1), onet.compute_view_2(X2)
        t_cpu = time.perf_counter() - t
        t = time.perf_counter()
        omr, omed, omd, ohr, omrr = metrics.eval_retrieval_ref(r1, r2)
        t_cpu_eval = time.perf_counter() - t
        cos = float(min(((c1 * r1).sum(1)).min(), ((c2 * r2).sum(1)).min()))
        out["parity"] = {"min_code_cosine_vs_oracle": cos, "mrr_abs_diff": abs(float(mrr) - float(omrr)),
                         "median_rank_abs_diff": abs(float(med) - float(omed)),
                         "recall_at_k_abs_diff_pct": dict((str(k), abs(100.0 * (hr[k] - ohr[k]) / n)) for k in (1, 5, 10, 25)),
                         "within_tolerance": bool(cos >= 0.999 and abs(float(mrr) - float(omrr)) <= 0.005
                                                  and all(abs(100.0 * (hr[k] - ohr[k]) / n) <= 0.5 for k in (1, 5, 10, 25)))}
        out["cpu_baseline"] = {"value": n / t_cpu, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "all 2000 pairs (%.1f s embed, %.2f s cdist + per-row argsort): torch-CPU fp32 "
                                         "oracle port of run_eval.py:102-174" % (t_cpu, t_cpu_eval),
                               "eval_retrieval_ms": t_cpu_eval * 1e3}
    return out


def config3_leg(torch, dist, dev, rank, world, with_cpu):
    """Config 3: CCA('svd') refit on 25 000 latent pairs, rows sharded over the ranks: one fused Gram pass per rank,
    ONE all-reduce of 3137 doubles, replicated single-CTA Jacobi solve."""
    from audio_sheet_retrieval_b200 import _lib
    from audio_sheet_retrieval_b200.dist import shard_bounds
    from audio_sheet_retrieval_b200.utils.cca import CCA, cca_solve_device
    n = 25000
    rng = np.random.RandomState(23)
    Z = rng.normal(size=(n, 32))
    H1 = (Z @ rng.normal(size=(32, 32)) * 0.05 + 0.3).astype(np.float32)
    H2 = (Z @ rng.normal(size=(32, 32)) * 0.05 + 0.02 * rng.normal(size=(n, 32)) - 0.1).astype(np.float32)
    lo, hi = shard_bounds(n, rank, world)
    h1, h2 = torch.as_tensor(H1[lo:hi]).to(dev), torch.as_tensor(H2[lo:hi]).to(dev)
    group = dist.group.WORLD if world > 1 else None
    c = CCA()
    for _ in range(3):
        sig = c.fit(h1, h2, group=group)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    iters = 10
    t = time.perf_counter()
    for _ in range(iters):
        sig = c.fit(h1, h2, group=group)
    torch.cuda.synchronize()
    wall = torch.tensor([(time.perf_counter() - t) / iters * 1e3], device=dev, dtype=torch.float64)
    # device-side breakdown: accumulate / all-reduce / solve
    sums = torch.zeros(_lib.CCA_NSUMS + 1, dtype=torch.float64, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    if world > 1:
        dist.barrier()
    ev[0].record()
    _lib.check(_lib.lib.asr_cca_accumulate_counted(_lib.dptr(h1), _lib.dptr(h2), int(h1.shape[0]), None, None,
                                                   _lib.dptr(sums), _lib.stream_ptr()))
    ev[1].record()
    if world > 1:
        dist.all_reduce(sums)
    ev[2].record()
    cca_solve_device(sums, _lib.CCA_COUNT_ON_DEVICE)
    ev[3].record()
    torch.cuda.synchronize()
    ph = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])], device=dev,
                      dtype=torch.float64)
    if world > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    out = {"samples": n, "n_gpus": world, "fit_wall_ms": float(wall.item()),
           "device_ms": {"gram_pass": float(ph[0]), "allreduce_3137_f64": float(ph[1]), "jacobi_solve": float(ph[2])},
           "regime": "latency-bound: 6.4 MB of latents, a 25 KB all-reduce, one single-CTA fp64 solve; fit_wall_ms is the "
                     "user-visible CCA.fit (includes the D2H copies of U, V, m, sigma)"}
    if with_cpu and rank == 0:
        from oracle import cca as occa
        o = occa.CCA()
        o.fit(H1, H2)
        t = time.perf_counter()
        for _ in range(5):
            sig_ref = o.fit(H1, H2)
        t_cpu = (time.perf_counter() - t) / 5
        out["cpu_baseline"] = {"value": t_cpu * 1e3, "unit": "ms per fit", "cores": os.cpu_count(), "kind": "port",
                               "sample": "the full 25 000 x 32 fit: NumPy fp32 sgemm covariances + SciPy sqrtm/inv + LAPACK svd "
                                         "(oracle/cca.py, pinned against the reference's own CCA.fit)"}
        out["sigma_abs_diff_vs_port"] = float(np.abs(np.asarray(sig) - sig_ref).max())
    return out


def read_traffic(mb):
    """Measured DRAM traffic of the tcgen05 kernels (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged
    over the launches of one chunk pair): written by tools/launch_table.py from the ncu launch list of this command."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        return None, "no capture committed"
    d = json.load(open(p))
    e = d.get(str(mb))
    if not e:
        return None, "no capture for max_batch %d" % mb
    return e["bytes_per_conv_launch"], "%s (%s)" % (e["source"], d.get("command", ""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=100000, help="pairs per GPU per step")
    ap.add_argument("--max-batch", type=int, default=4096, help="samples per launch (activation arena: ~1.7 MB per pair)")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="pairs per CPU step (default: 1024 per step for --impl reference, 4096 for the cpu_baseline leg)")
    ap.add_argument("--db-rows", type=int, default=10000000)
    ap.add_argument("--sweep-max-rows", type=int, default=100000000)
    ap.add_argument("--skip-extras", action="store_true", help="only the headline leg (used under ncu)")
    ap.add_argument("--quick", action="store_true", help="small extras (smoke runs)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from audio_sheet_retrieval_b200 import _lib, network
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont as model
    from audio_sheet_retrieval_b200.params import load_params
    pk = peaks()

    layers = model.build_model(show_model=False)
    net = layers[0].net
    net.max_batch = args.max_batch
    network.set_all_param_values(layers, load_params(PKL))
    e1 = net.encoder(1, model.prepare.asr_prepare_mode)
    e2 = net.encoder(2, _lib.PREP_NONE)
    flops_pair = e1.flops_per_sample + e2.flops_per_sample

    n, mb = args.pairs, args.max_batch
    X1, X2 = make_inputs(torch, n, dev, 23 + rank)
    codes1 = torch.empty((n, 32), device=dev)
    codes2 = torch.empty((n, 32), device=dev)

    # One stream per branch: the two encoders are independent, and a branch's persistent kernels leave SMs idle at
    # their tails that the other branch's next kernel can use (measured +4.5 % over one stream).
    two_streams = os.environ.get("ASR_BENCH_TWO_STREAMS", "1") == "1"
    s_a, s_b = (torch.cuda.Stream(), torch.cuda.Stream()) if two_streams else (None, None)

    def step_device():
        if not two_streams:
            for s in range(0, n, mb):
                e1.embed_device(X1[s:s + mb], codes=codes1[s:s + mb])
                e2.embed_device(X2[s:s + mb], codes=codes2[s:s + mb])
            return
        cur = torch.cuda.current_stream()
        s_a.wait_stream(cur); s_b.wait_stream(cur)
        for s in range(0, n, mb):
            e1.embed_device(X1[s:s + mb], codes=codes1[s:s + mb], stream=s_a)
            e2.embed_device(X2[s:s + mb], codes=codes2[s:s + mb], stream=s_b)
        cur.wait_stream(s_a); cur.wait_stream(s_b)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total.item()) / args.steps
    value = world * n / (ms_step * 1e-3)
    # Kernel durations for the roofline: CUDA events around [layer 0] [tcgen05 conv kernels] [head] on the stream they are
    # launched on.  With one stream per branch the two branches' kernels overlap and their intervals would count the
    # overlap twice, so the durations come from a second pass of the same steps on ONE stream, right after the timed
    # region and under the same clock sampling; `value` is never taken from this pass.
    roof_steps = max(1, min(args.steps, 3))
    e1.set_timing(True); e2.set_timing(True)
    rv0, rv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rv0.record()
    for _ in range(roof_steps):
        for s in range(0, n, mb):
            e1.embed_device(X1[s:s + mb], codes=codes1[s:s + mb])
            e2.embed_device(X2[s:s + mb], codes=codes2[s:s + mb])
    rv1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    t1, t2 = e1.get_timing(), e2.get_timing()
    e1.set_timing(False); e2.set_timing(False)
    ms_roof_step = rv0.elapsed_time(rv1) / roof_steps

    # Roofline of the dominant kernel family: the tcgen05 kernels.  Layer 0 of a branch counts with them when it runs
    # inside the fused layer-0 + layer-1 kernel (its time cannot be separated any more); an unfused layer 0 and the
    # fp32 head are reported beside them.
    f1, f2 = e1.fusion & 1, e2.fusion & 1
    l0_flops = {1: 2.0 * 9 * 12 * 160 * 200, 2: 2.0 * 9 * 12 * 92 * 42}
    head_flops = 2.0 * 48 * 32 * (10 * 12 + 5 * 2)
    conv_flops_pair = flops_pair - head_flops - (0 if f1 else l0_flops[1]) - (0 if f2 else l0_flops[2])
    conv_ms = t1["ms_conv_tc"] + t2["ms_conv_tc"]
    # tcgen05 launches per call: layers 1..7 (the fused layer-0 + layer-1 kernel counts as layer 1), one fewer where
    # layers 2 + 3 run as one kernel
    conv_launches = ((7 - ((e1.fusion >> 1) & 1)) * t1["calls"]) + ((7 - ((e2.fusion >> 1) & 1)) * t2["calls"])
    achieved = conv_flops_pair * n * roof_steps / (conv_ms * 1e-3) / 1e12
    traffic, traffic_src = read_traffic(mb)
    roofline = {"bound": "tensor",
                "kernel": "l01_fused_kernel (prepare + layer 0 + layer 1, sheet branch) + l23_fused_kernel (layers 2 + 3, sheet "
                          "branch) + conv3x3_rows_kernel + conv3x3_tc_kernel (tcgen05 implicit GEMM, layers 1-7 of both branches)",
                "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                "peak_source": pk["src"] + ", sustained figure (kernel timed inside a long step)",
                "traffic": traffic, "traffic_source": traffic_src,
                "avg_launch_ms": conv_ms / max(conv_launches, 1), "launches": conv_launches,
                "algorithmic_flops_per_pair": conv_flops_pair,
                "fused_layer01": {"sheet_branch": bool(f1), "spectrogram_branch": bool(f2)},
                "fused_layer23": {"sheet_branch": bool(e1.fusion & 2), "spectrogram_branch": bool(e2.fusion & 2)},
                "timed_in": "a second pass of %d step(s) on one stream right after the timed region (%.2f ms per step there; the "
                            "timed region runs the branches on two streams: %.2f ms per step)" % (roof_steps, ms_roof_step, ms_step),
                "share_of_step": conv_ms / (ms_roof_step * roof_steps),
                "whole_step_frac": flops_pair * n / (ms_step * 1e-3) / 1e12 / pk["bf16_sustained"],
                "other_ms_per_step": {"layer0_unfused_tcgen05_toeplitz": (t1["ms_layer0"] + t2["ms_layer0"]) / roof_steps,
                                      "head": (t1["ms_head"] + t2["ms_head"]) / roof_steps}}

    # ---- e2e: host buffers through the C-ABI entry the wrapper uses ----
    n_e2e = n
    h1 = torch.empty((n_e2e, 1, 160, 200), dtype=torch.uint8, pin_memory=True)
    h2 = torch.empty((n_e2e, 1, 92, 42), dtype=torch.float32, pin_memory=True)
    h1.copy_(X1[:n_e2e]); h2.copy_(X2[:n_e2e])
    torch.cuda.synchronize()
    e1.embed_host(h1[:mb * 4]); e2.embed_host(h2[:mb * 4])          # warm the staging buffers / streams
    barrier()
    k_e2e = max(1, args.steps)
    t0 = time.perf_counter()
    t_v1 = 0.0
    # the two branches are independent calls (compute_view_1 / compute_view_2): issue them from two host threads so
    # that the spectrogram branch's copies overlap the sheet branch's kernels (each handle has its own streams)
    res, t_br = {}, {}

    def run_branch(name, enc, h):
        ta = time.perf_counter()
        res[name] = enc.embed_host(h)
        t_br[name] = t_br.get(name, 0.0) + time.perf_counter() - ta

    step_ms = []
    for _ in range(k_e2e):
        ts = time.perf_counter()
        th = threading.Thread(target=run_branch, args=("v2", e2, h2))
        th.start()
        run_branch("v1", e1, h1)
        th.join()
        step_ms.append((time.perf_counter() - ts) * 1e3)
    torch.cuda.synchronize()
    c1h, c2h = res["v1"], res["v2"]
    dt = torch.tensor([(time.perf_counter() - t0) / k_e2e], device=dev, dtype=torch.float64)
    ms_v1, ms_v2 = t_br["v1"] / k_e2e * 1e3, t_br["v2"] / k_e2e * 1e3
    dt_all = [float(dt.item())]
    if world > 1:
        gathered = [torch.zeros_like(dt) for _ in range(world)]
        dist.all_gather(gathered, dt)
        dt_all = [float(g.item()) for g in gathered]
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * n_e2e / float(dt.item())
    same = bool(np.array_equal(c1h, codes1[:n_e2e].cpu().numpy()) and np.array_equal(c2h, codes2[:n_e2e].cpu().numpy()))
    # what the host link gives a plain pinned copy of the same bytes (explains e2e vs value), all ranks copying at once
    hv = h1.view(-1)[:min(h1.numel(), 1 << 30)]
    dv = torch.empty_like(hv, device=dev)
    dv.copy_(hv, non_blocking=True)
    barrier()
    la, lb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    la.record(); dv.copy_(hv, non_blocking=True); lb.record(); torch.cuda.synchronize()
    link = torch.tensor([hv.numel() / (la.elapsed_time(lb) * 1e-3) / 1e9], device=dev)
    if world > 1:
        dist.all_reduce(link, op=dist.ReduceOp.MIN)
    link_gbs = float(link.item())
    del dv
    h2d = n_e2e * (160 * 200 + 92 * 42 * 4)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": n_e2e * 2 * 32 * 4, "steps": k_e2e,
           "api": "asr_encoder_embed_host (what RetrievalWrapper.compute_view_1/2 call)", "codes_equal_device_path": same,
           "ms_per_step_sheet_branch": ms_v1, "ms_per_step_spectrogram_branch": ms_v2,
           "concurrency": "the two branch calls run concurrently from two host threads",
           "ms_per_step_by_rank": [x * 1e3 for x in dt_all], "ms_of_each_step_rank0": step_ms,
           "h2d_gbs_used": h2d / float(dt.item()) / 1e9,
           "h2d_gbs_plain_pinned_copy": link_gbs,
           "host_link_note": "min over ranks of a plain pinned copy with all %d rank(s) copying at once; e2e needs "
                             "value x 47.5 KB per pair = %.1f GB/s per GPU, so e2e is bound by this link whenever it is "
                             "lower (all GPUs of the box share the host's PCIe root complexes)" % (world, value / world * 47456 / 1e9)}
    del h1, h2

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "configs[1]: %s full-res encoders (12/24/48/48), %d synthetic pairs per GPU per step "
                               "+ CCA projection + length norm" % (MODEL, n),
                   "pairs_per_gpu": n, "max_batch": mb, "weights": "synthetic, reference pickle format (none shipped for this model)",
                   "generator": "bench.make_inputs (also the generator of --impl reference)",
                   "l2": "inputs (%.1f GB per GPU) are larger than L2; no flush needed" % ((n * (32000 + 15456)) / 1e9),
                   "algorithmic_mflop_per_pair": flops_pair / 1e6},
        "algorithmic_tflops": flops_pair * value / 1e12,
        "streams": "one CUDA stream per branch" if two_streams else "one CUDA stream",
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
    }
    del X1, X2, codes1, codes2
    e1.close(); e2.close()
    torch.cuda.empty_cache()
    with_cpu = rank == 0 and world == 1
    if not args.skip_extras:
        legs = [
            ("retrieval_sweep", lambda: retrieval_sweep_leg(
                torch, dist, dev, pk, rank, world,
                [r for r in (100000, 1000000, 10000000, 100000000) if r <= args.sweep_max_rows], quick=args.quick)),
            ("piece_identification", lambda: piece_id_leg(torch, dist, dev, rank, world, with_cpu, quick=args.quick)),
            ("config3_refit", lambda: config3_leg(torch, dist, dev, rank, world, with_cpu)),
        ]
        if world == 1:
            legs.append(("config1_rsz_eval", lambda: config1_leg(torch, dev, with_cpu)))
            legs.append(("streaming_identification", lambda: streaming_leg(torch, dev)))
        for name, fn in legs:
            try:
                line[name] = fn()
            except Exception as ex:  # report, never hide
                line[name + "_error"] = repr(ex)
        sw = line.get("retrieval_sweep")
        if sw:
            big = [c for c in sw["cells"] if c["q"] == 1]
            if big:
                top = max(big, key=lambda c: c["rows"])
                line["roofline_retrieval"] = {"bound": "hbm", "kernel": "topk_stream_kernel (Q=1, k=25, %d-row fp32 DB over %d GPU(s))"
                                                                        % (top["rows"], world),
                                              "achieved": top["algorithmic_gbs"], "peak": pk["hbm_gbs"] * world, "unit": "GB/s",
                                              "frac": top["hbm_frac"],
                                              "traffic": None, "traffic_note": "ncu dram__bytes_read.sum of one launch over a 10^7-row DB = "
                                              "1.280 GB = the algorithmic bytes (profiles/r1_topk_ncu_summary.md)"}
    if with_cpu and not args.skip_extras:
        try:
            cpu_n = args.cpu_sample or 4096             # ~10-15 s of CPU work on 16 cores
            cpu_reference_pairs_per_s(128, steps=1, warmup=0)   # warm the thread pool / allocator on a small sample
            v, dt_cpu, cores = cpu_reference_pairs_per_s(cpu_n, steps=1, warmup=0, chunk=mb)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d pairs (%.1f s), torch-CPU fp32 oracle port of the reference path, the generator "
                                              "of the GPU arm" % (cpu_n, dt_cpu)}
        except Exception as ex:
            line["cpu_baseline"] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
