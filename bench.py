#!/usr/bin/env python
"""bench.py - headline benchmark of the retrieval hot path (one JSON line on rank 0).

Workload (BASELINE.json configs[1]): the full-resolution encoder pair `mutopia_ccal_cont`
(12/24/48/48 filters; synthetic weights in the reference's pickle format, no weights are shipped
for this model) batch-embedding `--pairs` synthetic snippet pairs (sheet 160x200 uint8 +
spectrogram 92x42 float32) + CCA projection + length norm on every GPU.  One step = one pass
over all pairs.  Weak scaling: every rank embeds its own `--pairs` pairs, no data-path collective.

  value      pairs/s with inputs resident in HBM (asr_encoder_embed, chunks of --max-batch)
  e2e        pairs/s through the host-buffer C-ABI entry RetrievalWrapper uses
             (asr_encoder_embed_host): pinned host inputs, H2D + D2H inside the timed region
  roofline   tcgen05 conv kernel (layers 1..7, both views): algorithmic TFLOP/s vs measured bf16 peak,
             kernel time measured with CUDA events on the launching stream inside the timed region
  retrieval  fused normalise+dot+top-k over a resident DB: HBM GB/s (Q=1, k=25) and queries/s
  cpu_baseline  the oracle port of the reference's CPU path on a bounded sample

`--impl reference` times that CPU path alone (rank 0 only).
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
METRIC = "snippet pairs embedded per second (both encoder branches + CCA projection)"
UNIT = "pairs/s"


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


def cpu_reference_pairs_per_s(n_sample, steps=1, warmup=0):
    """The reference's CPU path (oracle port: torch-CPU fp32 NCHW conv/BN/ELU/pool, batch 100,
    CCA projection, length norm; asr/retrieval_wrapper.py + model graph) on all host threads."""
    import torch
    from oracle.encoders import OracleNet, load_param_list, synth_inputs
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1; use every host core
    net = OracleNet(MODEL, load_param_list(PKL))
    X1, X2 = synth_inputs(min(n_sample, 64), seed=3)
    reps = (n_sample + len(X1) - 1) // len(X1)
    X1 = np.concatenate([X1] * reps)[:n_sample]
    X2 = np.concatenate([X2] * reps)[:n_sample]
    times = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        net.compute_view_1(X1)
        net.compute_view_2(X2)
        dt = time.perf_counter() - t
        if i >= warmup:
            times.append(dt)
    return n_sample / float(np.median(times)), float(np.median(times)), torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    n_sample = args.cpu_sample or 1024
    v, dt, cores = cpu_reference_pairs_per_s(n_sample, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: %s full-res encoders, batch-embed synthetic pairs + CCA projection" % MODEL,
                   "pairs_per_step": n_sample, "note": "bounded sample of the %d-pair job" % args.pairs},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d pairs per step, torch-CPU fp32 oracle port (Theano/Lasagne cannot be installed)" % n_sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def retrieval_leg(torch, dev, pk, rows, quick=False):
    """Fused top-k: HBM-regime bandwidth (Q=1) and throughput at larger query counts."""
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    g = torch.Generator(device=dev).manual_seed(1)
    D = torch.randn((rows, 32), generator=g, device=dev)
    D = D / D.norm(dim=1, keepdim=True)
    db = EmbeddingDB(D)
    out = {"db_rows": rows, "k": 25, "db_bytes": rows * 128}

    def timed(q, iters):
        s = torch.empty((q.shape[0], 25), device=dev)
        i = torch.empty((q.shape[0], 25), dtype=torch.int64, device=dev)
        for _ in range(3):
            db.topk_device(q, 25, out_scores=s, out_idx=i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            db.topk_device(q, 25, out_scores=s, out_idx=i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    for nq in ((1, 16, 100) if not quick else (1,)):
        q = torch.randn((nq, 32), generator=g, device=dev)
        ms = timed(q, 10 if nq <= 16 else 5)
        gbs = rows * 128 / (ms * 1e-3) / 1e9
        out["q%d" % nq] = {"ms": ms, "queries_per_s": nq / (ms * 1e-3), "algorithmic_gbs": gbs,
                           "hbm_frac": gbs / pk["hbm_gbs"]}
    db.close()
    return out


def piece_id_leg(torch, dev, rank, world, quick=False):
    """Config 4: 10k queries (100 recordings x 100 windows) vs a 1M-row DB sharded over the ranks,
    k=25, all-gather + merge + vote."""
    import torch.distributed as dist
    from audio_sheet_retrieval_b200.dist import ShardedDB, shard_bounds
    n_db, n_rec, win = 1000000, (100 if not quick else 10), 100
    g = torch.Generator(device=dev).manual_seed(1)
    D = torch.randn((n_db, 32), generator=g, device=dev)
    D = D / D.norm(dim=1, keepdim=True)
    ids = (torch.arange(n_db, device=dev) // 100).to(torch.int32)
    g2 = torch.Generator(device=dev).manual_seed(2)
    true_piece = torch.randint(0, n_db // 100, (n_rec,), generator=g2, device=dev)
    rows = (true_piece[:, None] * 100 + torch.randint(0, 100, (n_rec, win), generator=g2, device=dev)).view(-1)
    Q = D[rows] + 0.12 * torch.randn((n_rec * win, 32), generator=g2, device=dev)
    lo, hi = shard_bounds(n_db, rank, world)
    sdb = ShardedDB(D[lo:hi].contiguous(), lo, row_ids_global=ids, group=None)
    for _ in range(2):
        pid, cnt = sdb.identify(Q, n_rec, 5, 25)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 3
    for _ in range(iters):
        pid, cnt = sdb.identify(Q, n_rec, 5, 25)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    acc = float((pid[:, 0].long() == true_piece).float().mean().item())
    return {"db_rows": n_db, "queries": n_rec * win, "k": 25, "ms": float(ms.item()),
            "queries_per_s": n_rec * win / (float(ms.item()) * 1e-3), "top1_piece_accuracy": acc,
            "regime": "tcgen05 tf32 pre-filter + exact fp32 re-scoring (bit-exact results); bound by the TMEM read-out of "
                      "the 128x256 score tiles; DB sharded over %d GPU(s)" % world}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=100000, help="pairs per GPU per step")
    ap.add_argument("--max-batch", type=int, default=4096, help="samples per launch (activation arena: ~2.7 MB per pair)")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="pairs per CPU step (default: 1024 per step for --impl reference, 4096 for the cpu_baseline leg)")
    ap.add_argument("--db-rows", type=int, default=10000000)
    ap.add_argument("--skip-extras", action="store_true", help="only the headline leg (used under ncu)")
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
    g = torch.Generator(device=dev).manual_seed(23 + rank)
    # sheet-like uint8 (mostly white, ~19 % dark) and spectrogram-like float32 (sparse, mean ~0.1)
    X1 = torch.where(torch.rand((n, 1, 160, 200), generator=g, device=dev) < 0.19,
                     torch.randint(0, 120, (n, 1, 160, 200), generator=g, device=dev, dtype=torch.uint8),
                     torch.full((1,), 255, device=dev, dtype=torch.uint8))
    X2 = torch.relu(torch.randn((n, 1, 92, 42), generator=g, device=dev) - 1.2) * 0.8
    codes1 = torch.empty((n, 32), device=dev)
    codes2 = torch.empty((n, 32), device=dev)

    def step_device():
        for s in range(0, n, mb):
            e1.embed_device(X1[s:s + mb], codes=codes1[s:s + mb])
            e2.embed_device(X2[s:s + mb], codes=codes2[s:s + mb])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    e1.set_timing(True); e2.set_timing(True)
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
    clk = clocks.stop()
    t1, t2 = e1.get_timing(), e2.get_timing()
    e1.set_timing(False); e2.set_timing(False)
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total.item()) / args.steps
    value = world * n / (ms_step * 1e-3)

    # roofline of the dominant kernel family: the tcgen05 convolutions of layers 1-7 (7 launches per embed call)
    conv_flops_pair = flops_pair - 2.0 * 9 * 12 * (160 * 200 + 92 * 42) - 2.0 * 48 * 32 * (10 * 12 + 5 * 2)
    conv_ms = t1["ms_conv_tc"] + t2["ms_conv_tc"]
    conv_launches = 7 * (t1["calls"] + t2["calls"])
    achieved = conv_flops_pair * n * args.steps / (conv_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "conv3x3_rows_kernel + conv3x3_tc_kernel (tcgen05 implicit GEMM, layers 1-7 of both branches)",
                "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                "peak_source": pk["src"] + ", sustained figure (kernel timed inside a long step)",
                # dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the 14 conv launches of one
                # chunk pair: ncu captures profiles/r1_launches_mb4096.csv (14945 MB / 14) and r1_launches_final.csv
                # (1024-pair chunks, 3397 MB / 14); other chunk sizes: not captured
                "traffic": {4096: 1067.5e6, 1024: 242.6e6}.get(mb),
                "avg_launch_ms": conv_ms / max(conv_launches, 1), "launches": conv_launches,
                "algorithmic_flops_per_pair": conv_flops_pair,
                "share_of_step": conv_ms / (ms_step * args.steps),
                "other_ms_per_step": {"layer0_tcgen05_toeplitz": (t1["ms_layer0"] + t2["ms_layer0"]) / args.steps,
                                      "head": (t1["ms_head"] + t2["ms_head"]) / args.steps}}

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
    for _ in range(k_e2e):
        c1h = e1.embed_host(h1)
        c2h = e2.embed_host(h2)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / k_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * n_e2e / float(dt.item())
    same = bool(np.array_equal(c1h, codes1[:n_e2e].cpu().numpy()) and np.array_equal(c2h, codes2[:n_e2e].cpu().numpy()))
    # what the host link gives a plain pinned copy of the same bytes (explains e2e vs value)
    hv = h1.view(-1)[:min(h1.numel(), 1 << 30)]
    dv = torch.empty_like(hv, device=dev)
    dv.copy_(hv, non_blocking=True); torch.cuda.synchronize()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record(); dv.copy_(hv, non_blocking=True); l1.record(); torch.cuda.synchronize()
    link_gbs = hv.numel() / (l0.elapsed_time(l1) * 1e-3) / 1e9
    del dv
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_e2e * (160 * 200 + 92 * 42 * 4),
           "d2h_bytes_per_step": n_e2e * 2 * 32 * 4, "steps": k_e2e,
           "api": "asr_encoder_embed_host (what RetrievalWrapper.compute_view_1/2 call)", "codes_equal_device_path": same,
           "h2d_gbs_used": n_e2e * (160 * 200 + 92 * 42 * 4) / float(dt.item()) / 1e9,
           "h2d_gbs_plain_pinned_copy": link_gbs}
    del h1, h2

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "configs[1]: %s full-res encoders (12/24/48/48), %d synthetic pairs per GPU per step "
                               "+ CCA projection + length norm" % (MODEL, n),
                   "pairs_per_gpu": n, "max_batch": mb, "weights": "synthetic, reference pickle format (none shipped for this model)",
                   "l2": "inputs (%.1f GB per GPU) are larger than L2; no flush needed" % ((n * (32000 + 15456)) / 1e9),
                   "algorithmic_mflop_per_pair": flops_pair / 1e6},
        "algorithmic_tflops": flops_pair * value / 1e12,
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
    }
    del X1, X2
    torch.cuda.empty_cache()
    if not args.skip_extras:
        try:
            r = retrieval_leg(torch, dev, pk, args.db_rows)
            r["peak_gbs"] = pk["hbm_gbs"]
            line["retrieval"] = r
            line["roofline_retrieval"] = {"bound": "hbm", "kernel": "topk_stream_kernel (Q=1, k=25, %d-row fp32 DB)" % args.db_rows,
                                          "achieved": r["q1"]["algorithmic_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                                          "frac": r["q1"]["hbm_frac"],
                                          # ncu dram__bytes_read.sum of one launch over the 10^7-row DB = 1.280 GB,
                                          # exactly the algorithmic bytes (profiles/r1_topk_ncu_summary.md)
                                          "traffic": 1.28e9 if args.db_rows == 10000000 else None}
            line["piece_identification"] = piece_id_leg(torch, dev, rank, world)
        except Exception as ex:  # report, never hide
            line["retrieval_error"] = repr(ex)
    if rank == 0 and world == 1 and not args.skip_extras:
        try:
            cpu_n = args.cpu_sample or 4096             # ~10-15 s of CPU work on 16 cores
            cpu_reference_pairs_per_s(128, steps=1, warmup=0)   # warm the thread pool / allocator on a small sample
            v, dt_cpu, cores = cpu_reference_pairs_per_s(cpu_n, steps=1, warmup=0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d pairs (%.1f s), torch-CPU fp32 oracle port of the reference path" % (cpu_n, dt_cpu)}
        except Exception as ex:
            line["cpu_baseline"] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
