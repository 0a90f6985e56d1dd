#!/usr/bin/env python
"""SURVEY 8(d) configs 3 and 5 on one GPU, CUDA-event timed:
  3. CCA refit on 25 000 latent pairs (means + 3136 fp64 sums + single-CTA solve): wall time and breakdown;
  5. top-k (k = 25) over 1e5 / 1e6 / 1e7 rows at Q = 1, 16, 100: ms per call and algorithmic GB/s (128 B per row per pass;
     1e5 rows = 12.8 MB is L2-resident, reported as such)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_sheet_retrieval_b200.retrieval import EmbeddingDB  # noqa: E402
from audio_sheet_retrieval_b200.utils.cca import CCA, cca_solve_device, cca_sums_device  # noqa: E402

dev = torch.device("cuda")
out = {}


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ---- config 3
rng = np.random.RandomState(23)
Z = rng.randn(25000, 32)
H1 = torch.as_tensor((Z @ rng.randn(32, 32) * 0.05 + 0.02 * rng.randn(25000, 32)).astype(np.float32)).to(dev)
H2 = torch.as_tensor((Z @ rng.randn(32, 32) * 0.05 + 0.02 * rng.randn(25000, 32)).astype(np.float32)).to(dev)
sh1, sh2 = H1.mean(0).contiguous(), H2.mean(0).contiguous()
sums = cca_sums_device(H1, H2, sh1, sh2)
out["config3"] = {
    "rows": 25000,
    "ms_sums_kernel": timed(lambda: cca_sums_device(H1, H2, sh1, sh2), 20),
    "ms_solve_kernel": timed(lambda: cca_solve_device(sums, 25000, sh1, sh2, 1e-3, 1e-3, 0.0, mode=0), 20),
    "ms_fit_end_to_end_incl_host_copies": timed(lambda: CCA().fit(H1, H2), 5),
    "bytes_read": 25000 * 2 * 128,
}
# ---- config 5
g = torch.Generator(device=dev).manual_seed(1)
rows_out = {}
for rows in (100000, 1000000, 10000000):
    D = torch.randn((rows, 32), generator=g, device=dev)
    D = D / D.norm(dim=1, keepdim=True)
    db = EmbeddingDB(D)
    r = {}
    for nq in (1, 16, 100):
        q = torch.randn((nq, 32), generator=g, device=dev)
        s = torch.empty((nq, 25), device=dev)
        i = torch.empty((nq, 25), dtype=torch.int64, device=dev)
        ms = timed(lambda: db.topk_device(q, 25, out_scores=s, out_idx=i), 20 if rows < 10000000 else 10)
        r["q%d" % nq] = {"ms": ms, "algorithmic_gbs": rows * 128 / ms / 1e6, "queries_per_s": nq / ms * 1e3}
    rows_out[str(rows)] = r
    db.close()
    del D
out["config5_one_gpu"] = rows_out
print(json.dumps(out))
