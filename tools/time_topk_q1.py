#!/usr/bin/env python
"""Event-timed top-k at Q = 1 over a 1e7-row DB (the roofline_retrieval case); ASR_TOPK_Q1 selects the variant."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_sheet_retrieval_b200.retrieval import EmbeddingDB  # noqa: E402

rows = 10000000
g = torch.Generator(device="cuda").manual_seed(1)
D = torch.randn((rows, 32), generator=g, device="cuda")
db = EmbeddingDB(D)
q = torch.randn((1, 32), generator=g, device="cuda")
for _ in range(5):
    db.topk_device(q, 25)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    db.topk_device(q, 25)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print("ASR_TOPK_Q1=%s: %.4f ms  %.1f GB/s" % (os.environ.get("ASR_TOPK_Q1", "0"), ms, rows * 128 / ms / 1e6))
