#!/usr/bin/env python
"""Small driver for ncu: top-k over a 1e7-row DB at Q = 1, 16, 100 (2 launches each)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_sheet_retrieval_b200.retrieval import EmbeddingDB  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
g = torch.Generator(device="cuda").manual_seed(1)
D = torch.randn((rows, 32), generator=g, device="cuda")
db = EmbeddingDB(D)
for nq in (1, 16, 100):
    q = torch.randn((nq, 32), generator=g, device="cuda")
    for _ in range(2):
        db.topk_device(q, 25)
torch.cuda.synchronize()
print("done")
