"""Top-k time against a 10^6-row DB for the query counts one rank sees when 10 000 queries are split over 1/2/4/8 ranks."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
g = torch.Generator(device="cuda").manual_seed(1)
D = torch.randn((1000000, 32), generator=g, device="cuda"); D = D / D.norm(dim=1, keepdim=True)
rows = torch.randint(0, 1000000, (10000,), generator=g, device="cuda")
Q = D[rows] + 0.12 * torch.randn((10000, 32), generator=g, device="cuda")
db = EmbeddingDB(D)
out = []
for nq in (10000, 5000, 2500, 1300, 1200, 640, 320):
    q = Q[:nq].contiguous()
    s = torch.empty((nq, 25), device="cuda"); i = torch.empty((nq, 25), dtype=torch.int64, device="cuda")
    for _ in range(2): db.topk_device(q, 25, out_scores=s, out_idx=i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): db.topk_device(q, 25, out_scores=s, out_idx=i)
    e1.record(); torch.cuda.synchronize()
    out.append("%d: %.3f ms" % (nq, e0.elapsed_time(e1) / 5))
print(" | ".join(out), flush=True)
