"""ASR_TC_STATS=1 python tools/tc_stats.py : one pre-filter call per regime with the kernel's device counters on stderr."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
g = torch.Generator(device="cuda").manual_seed(1)
D1 = torch.randn((1000000, 32), generator=g, device="cuda"); D1 = D1 / D1.norm(dim=1, keepdim=True)
D7 = torch.randn((10000000, 32), generator=g, device="cuda"); D7 = D7 / D7.norm(dim=1, keepdim=True)
rows = torch.randint(0, 1000000, (10000,), generator=g, device="cuda")
Q10k = D1[rows] + 0.12 * torch.randn((10000, 32), generator=g, device="cuda")
Q100 = torch.randn((100, 32), generator=g, device="cuda")
db1, db7 = EmbeddingDB(D1), EmbeddingDB(D7, normalise_in_place=True)
for name, db, q in (("10k x 1M", db1, Q10k), ("100 x 1e7", db7, Q100), ("16 x 1e7", db7, Q100[:16].contiguous())):
    s = torch.empty((q.shape[0], 25), device="cuda"); i = torch.empty((q.shape[0], 25), dtype=torch.int64, device="cuda")
    for _ in range(2):
        db.topk_device(q, 25, out_scores=s, out_idx=i)
    torch.cuda.synchronize()
    print(name, flush=True)
