"""Timing of asr_topk over the regimes the dispatch rule separates (env switches select path / partition)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
def timed(db, q, iters=5):
    s = torch.empty((q.shape[0], 25), device="cuda"); i = torch.empty((q.shape[0], 25), dtype=torch.int64, device="cuda")
    for _ in range(2): db.topk_device(q, 25, out_scores=s, out_idx=i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): db.topk_device(q, 25, out_scores=s, out_idx=i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
g = torch.Generator(device="cuda").manual_seed(1)
D1 = torch.randn((1000000, 32), generator=g, device="cuda"); D1 = D1 / D1.norm(dim=1, keepdim=True)
D7 = torch.randn((10000000, 32), generator=g, device="cuda"); D7 = D7 / D7.norm(dim=1, keepdim=True)
rows = torch.randint(0, 1000000, (10000,), generator=g, device="cuda")
Q10k = D1[rows] + 0.12 * torch.randn((10000, 32), generator=g, device="cuda")
Q100 = torch.randn((100, 32), generator=g, device="cuda")
db1, db7, db8 = EmbeddingDB(D1), EmbeddingDB(D7, normalise_in_place=True), EmbeddingDB(D1[:125000].contiguous())
env = " ".join("%s=%s" % (k[4:], v) for k, v in sorted(os.environ.items()) if k.startswith("ASR_T"))
print("[%s] 10k x 1M %.2f ms | 10k x 125k %.2f ms | 1e7 rows: Q100 %.3f Q64 %.3f Q32 %.3f Q16 %.3f Q8 %.3f ms | 1e6 rows: Q100 %.3f Q16 %.3f ms" % (
    env, timed(db1, Q10k, 3), timed(db8, Q10k, 3), timed(db7, Q100), timed(db7, Q100[:64].contiguous()), timed(db7, Q100[:32].contiguous()),
    timed(db7, Q100[:16].contiguous()), timed(db7, Q100[:8].contiguous()), timed(db1, Q100), timed(db1, Q100[:16].contiguous())), flush=True)
