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
print("items/SM=%s min_tiles=%s: 10k x 1M %.2f ms | 100 x 1e7 %.3f ms | 16 x 1e7 %.3f ms" % (
    os.environ.get("ASR_TC_ITEMS_PER_SM"), os.environ.get("ASR_TC_MIN_TILES"),
    timed(EmbeddingDB(D1), Q10k, 3), timed(EmbeddingDB(D7), Q100), timed(EmbeddingDB(D7), Q100[:16].contiguous())))
