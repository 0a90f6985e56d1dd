import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
g = torch.Generator(device="cuda").manual_seed(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
D1 = torch.randn((n, 32), generator=g, device="cuda"); D1 = D1 / D1.norm(dim=1, keepdim=True)
rows = torch.randint(0, n, (10000,), generator=g, device="cuda")
Q = D1[rows] + 0.12 * torch.randn((10000, 32), generator=g, device="cuda")
db = EmbeddingDB(D1)
for _ in range(2):
    db.topk_device(Q, 25)
torch.cuda.synchronize()
