"""ASR_TC_STATS=1 python tools/tc_one.py NQ [ROWS]: one pre-filter call, device counters on stderr."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
nq = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
g = torch.Generator(device="cuda").manual_seed(1)
D = torch.randn((n, 32), generator=g, device="cuda"); D /= D.norm(dim=1, keepdim=True)
Q = torch.randn((nq, 32), generator=g, device="cuda")
db = EmbeddingDB(D)
for _ in range(2):
    db.topk_device(Q, 25)
torch.cuda.synchronize()
