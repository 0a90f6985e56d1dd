import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
from audio_sheet_retrieval_b200.utils.train_dcca_pool import eval_retrieval
rng = np.random.RandomState(0)
a = rng.normal(size=(2000, 32)).astype(np.float32); b = (a + 0.5 * rng.normal(size=(2000, 32))).astype(np.float32)
eval_retrieval(a[:100], b[:100])
for _ in range(3):
    t = time.perf_counter(); eval_retrieval(a, b); print("eval_retrieval %.2f ms" % ((time.perf_counter() - t) * 1e3))
bg = torch.as_tensor(b).cuda()
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); db = EmbeddingDB(bg); torch.cuda.synchronize(); t1 = time.perf_counter()
    r = db.ranks_device(torch.as_tensor(a).cuda()); torch.cuda.synchronize(); t2 = time.perf_counter(); db.close(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("create %.2f ms, ranks %.2f ms, close %.2f ms" % ((t1 - t) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
