import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_sheet_retrieval_b200 import _lib
from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
rng = np.random.RandomState(0)
D = rng.normal(size=(4096, 32)).astype(np.float32); Q = rng.normal(size=(128, 32)).astype(np.float32)
db = EmbeddingDB(D)
q = torch.as_tensor(Q).cuda()
out = np.zeros((128, 256), np.float32)
_lib.check(_lib.lib.asr_debug_tc_scores(db.handle, _lib.dptr(q), 128, _lib.dptr(out)))
Dn = D / np.linalg.norm(D, axis=1, keepdims=True); Qn = Q / np.linalg.norm(Q, axis=1, keepdims=True)
ref = Qn @ Dn[:256].T
print("max abs err", np.abs(out - ref).max(), "nan", np.isnan(out).sum(), "out[0,:6]", out[0, :6], "ref[0,:6]", ref[0, :6])
# hypotheses: which (q', d') pairing reproduces out?
err_T = np.abs(out - ref.T[:128, :256]).max() if ref.shape[0] == ref.shape[1] else None
print("corr with ref", np.corrcoef(out.ravel(), ref.ravel())[0, 1])
for kk in range(4):
    part = Qn[:, kk*8:(kk+1)*8] @ Dn[:256, kk*8:(kk+1)*8].T
    print("partial k-block", kk, "corr", np.corrcoef(out.ravel(), part.ravel())[0, 1])
np.save(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "tc_dbg.npy"), out)
