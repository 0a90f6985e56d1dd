"""e2e (host buffers) vs device-resident timing of the sheet branch, fused / unfused."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import _use_so  # noqa: F401
from audio_sheet_retrieval_b200 import _lib, network
from audio_sheet_retrieval_b200.models import mutopia_ccal_cont as model
from audio_sheet_retrieval_b200.params import load_params
PKL = os.path.join(ROOT, "tests", "golden", "params_synth_mutopia_ccal_cont.pkl")
n, mb = 50000, 4096
layers = model.build_model(show_model=False)
net = layers[0].net
net.max_batch = mb
network.set_all_param_values(layers, load_params(PKL))
enc = net.encoder(1, model.prepare.asr_prepare_mode)
h = torch.randint(0, 256, (n, 1, 160, 200), dtype=torch.uint8).pin_memory()
X = h.cuda()
codes = torch.empty((n, 32), device="cuda")
for fuse in (1, 0, 1, 0):
    enc.set_fusion(fuse)
    enc.embed_host(h[:mb * 4])
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        enc.embed_host(h)
    te = (time.perf_counter() - t) / 3
    for s in range(0, n, mb):
        enc.embed_device(X[s:s + mb], codes=codes[s:s + mb])
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        for s in range(0, n, mb):
            enc.embed_device(X[s:s + mb], codes=codes[s:s + mb])
    torch.cuda.synchronize()
    td = (time.perf_counter() - t) / 3
    print("fuse=%d: e2e %.1f ms (%.0f/s, %.1f GB/s H2D), device %.1f ms (%.0f/s)" % (fuse, te * 1e3, n / te, n * 32000 / te / 1e9, td * 1e3, n / td), flush=True)
