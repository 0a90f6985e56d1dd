"""Fused layer-0 + layer-1 kernel vs the unfused path: layer-1 activations, codes, and event timings.
Run under gpurun:  python tools/f01_check.py [n_time]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_sheet_retrieval_b200 import _lib, network  # noqa: E402
from audio_sheet_retrieval_b200.models import mutopia_ccal_cont as model  # noqa: E402
from audio_sheet_retrieval_b200.params import load_params  # noqa: E402

PKL = os.path.join(ROOT, "tests", "golden", "params_synth_mutopia_ccal_cont.pkl")
n_time = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
layers = model.build_model(show_model=False)
net = layers[0].net
net.max_batch = n_time
network.set_all_param_values(layers, load_params(PKL))
enc = net.encoder(1, model.prepare.asr_prepare_mode)
print("fusion available/in effect:", enc.fusion, flush=True)
rng = np.random.RandomState(3)
for dtype in (np.uint8, np.float32):
    for n in (1, 5, 149, 300):
        X = rng.randint(0, 256, size=(n, 1, 160, 200)).astype(dtype)
        enc.set_fusion(1)
        c1 = enc.embed_host(X)
        a1 = enc.debug_activation(1, n)
        enc.set_fusion(0)
        c0 = enc.embed_host(X)
        a0 = enc.debug_activation(1, n)
        d = np.abs(a1 - a0)
        cos = (c1 * c0).sum(1).min()
        print("%s n=%d: layer-1 max diff %.4g (scale %.3g), frac differing %.4g, nan %d, min cos %.7f"
              % (dtype.__name__, n, d.max(), np.abs(a0).max(), (d > 0).mean(), int(np.isnan(a1).sum()), cos), flush=True)
        if d.max() > 0.05 * np.abs(a0).max():
            bad = np.argwhere(d > 0.05 * np.abs(a0).max())
            print("  first bad (n,c,y,x):", bad[:8].tolist(), "count", len(bad), flush=True)

# timing: device-resident u8 input
X = torch.randint(0, 256, (n_time, 1, 160, 200), dtype=torch.uint8, device="cuda")
codes = torch.empty((n_time, 32), device="cuda")
for fuse in (0, 1, 0, 1):
    enc.set_fusion(fuse)
    for _ in range(3):
        enc.embed_device(X, codes=codes)
    torch.cuda.synchronize()
    enc.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        enc.embed_device(X, codes=codes)
    e1.record()
    torch.cuda.synchronize()
    t = enc.get_timing()
    enc.set_timing(False)
    print("fuse=%d: %.3f ms per %d samples (layer0 %.3f, conv %.3f, head %.3f)"
          % (fuse, e0.elapsed_time(e1) / 10, n_time, t["ms_layer0"] / 10, t["ms_conv_tc"] / 10, t["ms_head"] / 10), flush=True)
