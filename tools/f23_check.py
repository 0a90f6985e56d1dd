"""Fused layer-2 + layer-3 kernel: parity against the per-layer launches, then event timing of the sheet branch.

    python tools/f23_check.py [n_parity] [n_timing]
"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import _use_so  # noqa: F401
from audio_sheet_retrieval_b200 import _lib, network
from audio_sheet_retrieval_b200.models import mutopia_ccal_cont as model
from audio_sheet_retrieval_b200.params import load_params
PKL = os.path.join(ROOT, "tests", "golden", "params_synth_mutopia_ccal_cont.pkl")
n_par = int(sys.argv[1]) if len(sys.argv) > 1 else 333
n_time = int(sys.argv[2]) if len(sys.argv) > 2 else 4096


def make(n):
    layers = model.build_model(show_model=False)
    net = layers[0].net
    net.max_batch = n
    network.set_all_param_values(layers, load_params(PKL))
    return net.encoder(1, model.prepare.asr_prepare_mode)


def cos(a, b):
    return (a * b).sum(1) / np.sqrt((a * a).sum(1) * (b * b).sum(1))


DBG = os.environ.get("ASR_F23_DEBUG", "0") != "0"
if DBG:
    n_par = 4
enc = make(n_par)
print("fusion mask at create:", enc.fusion, flush=True)
g = torch.Generator(device="cuda").manual_seed(3)
X = torch.randint(0, 256, (n_par, 1, 160, 200), dtype=torch.uint8, device="cuda", generator=g)
# sheet-like: mostly white with dark strokes
X = torch.where(X > 48, torch.full_like(X, 255), X)
codes = {}
acts = {}
for mask in (1, 3, 1, 3):
    assert enc.set_fusion(mask) == mask, (mask, enc.fusion)
    c = torch.empty((n_par, 32), device="cuda")
    enc.embed_device(X, codes=c)
    torch.cuda.synchronize()
    codes[mask] = c.cpu().numpy()
    acts[mask] = enc.debug_activation(3, min(n_par, 160), path=_lib.PATH_TCGEN05)
    print("mask", mask, "done", flush=True)
d = np.abs(acts[3] - acts[1])
print("layer-3 activations: max |diff| %.4g (scale %.3g), fraction differing %.4g" % (d.max(), np.abs(acts[1]).max(), (d > 0).mean()))
per_sample = d.reshape(d.shape[0], -1).max(1)
print("worst samples:", np.argsort(per_sample)[-5:], np.sort(per_sample)[-5:])
print("codes: min cosine fused vs unfused %.7f" % cos(codes[3], codes[1]).min(), flush=True)
ok = DBG or (d.max() <= 2.0 ** -6 * np.abs(acts[1]).max() + 1e-3 and cos(codes[3], codes[1]).min() > 0.9999)
print("PARITY", "OK" if ok else "FAIL", flush=True)
enc.close()
if not ok or n_time <= 0:
    sys.exit(0 if ok else 1)

enc = make(n_time)
X = torch.randint(0, 256, (n_time, 1, 160, 200), dtype=torch.uint8, device="cuda")
c = torch.empty((n_time, 32), device="cuda")
for mask in ((3, 3) if DBG else (1, 3, 1, 3)):
    enc.set_fusion(mask)
    for _ in range(3):
        enc.embed_device(X, codes=c)
    torch.cuda.synchronize()
    enc.set_timing(True)
    for _ in range(10):
        enc.embed_device(X, codes=c)
    torch.cuda.synchronize()
    t = enc.get_timing()
    enc.set_timing(False)
    print("mask %d F23_DEBUG=%s: conv %.3f ms head %.3f ms per %d samples" % (
        mask, os.environ.get("ASR_F23_DEBUG", "-"), t["ms_conv_tc"] / 10, t["ms_head"] / 10, n_time), flush=True)
