"""Event timing of both encoder branches (full-res model), conv kernels only, per n samples (env switches select variants)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import _use_so  # noqa: F401
from audio_sheet_retrieval_b200 import _lib, network
from audio_sheet_retrieval_b200.models import mutopia_ccal_cont as model
from audio_sheet_retrieval_b200.params import load_params
PKL = os.path.join(ROOT, "tests", "golden", "params_synth_mutopia_ccal_cont.pkl")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
layers = model.build_model(show_model=False)
net = layers[0].net
net.max_batch = n
network.set_all_param_values(layers, load_params(PKL))
tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("ASR_"))
for view, shape, dt in ((1, (n, 1, 160, 200), torch.uint8), (2, (n, 1, 92, 42), torch.float32)):
    enc = net.encoder(view, model.prepare.asr_prepare_mode if view == 1 else _lib.PREP_NONE)
    X = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda").to(dt)
    if view == 2:
        X = X / 255.0
    codes = torch.empty((n, 32), device="cuda")
    for _ in range(3):
        enc.embed_device(X, codes=codes)
    torch.cuda.synchronize()
    enc.set_timing(True)
    for _ in range(10):
        enc.embed_device(X, codes=codes)
    torch.cuda.synchronize()
    t = enc.get_timing()
    print("[%s] view %d: layer0 %.3f conv %.3f head %.3f ms per %d (fusion %d)" % (
        tag, view, t["ms_layer0"] / 10, t["ms_conv_tc"] / 10, t["ms_head"] / 10, n, enc.fusion), flush=True)
