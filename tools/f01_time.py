"""Event timing of the sheet branch with the fused layer-0 + layer-1 kernel (env switches select variants)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import _use_so  # noqa: F401
from audio_sheet_retrieval_b200 import _lib, network
from audio_sheet_retrieval_b200.models import mutopia_ccal_cont as model
from audio_sheet_retrieval_b200.params import load_params
PKL = os.path.join(ROOT, "tests", "golden", "params_synth_mutopia_ccal_cont.pkl")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
layers = model.build_model(show_model=False)
net = layers[0].net
net.max_batch = n
network.set_all_param_values(layers, load_params(PKL))
enc = net.encoder(1, model.prepare.asr_prepare_mode)
X = torch.randint(0, 256, (n, 1, 160, 200), dtype=torch.uint8, device="cuda")
codes = torch.empty((n, 32), device="cuda")
for _ in range(3):
    enc.embed_device(X, codes=codes)
torch.cuda.synchronize()
enc.set_timing(True)
for _ in range(10):
    enc.embed_device(X, codes=codes)
torch.cuda.synchronize()
t = enc.get_timing()
print("env F01_DEBUG=%s VARIANT=%s FUSE=%s: layer0 %.3f conv %.3f head %.3f ms per %d" % (
    os.environ.get("ASR_F01_DEBUG", "-"), os.environ.get("ASR_F01_VARIANT", "-"), os.environ.get("ASR_FUSE01", "-"),
    t["ms_layer0"] / 10, t["ms_conv_tc"] / 10, t["ms_head"] / 10, n), flush=True)
