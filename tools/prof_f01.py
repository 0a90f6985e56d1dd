import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_sheet_retrieval_b200 import _lib, network
from audio_sheet_retrieval_b200.models import mutopia_ccal_cont as model
from audio_sheet_retrieval_b200.params import load_params
PKL = os.path.join(ROOT, "tests", "golden", "params_synth_mutopia_ccal_cont.pkl")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
layers = model.build_model(show_model=False)
net = layers[0].net
net.max_batch = n
network.set_all_param_values(layers, load_params(PKL))
enc = net.encoder(1, model.prepare.asr_prepare_mode)
X = torch.where(torch.rand((n, 1, 160, 200), device="cuda") < 0.19, torch.randint(0, 120, (n, 1, 160, 200), device="cuda", dtype=torch.uint8),
                torch.full((1,), 255, device="cuda", dtype=torch.uint8))
codes = torch.empty((n, 32), device="cuda")
for _ in range(2):
    enc.embed_device(X, codes=codes)
torch.cuda.synchronize()
