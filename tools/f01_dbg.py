import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio_sheet_retrieval_b200 import _lib, network
from audio_sheet_retrieval_b200.models import mutopia_ccal_cont as model
from audio_sheet_retrieval_b200.params import load_params
PKL = os.path.join(ROOT, "tests", "golden", "params_synth_mutopia_ccal_cont.pkl")
layers = model.build_model(show_model=False)
net = layers[0].net
net.max_batch = 8
network.set_all_param_values(layers, load_params(PKL))
enc = net.encoder(1, model.prepare.asr_prepare_mode)
X = np.random.RandomState(3).randint(0, 256, size=(int(sys.argv[1]) if len(sys.argv) > 1 else 1, 1, 160, 200)).astype(np.uint8)
enc.set_fusion(0)
c0 = enc.embed_host(X); a0 = enc.debug_activation(1, len(X))
enc.set_fusion(1)
c1 = enc.embed_host(X); a1 = enc.debug_activation(1, len(X))
d = np.abs(a1 - a0)
print("max diff", d.max(), "scale", np.abs(a0).max(), "frac", (d > 0).mean(), "cos", (c0 * c1).sum(1).min())
