#!/usr/bin/env python
"""Throughput of the shipped `mutopia_ccal_cont_rsz` model (tutorial weights) on device-resident inputs."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_sheet_retrieval_b200 import _lib, network  # noqa: E402
from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model  # noqa: E402
from audio_sheet_retrieval_b200.params import load_params  # noqa: E402

n, mb = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 1024
layers = model.build_model(False)
net = layers[0].net
net.max_batch = mb
network.set_all_param_values(layers, load_params(os.path.join(os.path.dirname(__file__), "..", "tests", "golden",
                                                               "params_all_split_mutopia_full_aug.pkl")))
e1, e2 = net.encoder(1, _lib.PREP_SCALE_HALF), net.encoder(2, _lib.PREP_NONE)
g = torch.Generator(device="cuda").manual_seed(0)
X1 = torch.randint(0, 256, (n, 1, 160, 200), generator=g, device="cuda", dtype=torch.uint8)
X2 = torch.rand((n, 1, 92, 42), generator=g, device="cuda")
c1, c2 = torch.empty((n, 32), device="cuda"), torch.empty((n, 32), device="cuda")


def step():
    for s in range(0, n, mb):
        e1.embed_device(X1[s:s + mb], codes=c1[s:s + mb])
        e2.embed_device(X2[s:s + mb], codes=c2[s:s + mb])


step(); torch.cuda.synchronize()
e1.set_timing(True); e2.set_timing(True)
t = time.perf_counter(); step(); step(); torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 2
t1, t2 = e1.get_timing(), e2.get_timing()
fl = e1.flops_per_sample + e2.flops_per_sample
print("rsz: %.0f pairs/s, %.1f algorithmic TFLOP/s; per-step ms v1 %s v2 %s" % (n / dt, fl * n / dt / 1e12,
      {k: round(v / 2, 2) for k, v in t1.items()}, {k: round(v / 2, 2) for k, v in t2.items()}))
