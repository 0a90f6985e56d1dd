#!/usr/bin/env python
"""No parameter pickle is shipped for the full-resolution model `mutopia_ccal_cont`
(12/24/48/48 filters).  This writes a synthetic one in the reference's pickle format: He-uniform
weights, BN statistics calibrated on synthetic inputs (oracle/encoders.py:synth_params), seed 23.

    python tests/golden/make_synth_params.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle.encoders import synth_params  # noqa: E402
from audio_sheet_retrieval_b200.params import save_params  # noqa: E402

if __name__ == "__main__":
    params = synth_params("mutopia_ccal_cont", seed=23, calib_n=16)
    out = os.path.join(HERE, "params_synth_mutopia_ccal_cont.pkl")
    save_params(out, params)
    print("wrote", out, sum(p.size for p in params), "parameters")
