#!/usr/bin/env python
"""Golden vectors for the contrastive cosine loss: executes the REFERENCE's own
`get_contrastive_cos_loss` (audio_sheet_retrieval/models/objectives.py:30-69) on NumPy arrays.

The function is written against `theano.tensor as T`; every operation it uses has a NumPy twin
(`T.identity_like`, `T.repeat`, `T.clip`, ndarray `.dot/.diagonal/.nonzero/.mean`), so its source text
is sliced out of the reference file where it lies and executed with a three-name shim for `T`.
Run in the build container only (needs /root/reference):  python tests/golden/make_golden_loss.py
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, _slice_def, _src  # noqa: E402


def load_loss_factory():
    code = _slice_def(_src("models/objectives.py"), "get_contrastive_cos_loss")
    T = types.SimpleNamespace(identity_like=lambda D: np.eye(D.shape[0], dtype=D.dtype), repeat=np.repeat, clip=np.clip)
    g = {"T": T}
    exec(compile(code, "ref:get_contrastive_cos_loss", "exec"), g)
    return g["get_contrastive_cos_loss"]


def unit(x):
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def main():
    factory = load_loss_factory()
    out = {}
    rng = np.random.RandomState(2024)
    cases = []
    for n, noise in ((100, 0.6), (37, 0.15), (2, 0.5), (128, 2.0)):
        a = rng.randn(n, 32)
        lv1, lv2 = unit(a), unit(a + noise * rng.randn(n, 32))     # matching pairs correlate, as trained codes do
        cases.append((lv1, lv2))
    for ci, (lv1, lv2) in enumerate(cases):
        out["c%d_lv1" % ci], out["c%d_lv2" % ci] = lv1, lv2
        for weight, gamma, sym in ((1.0, 0.7, False), (1.0, 0.7, True), (0.35, 0.2, True)):
            key = "c%d_w%g_g%g_s%d" % (ci, weight, gamma, int(sym))
            out[key] = np.float64(factory(weight, gamma, symmetric=sym)(lv1, lv2))
    np.savez_compressed(os.path.join(OUT, "reference_contrastive_loss.npz"), **out)
    print("wrote reference_contrastive_loss.npz", sorted(k for k in out if "_w" in k)[:4], "...")


if __name__ == "__main__":
    main()
