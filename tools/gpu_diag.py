#!/usr/bin/env python
"""First-contact diagnostics on a B200: run each stage in isolation, print what matched, and dump
raw activations of the tcgen05 path vs the fp32 path to gpurun_out/ for offline analysis."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def stage(name):
    def deco(fn):
        def run():
            t = time.time()
            try:
                fn()
                print("[diag] %-28s OK   (%.1fs)" % (name, time.time() - t), flush=True)
            except Exception:
                print("[diag] %-28s FAIL (%.1fs)" % (name, time.time() - t), flush=True)
                traceback.print_exc()
        return run
    return deco


@stage("topk small")
def d_topk():
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    from oracle import clib
    rng = np.random.RandomState(0)
    D = rng.normal(size=(5000, 32)).astype(np.float32)
    Q = rng.normal(size=(20, 32)).astype(np.float32)
    s, i = EmbeddingDB(D).topk(Q, 25)
    s_ref, i_ref = clib.topk(Q, D, 25)
    print("   idx equal:", (i == i_ref).mean(), "score equal:", (s == s_ref).mean())
    assert (i == i_ref).all() and (s == s_ref).all()


@stage("cca solve")
def d_cca():
    from audio_sheet_retrieval_b200.utils.cca import CCA
    from oracle import cca as occa
    H1, H2 = occa.synth_latents(3000, seed=1)
    sig = CCA().fit(H1, H2)
    o = occa.CCA()
    ref = o.fit(H1, H2)
    print("   sigma max err:", np.abs(sig - ref).max())
    assert np.abs(sig - ref).max() < 1e-5


def _enc(model_name, view):
    import importlib
    from audio_sheet_retrieval_b200 import network, _lib
    from audio_sheet_retrieval_b200.params import load_params
    pkl = {"mutopia_ccal_cont_rsz": "params_all_split_mutopia_full_aug.pkl",
           "mutopia_ccal_cont": "params_synth_mutopia_ccal_cont.pkl"}[model_name]
    model = importlib.import_module("audio_sheet_retrieval_b200.models." + model_name)
    layers = model.build_model(False)
    net = layers[0].net
    net.max_batch = 8
    network.set_all_param_values(layers, load_params(os.path.join(ROOT, "tests", "golden", pkl)))
    mode = model.prepare.asr_prepare_mode if view == 1 else _lib.PREP_NONE
    return net.encoder(view, mode)


def _enc_stage(model_name, view):
    from audio_sheet_retrieval_b200 import _lib
    from oracle.encoders import synth_inputs
    X1, X2 = synth_inputs(3, seed=4)
    X = X1 if view == 1 else X2
    enc = _enc(model_name, view)
    c_fp = enc.embed_host(X, path=_lib.PATH_FP32)
    ref = [enc.debug_activation(l, 3, path=_lib.PATH_FP32) for l in range(8)]
    print("   fp32 path done", flush=True)
    c_tc = enc.embed_host(X, path=_lib.PATH_TCGEN05)
    tc = [enc.debug_activation(l, 3, path=_lib.PATH_TCGEN05) for l in range(8)]
    bad = False
    for l in range(8):
        err = np.abs(tc[l] - ref[l]).max()
        print("   layer %d shape %s max|ref| %.3f max err %.4f" % (l, ref[l].shape, np.abs(ref[l]).max(), err), flush=True)
        bad = bad or not (err < 0.05 * np.abs(ref[l]).max() + 0.02)
    cos = (c_tc * c_fp).sum(1)
    print("   codes cosine tc vs fp32:", cos)
    if bad:
        np.savez_compressed(os.path.join(OUT, "diag_%s_v%d.npz" % (model_name, view)),
                            **{"ref%d" % l: ref[l] for l in range(3)}, **{"tc%d" % l: tc[l] for l in range(3)})
        raise AssertionError("tcgen05 path deviates; activations dumped")


for _m in ("mutopia_ccal_cont", "mutopia_ccal_cont_rsz"):
    for _v in (2, 1):
        globals()["d_enc_%s_%d" % (_m, _v)] = stage("encoder %s v%d" % (_m, _v))(
            (lambda m, v: (lambda: _enc_stage(m, v)))(_m, _v))

if __name__ == "__main__":
    import torch
    print("[diag] device:", torch.cuda.get_device_name(0), flush=True)
    d_topk()
    d_cca()
    for _m in ("mutopia_ccal_cont", "mutopia_ccal_cont_rsz"):
        for _v in (2, 1):
            globals()["d_enc_%s_%d" % (_m, _v)]()
