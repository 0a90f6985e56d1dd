"""Encoders through the C ABI vs the CPU oracle.

Tolerances (BASELINE.json north_star): codes cosine >= 0.999 per row vs the fp32 oracle for the
tcgen05/bf16 path; the fp32 CUDA-core path must agree to fp32 round-off.
"""
import os

import numpy as np
import pytest
import torch

from oracle.encoders import OracleNet, load_param_list, synth_inputs

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PKL = {"mutopia_ccal_cont_rsz": os.path.join(GOLDEN, "params_all_split_mutopia_full_aug.pkl"),
       "mutopia_ccal_cont": os.path.join(GOLDEN, "params_synth_mutopia_ccal_cont.pkl")}
COS_TOL = 0.999


def _cos(a, b):
    return (a * b).sum(1) / np.linalg.norm(a, axis=1) / np.linalg.norm(b, axis=1)


def _net(model_name, max_batch=64):
    import importlib
    from audio_sheet_retrieval_b200 import network
    from audio_sheet_retrieval_b200.params import load_params
    model = importlib.import_module("audio_sheet_retrieval_b200.models." + model_name)
    layers = model.build_model(show_model=False)
    net = layers[0].net
    net.max_batch = max_batch
    network.set_all_param_values(layers, load_params(PKL[model_name]))
    return model, net


@pytest.mark.parametrize("model_name", ["mutopia_ccal_cont_rsz", "mutopia_ccal_cont"])
@pytest.mark.parametrize("view", [1, 2])
def test_fp32_path_matches_oracle_per_layer(model_name, view):
    from audio_sheet_retrieval_b200 import _lib
    model, net = _net(model_name)
    onet = OracleNet(model_name, load_param_list(PKL[model_name]))
    X1, X2 = synth_inputs(6, seed=3)
    X = X1 if view == 1 else X2
    mode = model.prepare.asr_prepare_mode if view == 1 else _lib.PREP_NONE
    enc = net.encoder(view, mode)
    codes, lats = enc.embed_host(X, want="both", path=_lib.PATH_FP32)
    Xp = onet.prepare(X) if view == 1 else X
    ref_c = onet.code(view, Xp).numpy()
    ref_l = onet.latent(view, Xp).numpy()
    np.testing.assert_allclose(lats, ref_l, atol=2e-5, rtol=1e-4)
    assert _cos(codes, ref_c).min() > 0.99999
    # per-layer activations
    import torch.nn.functional as F
    from oracle.encoders import _elu
    h = torch.as_tensor(Xp)
    for l in range(8):
        L = onet.views[view - 1][l]
        y = F.conv2d(h, torch.as_tensor(L["W"]), padding=1)
        y = (y - torch.as_tensor(L["mean"]).view(1, -1, 1, 1)) * torch.as_tensor(L["gamma"] * L["inv_std"]).view(1, -1, 1, 1) \
            + torch.as_tensor(L["beta"]).view(1, -1, 1, 1)
        y = _elu(y)
        if l % 2:
            y = F.max_pool2d(y, 2)
        got = enc.debug_activation(l, 6, path=_lib.PATH_FP32)
        assert got.shape == tuple(y.shape)
        np.testing.assert_allclose(got, y.numpy(), atol=5e-4, rtol=1e-3, err_msg="layer %d" % l)
        h = y


@pytest.mark.parametrize("model_name", ["mutopia_ccal_cont_rsz", "mutopia_ccal_cont"])
@pytest.mark.parametrize("view", [1, 2])
def test_tcgen05_path_per_layer_vs_fp32_path(model_name, view):
    """Layer by layer: bf16 tensor-core activations vs the fp32 kernels on the same device."""
    from audio_sheet_retrieval_b200 import _lib
    model, net = _net(model_name)
    X1, X2 = synth_inputs(5, seed=4)
    X = X1 if view == 1 else X2
    mode = model.prepare.asr_prepare_mode if view == 1 else _lib.PREP_NONE
    enc = net.encoder(view, mode)
    fused = enc.fusion
    c_fused, l_fused = enc.embed_host(X, want="both", path=_lib.PATH_TCGEN05)
    a1_fused = enc.debug_activation(1, 5, path=_lib.PATH_TCGEN05)
    enc.set_fusion(0)            # one launch per layer: every activation reaches memory
    c_tc, l_tc = enc.embed_host(X, want="both", path=_lib.PATH_TCGEN05)
    acts_tc = [enc.debug_activation(l, 5, path=_lib.PATH_TCGEN05) for l in range(8)]
    if fused:
        # layers 0 + 1 in one kernel: same arithmetic up to the accumulation order inside the tensor core
        # (8 instead of 12 rows per Toeplitz tile), so a few layer-1 outputs round to the neighbouring bf16
        assert model_name == "mutopia_ccal_cont" and view == 1
        d = np.abs(a1_fused - acts_tc[1])
        assert d.max() <= 2.0 ** -7 * np.abs(acts_tc[1]).max() + 1e-3, d.max()
        assert (d > 0).mean() < 0.02, (d > 0).mean()
        assert _cos(c_fused, c_tc).min() > 0.9999
        enc.set_fusion(1)
        enc.embed_host(X, path=_lib.PATH_TCGEN05)
        with pytest.raises(Exception):
            enc.debug_activation(0, 5, path=_lib.PATH_TCGEN05)     # never materialised when fused
        enc.set_fusion(0)
    if fused & 2:
        # layers 2 + 3 in one kernel, layers 0 + 1 unfused: layer 2 is the same arithmetic in another accumulation
        # order (row-stacked instead of raster tiles), so a few of its bf16 outputs differ by one ulp and layer 3
        # inherits that
        assert model_name == "mutopia_ccal_cont" and view == 1
        assert enc.set_fusion(2) == 2
        c_f23 = enc.embed_host(X, path=_lib.PATH_TCGEN05)
        a3_f23 = enc.debug_activation(3, 5, path=_lib.PATH_TCGEN05)
        d = np.abs(a3_f23 - acts_tc[3])
        assert d.max() <= 2.0 ** -6 * np.abs(acts_tc[3]).max() + 1e-3, d.max()
        assert (d > 0).mean() < 0.05, (d > 0).mean()
        assert _cos(c_f23, c_tc).min() > 0.9999
        with pytest.raises(Exception):
            enc.debug_activation(2, 5, path=_lib.PATH_TCGEN05)     # never materialised when fused
        enc.set_fusion(0)
        enc.embed_host(X, path=_lib.PATH_TCGEN05)
    c_fp, l_fp = enc.embed_host(X, want="both", path=_lib.PATH_FP32)
    for l in range(8):
        ref = enc.debug_activation(l, 5, path=_lib.PATH_FP32)
        err = np.abs(acts_tc[l] - ref).max()
        scale = np.abs(ref).max()
        assert err <= 0.05 * scale + 0.02, "layer %d: max err %.4g (scale %.3g)" % (l, err, scale)
    assert _cos(c_tc, c_fp).min() >= COS_TOL


@pytest.mark.parametrize("model_name,n", [("mutopia_ccal_cont_rsz", 150), ("mutopia_ccal_cont", 70)])
def test_codes_cosine_vs_oracle(model_name, n):
    """The headline parity number: per-row cosine >= 0.999 vs the fp32 oracle, both views, u8 and
    f32 sheet inputs, n > max_batch so the chunked host path is exercised."""
    from audio_sheet_retrieval_b200 import _lib
    model, net = _net(model_name, max_batch=64)
    onet = OracleNet(model_name, load_param_list(PKL[model_name]))
    X1, X2 = synth_inputs(n, seed=5)
    ref1, ref2 = onet.compute_view_1(X1), onet.compute_view_2(X2)
    e1 = net.encoder(1, model.prepare.asr_prepare_mode)
    e2 = net.encoder(2, _lib.PREP_NONE)
    c1 = e1.embed_host(X1)
    c1_u8 = e1.embed_host(X1.astype(np.uint8))
    c2 = e2.embed_host(X2)
    # u8 pixels enter the layer-0 GEMM as exact integers with 1/255 folded into the weights, fp32 pixels as
    # x/255 split into bf16 hi + lo: the same value up to ~2^-16, so a few layer-0 activations round to the
    # neighbouring bf16 and the two code sets differ like two bf16 pipelines do (both within COS_TOL of fp32)
    assert _cos(c1, c1_u8).min() >= COS_TOL, _cos(c1, c1_u8).min()
    assert _cos(c1_u8, ref1).min() >= COS_TOL, _cos(c1_u8, ref1).min()
    np.testing.assert_allclose(np.linalg.norm(c1, axis=1), 1.0, atol=1e-5)
    assert _cos(c1, ref1).min() >= COS_TOL, _cos(c1, ref1).min()
    assert _cos(c2, ref2).min() >= COS_TOL, _cos(c2, ref2).min()
    # prepared input + PREP_NONE handle == raw input + fused prepare
    c1_prep = net.encoder(1, _lib.PREP_NONE).embed_host(model.prepare(X1))
    assert _cos(c1_prep, c1).min() > 0.99999


def test_real_sheet_windows_shipped_weights():
    """Non-degenerate case from shipped data: windows of tutorials/sheet_image.png."""
    import cv2
    from audio_sheet_retrieval_b200 import _lib
    model, net = _net("mutopia_ccal_cont_rsz")
    onet = OracleNet("mutopia_ccal_cont_rsz", load_param_list(PKL["mutopia_ccal_cont_rsz"]))
    im = cv2.imread(os.path.join(GOLDEN, "sheet_image.png"), 0).astype(np.float32)
    wins = np.stack([im[y:y + 160, x:x + 200] for y in range(60, 1000, 97) for x in range(0, 600, 61)])[:, None]
    ref = onet.compute_view_1(wins)
    got = net.encoder(1, model.prepare.asr_prepare_mode).embed_host(wins)
    assert _cos(got, ref).min() >= COS_TOL
    # shifting a window by 4 px keeps it the nearest neighbour of its origin among all windows
    shifted = np.stack([im[y:y + 160, x + 4:x + 204] for y in range(60, 1000, 97) for x in range(0, 600, 61)])[:, None]
    got_s = net.encoder(1, model.prepare.asr_prepare_mode).embed_host(shifted)
    ref_s = onet.compute_view_1(shifted)
    assert ((got_s @ got.T).argmax(1) == (ref_s @ ref.T).argmax(1)).mean() >= 0.98


def test_embed_device_and_edge_sizes():
    from audio_sheet_retrieval_b200 import _lib
    model, net = _net("mutopia_ccal_cont_rsz", max_batch=8)
    X1, X2 = synth_inputs(9, seed=6)
    e2 = net.encoder(2, _lib.PREP_NONE)
    full = e2.embed_host(X2)
    one = e2.embed_host(X2[:1])
    assert (one == full[:1]).all()
    x = torch.as_tensor(X2[:8]).cuda()
    codes = torch.empty((8, 32), device="cuda")
    e2.embed_device(x, codes=codes)
    torch.cuda.synchronize()
    assert (codes.cpu().numpy() == full[:8]).all()
    with pytest.raises(Exception):
        e2.embed_device(torch.as_tensor(X2).cuda(), codes=torch.empty((9, 32), device="cuda"))   # n > max_batch
    with pytest.raises(ValueError):
        e2.embed_host(X1)                                                                         # wrong shape
    assert e2.embed_host(X2[:0]).shape == (0, 32)


def test_flip_filters_switch_changes_result():
    from audio_sheet_retrieval_b200 import _lib
    model, net = _net("mutopia_ccal_cont_rsz")
    X1, X2 = synth_inputs(4, seed=7)
    a = net.encoder(2, _lib.PREP_NONE).embed_host(X2)
    model2, net2 = _net("mutopia_ccal_cont_rsz")
    net2.flip_filters = True
    b = net2.encoder(2, _lib.PREP_NONE).embed_host(X2)
    onet = OracleNet("mutopia_ccal_cont_rsz", load_param_list(PKL["mutopia_ccal_cont_rsz"]), flip_filters=True)
    assert _cos(b, onet.compute_view_2(X2)).min() >= COS_TOL
    assert _cos(a, b).min() < 0.999


@pytest.mark.parametrize("model_name", ["mutopia_ccal_cont_rsz", "mutopia_ccal_cont"])
@pytest.mark.parametrize("dtype", ["u8", "f32"])
@pytest.mark.parametrize("mode", ["none", "scale", "scale_half"])
def test_layer0_tensor_core_all_input_forms(model_name, dtype, mode):
    """Layer 0 on the tensor cores (banded-Toeplitz GEMM) for every input dtype x prepare mode -- the
    integer-pixel form (u8, x/255), the generic hi/lo converter and the 2x2 box filter -- against the
    fp32 CUDA-core path on the same device: layer-0 activations to bf16 round-off, codes to COS_TOL."""
    from audio_sheet_retrieval_b200 import _lib
    model, net = _net(model_name, max_batch=16)
    prep = {"none": _lib.PREP_NONE, "scale": _lib.PREP_SCALE, "scale_half": _lib.PREP_SCALE_HALF}[mode]
    rng = np.random.RandomState(11)
    enc = net.encoder(1, prep)      # raw input size follows the mode (the box filter halves it)
    X = rng.randint(0, 256, size=(5, 1, enc.in_h, enc.in_w)).astype(np.uint8 if dtype == "u8" else np.float32)
    c_fused = enc.embed_host(X, path=_lib.PATH_TCGEN05) if enc.fusion else None
    enc.set_fusion(0)
    c_tc = enc.embed_host(X, path=_lib.PATH_TCGEN05)
    if c_fused is not None:         # the fused layer-0 + layer-1 kernel takes the same input forms
        assert _cos(c_fused, c_tc).min() > 0.9999
    a_tc = enc.debug_activation(0, 5, path=_lib.PATH_TCGEN05)
    c_fp = enc.embed_host(X, path=_lib.PATH_FP32)
    a_fp = enc.debug_activation(0, 5, path=_lib.PATH_FP32)
    scale = np.abs(a_fp).max()
    assert np.abs(a_tc - a_fp).max() <= 2.0 ** -7 * scale + 1e-3        # one bf16 ulp of the largest activation
    assert _cos(c_tc, c_fp).min() >= COS_TOL


@pytest.mark.parametrize("n", [1, 2, 3, 17])
def test_small_and_ragged_batches(n):
    """Batch sizes around the tile edges of the raster-over-samples layer-0 GEMM and the persistent conv
    grids (fewer items than SMs): every row must equal the row of a larger batch."""
    from audio_sheet_retrieval_b200 import _lib
    model, net = _net("mutopia_ccal_cont", max_batch=16)
    X1, X2 = synth_inputs(20, seed=9)
    e1 = net.encoder(1, model.prepare.asr_prepare_mode)
    e2 = net.encoder(2, _lib.PREP_NONE)
    full1, full2 = e1.embed_host(X1[:16].astype(np.uint8)), e2.embed_host(X2[:16])
    m = min(n, 16)
    got1, got2 = e1.embed_host(X1[:n].astype(np.uint8)), e2.embed_host(X2[:n])
    assert (got1[:m] == full1[:m]).all() and (got2[:m] == full2[:m]).all()


def test_fused_kernels_several_samples_per_cta():
    """More samples than SMs, not a multiple of the grid: the persistent fused kernels (layers 0 + 1, layers 2 + 3) walk
    several samples per CTA, so their ring / accumulator phases wrap many times; rows must equal the per-layer launches
    to bf16 round-off, and a sample's result must not depend on its position in the batch."""
    from audio_sheet_retrieval_b200 import _lib
    n = 2 * torch.cuda.get_device_properties(0).multi_processor_count + 37
    model, net = _net("mutopia_ccal_cont", max_batch=n)
    rng = np.random.RandomState(5)
    base = rng.randint(0, 256, size=(8, 1, 160, 200)).astype(np.uint8)
    base[base > 60] = 255                                   # sheet-like: mostly white
    X = base[rng.randint(0, 8, size=n)]                     # every sample is one of 8 images
    enc = net.encoder(1, model.prepare.asr_prepare_mode)
    if not (enc.fusion & 2):
        pytest.skip("layers 2 + 3 do not run fused on this geometry")
    fused = enc.embed_host(X, path=_lib.PATH_TCGEN05)
    enc.set_fusion(enc.fusion & 1)
    per_layer = enc.embed_host(X, path=_lib.PATH_TCGEN05)
    assert _cos(fused, per_layer).min() > 0.9999
    # identical inputs -> identical rows, wherever they sit in the batch
    first = {}
    for i in range(n):
        key = X[i].tobytes()
        if key in first:
            assert (fused[i] == fused[first[key]]).all(), (i, first[key])
        else:
            first[key] = i


_VARIANT_SNIPPET = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
import importlib
from oracle.encoders import synth_inputs
from audio_sheet_retrieval_b200 import _lib, network
from audio_sheet_retrieval_b200.params import load_params
out = {{}}
for name, pkl in {pkl!r}.items():
    model = importlib.import_module("audio_sheet_retrieval_b200.models." + name)
    layers = model.build_model(show_model=False)
    net = layers[0].net
    net.max_batch = 16
    network.set_all_param_values(layers, load_params(pkl))
    X1, X2 = synth_inputs(7, seed=21)
    out[name + "_1"] = net.encoder(1, model.prepare.asr_prepare_mode).embed_host(X1.astype(np.uint8))
    out[name + "_2"] = net.encoder(2, _lib.PREP_NONE).embed_host(X2)
np.savez({dst!r}, **out)
"""


@pytest.mark.parametrize("env", [{}, {"ASR_CONV_ROWS": "0"}, {"ASR_L0_TC": "0"}, {"ASR_CONV_ROWS_MULTI": "1"},
                                 {"ASR_CONV_ROWS": "2"}, {"ASR_FUSE01": "0"}, {"ASR_F01_VARIANT": "1"}, {"ASR_F01_VARIANT": "2"}, {"ASR_FUSE23": "0"}])
def test_kernel_variants_agree(env, tmp_path):
    """The kernel-selection switches are read once per process, so each variant runs in its own interpreter:
    raster-only conv, CUDA-core layer 0, side-by-side narrow tiles and row-stacked non-pooled layers must all
    reproduce the oracle to the same tolerance as the default selection."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = str(tmp_path / "codes.npz")
    e = dict(os.environ)
    e.update(env)
    subprocess.run([sys.executable, "-c", _VARIANT_SNIPPET.format(root=root, dst=dst, pkl=PKL)], check=True, env=e, cwd=root, timeout=600)
    got = np.load(dst)
    X1, X2 = synth_inputs(7, seed=21)
    for name in ("mutopia_ccal_cont", "mutopia_ccal_cont_rsz"):
        onet = OracleNet(name, load_param_list(PKL[name]))
        assert _cos(got[name + "_1"], onet.compute_view_1(X1)).min() >= COS_TOL, (env, name, 1)
        assert _cos(got[name + "_2"], onet.compute_view_2(X2)).min() >= COS_TOL, (env, name, 2)
