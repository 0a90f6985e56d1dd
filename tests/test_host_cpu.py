"""CPU-side checks: the C-ABI library loads and exports what include/asr_b200.h declares, compute
calls fail loudly without a GPU, host logic (pickle format, batching, model protocol, data pool)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NO_GPU = not torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    from audio_sheet_retrieval_b200 import _lib
    header = open(os.path.join(ROOT, "include", "asr_b200.h")).read()
    declared = set(re.findall(r"\b(asr_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert _lib.lib.asr_abi_version() == _lib.ABI_VERSION == 2


def test_library_is_sm100a_with_tcgen05_and_tma():
    """The shipped binary really contains Blackwell tensor-core / TMA code."""
    import subprocess
    from audio_sheet_retrieval_b200 import _lib
    try:
        sass = subprocess.run(["cuobjdump", "-sass", _lib.SO_PATH], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass, "no tcgen05.mma in SASS"
    assert "LDTM" in sass, "no tcgen05.ld in SASS"
    assert "UBLKCP" in sass and "UTMALDG" in sass, "no TMA bulk / tensor loads in SASS"
    # every hot kernel of the path is in the binary
    for kernel in ("l0_tc_kernel", "l01_fused_kernel", "l23_fused_kernel", "conv3x3_rows_kernel", "conv3x3_tc_kernel", "head_kernel", "topk_stream_kernel",
                   "topk_tc_kernel", "topk_merge_select_kernel", "cca_solve_kernel", "contrastive_rows_kernel"):
        assert kernel in sass, kernel


@pytest.mark.skipif(not NO_GPU, reason="checks the no-GPU failure mode")
def test_compute_calls_fail_loudly_without_gpu(shipped_params):
    from audio_sheet_retrieval_b200 import _lib, network
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont_rsz as model
    from audio_sheet_retrieval_b200.retrieval import EmbeddingDB
    from audio_sheet_retrieval_b200.utils.cca import CCA
    layers = model.build_model(show_model=False)
    network.set_all_param_values(layers, shipped_params)
    with pytest.raises(_lib.AsrError, match="no CUDA device"):
        layers[0].net.encoder(1)
    with pytest.raises(_lib.AsrError, match="no CPU fallback"):
        EmbeddingDB(np.zeros((4, 32), np.float32))
    with pytest.raises(_lib.AsrError):
        CCA().fit(np.zeros((4, 32), np.float32), np.zeros((4, 32), np.float32))
    h = ctypes.c_void_p()
    assert _lib.lib.asr_db_create(ctypes.byref(h), None, 4, 0) != 0
    assert b"no CUDA device" in _lib.lib.asr_last_error()
    # training objective: CPU tensors are refused by the mirror, the C entry itself fails without a device
    import torch
    from audio_sheet_retrieval_b200.models.objectives import get_contrastive_cos_loss
    with pytest.raises(_lib.AsrError, match="no CPU path"):
        get_contrastive_cos_loss(1.0, 0.7)(torch.zeros(4, 32), torch.zeros(4, 32))
    assert _lib.lib.asr_contrastive_loss(None, None, 4, 1.0, 0.7, 0, None, None, None, None, None) != 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "audio_sheet_retrieval_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/_" not in src.replace("oracle/search.py", ""), f


def test_pickle_roundtrip_is_py2_numpy_compatible(tmp_path, shipped_params):
    import pickletools
    from audio_sheet_retrieval_b200 import params as P
    out = tmp_path / "p.pkl"
    P.save_params(str(out), shipped_params)
    raw = open(out, "rb").read()
    assert raw[:2] == b"\x80\x02"                       # protocol 2, as `protocol=-1` under Python 2
    globs = set(arg for op, arg, _ in pickletools.genops(raw) if op.name == "GLOBAL")
    assert "numpy.core.multiarray _reconstruct" in globs and not any("_core" in g for g in globs)
    back = P.load_params(str(out))
    assert len(back) == 97 and all((a == b).all() and b.dtype == np.float32 for a, b in zip(shipped_params, back))
    views, cca = P.split_params(back)
    assert views[1][8]["W"].shape == (32, 96, 1, 1) and cca["S22"].shape == (32, 32)
    shapes = P.expected_shapes((24, 24, 48, 48, 96, 96, 96, 96))
    assert [tuple(p.shape) for p in back] == [tuple(s) for s in shapes]


def test_set_params_rejects_wrong_model(shipped_params):
    from audio_sheet_retrieval_b200 import network
    from audio_sheet_retrieval_b200.models import mutopia_ccal_cont as full
    layers = full.build_model(show_model=False)
    with pytest.raises(ValueError, match="mismatch"):
        network.set_all_param_values(layers, shipped_params)      # rsz weights into the 12-filter model
    with pytest.raises(ValueError, match="mismatch"):
        network.set_all_param_values(layers, shipped_params[:90])


def test_model_module_protocol():
    from audio_sheet_retrieval_b200.run_train import compile_tag, select_model
    model, _ = select_model("models/mutopia_ccal_cont_rsz.py")
    assert model.EXP_NAME == "mutopia_ccal_cont_rsz" and model.DIM_LATENT == 32 and model.BATCH_SIZE == 100
    l1, l2, a, b = model.build_model(show_model=False)
    assert l1.output_shape == (None, 1, 80, 100) and l2.output_shape == (None, 1, 92, 42) and a.output_shape == (None, 32)
    full, _ = select_model("mutopia_ccal_cont")
    assert full.build_model(False)[0].output_shape == (None, 1, 160, 200)
    assert compile_tag("splits/all_split.yaml", "exp_configs/mutopia_full_aug.yaml") == "all_split_mutopia_full_aug"
    x = np.arange(2 * 160 * 200, dtype=np.float32).reshape(2, 1, 160, 200) % 256
    assert model.prepare(x).shape == (2, 1, 80, 100) and full.prepare(x).shape == (2, 1, 160, 200)
    assert 0.5 < model.prepare(x).max() <= 1.0 and full.prepare(x).max() == 1.0


def test_batch_compute_generic_path_matches_reference(golden):
    from audio_sheet_retrieval_b200.utils.batch_iterators import batch_compute1, batch_compute2
    Xa, Xb = golden["bc_Xa"], golden["bc_Xb"]
    calls = []

    def f1(E):
        calls.append(E.shape[0])
        return E.reshape(E.shape[0], -1)[:, :3] * 2.0

    def f2(E1, E2):
        return np.hstack((E1.reshape(E1.shape[0], -1)[:, :2], E2.reshape(E2.shape[0], -1)[:, :2]))

    assert (batch_compute1(Xa, f1, 10, prepare=lambda e: e + 1.0) == golden["bc1"]).all()
    assert calls == list(golden["bc1_calls"])                        # 3 full-size batches (last one zero padded)
    assert (batch_compute2(Xa, Xb, f2, 10, prepare1=lambda e: e * 0.5) == golden["bc2"]).all()
    with pytest.raises(TypeError):                                   # the reference's latent prepare2 bug (:98-99)
        batch_compute2(Xa, Xb, f2, 10, prepare1=None, prepare2=lambda e: e)


def test_synthetic_pool_protocol():
    from audio_sheet_retrieval_b200.run_train import select_data
    with pytest.raises(RuntimeError, match="msmd"):                   # the real data set is never faked
        select_data("mutopia", None, None, seed=23, test_only=True)
    with pytest.raises(ValueError):
        select_data("imagenet", None, None)
    data = select_data("synthetic", None, None, seed=23, test_only=True)
    pool = data["test"]
    assert pool.shape[0] == 2000 and data["train"] is None
    X1, X2 = pool[np.array([0, 7, 1999])]
    assert X1.shape == (3, 1, 160, 200) and X2.shape == (3, 1, 92, 42) and X1.dtype == np.float32
    assert X1.max() == 255.0 and X1.min() >= 0 and X2.min() >= 0
    Y1, _ = pool[7:8]
    assert (Y1[0] == X1[1]).all()                                    # deterministic per index
    assert 0.7 < X1.mean() / 255 < 0.99
