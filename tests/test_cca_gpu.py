"""CCA statistics + Jacobi solve vs the oracle / the reference's golden fit."""
import numpy as np
import pytest

from oracle import cca as occa

pytestmark = pytest.mark.gpu


def _cos_cols(A, B):
    return np.abs((A * B).sum(0)) / np.linalg.norm(A, axis=0) / np.linalg.norm(B, axis=0)


def test_fit_matches_reference_golden(golden):
    from audio_sheet_retrieval_b200.utils.cca import CCA
    c = CCA(method="svd")
    sig = c.fit(golden["cca_H1"], golden["cca_H2"])
    np.testing.assert_allclose(sig, golden["cca_sigma"], atol=1e-5)      # tolerance: SURVEY 8d config 3
    np.testing.assert_allclose(c.m1, golden["cca_m1"], atol=1e-7)
    np.testing.assert_allclose(c.m2, golden["cca_m2"], atol=1e-7)
    # projections equal up to a joint sign per component
    assert _cos_cols(c.U, golden["cca_U"]).min() > 0.9999
    assert _cos_cols(c.V, golden["cca_V"]).min() > 0.9999
    sgn = np.sign((c.U * golden["cca_U"]).sum(0))
    np.testing.assert_allclose(np.sign((c.V * golden["cca_V"]).sum(0)), sgn)
    T1 = c.transform_V1(golden["cca_H1"][:16]) * sgn
    np.testing.assert_allclose(T1, golden["cca_T1"], atol=2e-3, rtol=2e-3)


def test_fit_invariants_25000():
    from audio_sheet_retrieval_b200.utils.cca import CCA
    H1, H2 = occa.synth_latents(25000, seed=23)
    c = CCA(method="svd")
    sig = c.fit(H1, H2)
    o = occa.CCA()
    sig_ref = o.fit(H1.astype(np.float64), H2.astype(np.float64))
    I = np.eye(32)
    np.testing.assert_allclose(c.U.T @ o.S11 @ c.U, I, atol=1e-8)
    np.testing.assert_allclose(c.V.T @ o.S22 @ c.V, I, atol=1e-8)
    np.testing.assert_allclose(c.U.T @ o.S12 @ c.V, np.diag(sig), atol=1e-8)
    np.testing.assert_allclose(sig, sig_ref, atol=1e-9)
    assert (np.diff(sig) <= 0).all()


def test_sharded_sums_equal_single_pass():
    import torch
    from audio_sheet_retrieval_b200.utils.cca import cca_sums_device, cca_solve_device
    H1, H2 = occa.synth_latents(5003, seed=5)
    h1, h2 = torch.as_tensor(H1).cuda(), torch.as_tensor(H2).cuda()
    full = cca_sums_device(h1, h2)
    part = None
    for r in range(4):
        lo, hi = 5003 * r // 4, 5003 * (r + 1) // 4
        part = cca_sums_device(h1[lo:hi].contiguous(), h2[lo:hi].contiguous(), sums=part)
    np.testing.assert_allclose(part.cpu().numpy(), full.cpu().numpy(), rtol=1e-12, atol=1e-14)
    ref = np.concatenate([H1.astype(np.float64).sum(0), H2.astype(np.float64).sum(0),
                          (H1.astype(np.float64).T @ H1.astype(np.float64)).ravel(),
                          (H2.astype(np.float64).T @ H2.astype(np.float64)).ravel(),
                          (H1.astype(np.float64).T @ H2.astype(np.float64)).ravel()])
    np.testing.assert_allclose(full.cpu().numpy(), ref, rtol=1e-11, atol=1e-13)
    m1, m2, U, V, sig = cca_solve_device(full, 5003)
    o = occa.CCA()
    sig_ref = o.fit(H1.astype(np.float64), H2.astype(np.float64))
    np.testing.assert_allclose(sig.cpu().numpy(), sig_ref, atol=1e-8)


def test_layer_train_forward_mode():
    """CCALayer non-deterministic forward (layers/cca.py:91-182): invariants of the eigh form."""
    import torch
    from audio_sheet_retrieval_b200.utils.cca import cca_sums_device, cca_solve_device
    H1, H2 = occa.synth_latents(100, seed=9)          # BATCH_SIZE = 100
    ref = occa.cca_layer_train_forward(H1, H2)
    h1, h2 = torch.as_tensor(H1).cuda(), torch.as_tensor(H2).cuda()
    sums = cca_sums_device(h1, h2)
    m1, m2, U, V, corr = cca_solve_device(sums, 100, mode=1)
    U, V, corr = U.cpu().numpy(), V.cpu().numpy(), corr.cpu().numpy()
    np.testing.assert_allclose(m1.cpu().numpy(), ref["mean1"], atol=1e-9)
    np.testing.assert_allclose(corr, ref["corr"], atol=1e-7)
    d = np.diag(U.T @ ref["S12"] @ V)
    assert (d >= -1e-10).all()                                      # sign fix
    np.testing.assert_allclose(U.T @ ref["S11"] @ U, np.eye(32), atol=1e-7)
    np.testing.assert_allclose(V.T @ ref["S22"] @ V, np.eye(32), atol=1e-7)
    # projected batch equals the oracle's up to a joint sign per component
    out = np.hstack(((H1 - ref["mean1"]) @ U, (H2 - ref["mean2"]) @ V))
    well = np.where(np.diff(ref["corr"], prepend=-1) > 1e-4)[0]   # skip (near-)degenerate eigenvalues
    for j in well:
        s = np.sign((out[:, j] * ref["out"][:, j]).sum())
        np.testing.assert_allclose(out[:, j] * s, ref["out"][:, j], atol=1e-5)
        np.testing.assert_allclose(out[:, 32 + j] * s, ref["out"][:, 32 + j], atol=1e-5)


def test_counted_layout_sharded_equals_fit():
    """The one-buffer layout of a row-sharded fit: four shards accumulate [sums | row count] into 3137 doubles (what
    the single all-reduce carries), the solve reads the count from the device -- equal to CCA.fit on all rows."""
    import torch
    from audio_sheet_retrieval_b200 import _lib
    from audio_sheet_retrieval_b200.utils.cca import CCA, cca_solve_device
    H1, H2 = occa.synth_latents(5003, seed=6)
    h1, h2 = torch.as_tensor(H1).cuda(), torch.as_tensor(H2).cuda()
    total = torch.zeros(_lib.CCA_NSUMS + 1, dtype=torch.float64, device="cuda")
    for r in range(4):
        lo, hi = 5003 * r // 4, 5003 * (r + 1) // 4
        part = torch.zeros_like(total)
        _lib.check(_lib.lib.asr_cca_accumulate_counted(_lib.dptr(h1[lo:hi].contiguous()), _lib.dptr(h2[lo:hi].contiguous()),
                                                       hi - lo, None, None, _lib.dptr(part), _lib.stream_ptr()))
        assert part[-1].item() == hi - lo
        total += part                                          # stands for the all-reduce
    m1, m2, U, V, sig = cca_solve_device(total, _lib.CCA_COUNT_ON_DEVICE)
    c = CCA()
    sig_ref = c.fit(H1, H2)
    np.testing.assert_allclose(sig.cpu().numpy(), sig_ref, atol=1e-12)
    np.testing.assert_allclose(U.cpu().numpy(), c.U, atol=1e-9 * np.abs(c.U).max())
    np.testing.assert_allclose(m1.cpu().numpy(), H1.astype(np.float64).mean(0), atol=1e-12)
