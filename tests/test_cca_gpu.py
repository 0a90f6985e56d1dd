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


@pytest.mark.parametrize("m,with_corr", [(64, False), (300, True), (2500, True)])
def test_layer_backward_matches_oracle(m, with_corr):
    """asr_cca_layer_backward (three Gram passes + one fp64 CTA + a row pass) vs the oracle's NumPy chain, which is pinned by
    central differences of the forward (tests/test_oracle_pins.py).  Inputs are float32 on both sides."""
    from audio_sheet_retrieval_b200.utils.cca import cca_layer_backward_device
    rng = np.random.RandomState(m)
    H1, H2 = occa.synth_latents(m, seed=m)
    G1 = rng.normal(size=(m, 32)).astype(np.float32)
    G2 = rng.normal(size=(m, 32)).astype(np.float32)
    gc = rng.normal(size=32) if with_corr else None
    # The forward is defined up to ONE sign per output column (shared by both views: whichever sign the eigensolver gives
    # F's columns; cca.py:174-175 only fixes U relative to V).  G1 / G2 are gradients w.r.t. the DEVICE's outputs, so the
    # oracle gets them in its own convention: column j times sign(<V_device[:, j], V_oracle[:, j]>).
    from audio_sheet_retrieval_b200.utils.cca import cca_sums_device, cca_solve_device
    import torch
    _, _, _, V_dev, _ = cca_solve_device(cca_sums_device(torch.as_tensor(H1).cuda(), torch.as_tensor(H2).cuda()), m, mode=1)
    sgn = np.sign((V_dev.cpu().numpy() * occa.cca_layer_train_forward(H1, H2)["V"]).sum(0))
    assert (sgn != 0).all()
    r1, r2 = occa.cca_layer_train_backward(H1, H2, G1 * sgn, G2 * sgn, g_corr=gc)
    d1, d2 = cca_layer_backward_device(H1, H2, G1, G2, g_corr=gc)
    d1, d2 = d1.cpu().numpy(), d2.cpu().numpy()
    for got, ref in ((d1, r1), (d2, r2)):
        assert np.isfinite(got).all()
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()      # fp32 outputs of an fp64 chain


def test_loss_and_layer_backward_chain():
    """contrastive loss gradient -> CCALayer backward: the gradient of the training objective w.r.t. the encoder latents,
    against central differences of (oracle forward -> oracle loss) on a few coordinates."""
    from audio_sheet_retrieval_b200.models.objectives import get_contrastive_cos_loss
    from audio_sheet_retrieval_b200.utils.cca import cca_layer_backward_device
    from oracle import objectives as oobj
    m = 100
    H1, H2 = occa.synth_latents(m, seed=11)
    H1, H2 = H1.astype(np.float64), H2.astype(np.float64)

    import torch

    def total(a, b):
        o = occa.cca_layer_train_forward(a, b)["out"]
        return oobj.contrastive_cos_loss_ref(o[:, :32], o[:, 32:], 1.0, 0.7, True)

    # forward on the device (its sign convention per output column is the one its backward differentiates)
    from audio_sheet_retrieval_b200.utils.cca import cca_sums_device, cca_solve_device
    m1, m2, U, V, _ = cca_solve_device(cca_sums_device(torch.as_tensor(H1.astype(np.float32)).cuda(),
                                                       torch.as_tensor(H2.astype(np.float32)).cuda()), m, mode=1)
    out = np.hstack(((H1 - m1.cpu().numpy()).dot(U.cpu().numpy()), (H2 - m2.cpu().numpy()).dot(V.cpu().numpy())))
    loss_fn = get_contrastive_cos_loss(1.0, 0.7, symmetric=True)
    val, g1, g2 = loss_fn.with_grads(torch.as_tensor(out[:, :32].astype(np.float32)).cuda(),
                                     torch.as_tensor(out[:, 32:].astype(np.float32)).cuda())
    assert abs(float(val) - total(H1, H2)) <= 1e-5 * abs(total(H1, H2))
    d1, d2 = cca_layer_backward_device(H1.astype(np.float32), H2.astype(np.float32), g1, g2)
    d1, d2 = d1.cpu().numpy(), d2.cpu().numpy()
    rng = np.random.RandomState(2)
    eps = 1e-5
    scale = max(np.abs(d1).max(), np.abs(d2).max())
    for _ in range(6):
        i, j, which = rng.randint(m), rng.randint(32), rng.randint(2)
        Hp, Hm = [H1.copy(), H2.copy()], [H1.copy(), H2.copy()]
        Hp[which][i, j] += eps
        Hm[which][i, j] -= eps
        fd = (total(*Hp) - total(*Hm)) / (2 * eps)
        an = (d1 if which == 0 else d2)[i, j]
        assert abs(fd - an) <= 2e-3 * scale, (which, i, j, fd, an)
