"""Contrastive cosine loss (first slice) through the C ABI vs the oracle and the
reference's golden values.  Tolerance: fp32 dot products and sums -> 2e-5 relative on the loss,
1e-5 absolute on the gradients (they are O(weight / n))."""
import os

import numpy as np
import pytest
import torch

from oracle.objectives import contrastive_cos_grads_ref, contrastive_cos_loss_ref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_contrastive_loss.npz")


def _cases():
    d = np.load(GOLD)
    for ci in range(4):
        for weight, gamma, sym in ((1.0, 0.7, False), (1.0, 0.7, True), (0.35, 0.2, True)):
            yield d["c%d_lv1" % ci], d["c%d_lv2" % ci], weight, gamma, sym, float(d["c%d_w%g_g%g_s%d" % (ci, weight, gamma, int(sym))])


def test_loss_matches_reference_golden_and_oracle_gradients():
    from audio_sheet_retrieval_b200.models.objectives import get_contrastive_cos_loss
    for lv1, lv2, weight, gamma, sym, want in _cases():
        f = get_contrastive_cos_loss(weight, gamma, symmetric=sym)
        a, b = torch.as_tensor(lv1).cuda(), torch.as_tensor(lv2).cuda()
        loss = float(f(a, b))
        assert abs(loss - want) <= 2e-5 * max(abs(want), 1e-3), (weight, gamma, sym, loss, want)
        loss2, g1, g2 = f.with_grads(a, b)
        assert float(loss2) == loss
        r1, r2 = contrastive_cos_grads_ref(lv1, lv2, weight, gamma, sym)
        np.testing.assert_allclose(g1.cpu().numpy(), r1, atol=1e-5, rtol=1e-4)
        np.testing.assert_allclose(g2.cpu().numpy(), r2, atol=1e-5, rtol=1e-4)


def test_loss_is_deterministic_and_scales_with_batch():
    from audio_sheet_retrieval_b200.models.objectives import contrastive_cos_loss
    rng = np.random.RandomState(0)
    for n in (2, 3, 129, 1000, 4096):
        a = rng.randn(n, 32).astype(np.float32)
        b = (a + 0.5 * rng.randn(n, 32)).astype(np.float32)
        a /= np.linalg.norm(a, axis=1, keepdims=True)
        b /= np.linalg.norm(b, axis=1, keepdims=True)
        ta, tb = torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda()
        l1, g1, h1 = contrastive_cos_loss(ta, tb, 1.0, 0.7, True, want_grads=True)
        l2, g2, h2 = contrastive_cos_loss(ta, tb, 1.0, 0.7, True, want_grads=True)
        assert float(l1) == float(l2) and torch.equal(g1, g2) and torch.equal(h1, h2)
        want = contrastive_cos_loss_ref(a, b, 1.0, 0.7, True)
        assert abs(float(l1) - want) <= 2e-5 * max(abs(want), 1e-3)


def test_bad_arguments_fail_loudly():
    from audio_sheet_retrieval_b200 import _lib
    from audio_sheet_retrieval_b200.models.objectives import contrastive_cos_loss
    a = torch.zeros(1, 32, device="cuda")
    with pytest.raises(_lib.AsrError):
        contrastive_cos_loss(a, a, 1.0, 0.7)
    with pytest.raises(ValueError):
        contrastive_cos_loss(torch.zeros(4, 16, device="cuda"), torch.zeros(4, 16, device="cuda"), 1.0, 0.7)
    with pytest.raises(_lib.AsrError):
        contrastive_cos_loss(torch.zeros(4, 32), torch.zeros(4, 32), 1.0, 0.7)
