"""The only pin the reference itself offers for the encoder restatement (SURVEY.md 8c, row "What does pin the
restatement", item i): the shipped pickle is self-consistent with real sheet snippets.

The stored BatchNorm statistics (`mean`, `inv_std` = 1/sqrt(var + 1e-4)) of every layer were accumulated by the
reference while it trained on sheet images.  Pushing windows of the reference's own tutorials/sheet_image.png
(committed as tests/golden/sheet_image.png) through oracle/encoders.py must therefore reproduce those statistics
layer by layer -- which checks, in one go, the 97-array layout (group-of-5 order W, beta, gamma, mean, inv_std),
eps, input polarity and scale (/255, white = 1), the half-size `prepare`, the pooling positions, and -- at the
deeper layers, where second-order statistics are no longer flip invariant -- the convolution convention:
cross-correlation (cuDNN's Conv2DDNNLayer, what the weights were trained with) must fit strictly better than the
true convolution of the reference's CPU-fallback layer (audio_sheet_retrieval/models/mutopia_ccal_cont.py:12-18).

CPU only (torch fp32), ~10 s.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle.encoders import _elu, load_param_list, prepare_rsz, split_params

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _windows():
    import cv2
    im = cv2.imread(os.path.join(GOLDEN, "sheet_image.png"), 0).astype(np.float32)
    assert im.shape == (1181, 835)
    return np.stack([im[y:y + 160, x:x + 200] for y in range(0, 1181 - 160, 40) for x in range(0, 835 - 200, 50)])[:, None]


def _layer_stats(views, cca, X, flip):
    """Per layer: mean |log(measured std / stored std)| over live channels, mean offset in stored-std units."""
    h = torch.as_tensor(X)
    out = []
    for li, L in enumerate(views[0]):
        W = torch.as_tensor(L["W"])
        if flip:
            W = torch.flip(W, dims=(2, 3))
        y = F.conv2d(h, W, padding=W.shape[2] // 2)
        mu, sd = y.mean(dim=(0, 2, 3)).numpy(), y.std(dim=(0, 2, 3)).numpy()
        stored_sd = 1.0 / L["inv_std"]
        live = stored_sd > 0.02                     # dead channels (inv_std at its cap of 100) carry no information
        out.append((float(np.abs(np.log(sd[live] / stored_sd[live])).mean()),
                    float(np.abs(mu - L["mean"])[live].mean() / stored_sd[live].mean())))
        y = (y - torch.as_tensor(L["mean"]).view(1, -1, 1, 1)) * torch.as_tensor(L["gamma"] * L["inv_std"]).view(1, -1, 1, 1) \
            + torch.as_tensor(L["beta"]).view(1, -1, 1, 1)
        if li < 8:
            y = _elu(y)
            if li % 2 == 1:
                y = F.max_pool2d(y, 2)
        h = y
    z = (h.mean(dim=(2, 3)) - torch.as_tensor(cca["mean1"])) @ torch.as_tensor(cca["U"])
    return out, z.numpy()


def test_stored_bn_statistics_match_real_sheet_windows(shipped_params):
    views, cca = split_params(shipped_params)
    wins = _windows()
    X = prepare_rsz(wins)
    noflip, z_noflip = _layer_stats(views, cca, X, flip=False)
    flip, z_flip = _layer_stats(views, cca, X, flip=True)
    # (1) layout / eps / scale / polarity: every layer's activation std within ~5-17 % of the stored one
    for li, (dev, off) in enumerate(noflip):
        assert dev < (0.10 if li <= 5 else 0.20), "layer %d: mean |log std ratio| %.3f" % (li, dev)
        assert off < 0.15, "layer %d: mean offset %.3f stored std" % (li, off)
    # (2) convolution convention: from layer 5 on cross-correlation fits strictly better than true convolution
    for li in range(5, 9):
        assert noflip[li][0] < flip[li][0], "layer %d: no-flip %.4f vs flip %.4f" % (li, noflip[li][0], flip[li][0])
    assert np.mean([noflip[l][0] for l in range(5, 9)]) < 0.85 * np.mean([flip[l][0] for l in range(5, 9)])
    # ... and the CCA-projected output is closer to zero mean (the refit centred it on real data)
    rms = lambda z: float(np.sqrt((z.mean(0) ** 2).mean()))  # noqa: E731
    assert rms(z_noflip) < 0.85 * rms(z_flip), (rms(z_noflip), rms(z_flip))
    # (3) negative controls: the same check must FAIL for a wrong polarity and for a wrong group-of-5 order
    inv, _ = _layer_stats(views, cca, prepare_rsz(255.0 - wins), flip=False)
    assert inv[0][1] > 1.0, inv[0]                                   # black-on-white swapped: layer-0 means are far off
    swapped = [dict(L, mean=L["gamma"], gamma=L["mean"]) for L in views[0]]
    bad, _ = _layer_stats([swapped], cca, X, flip=False)
    assert max(b[1] for b in bad[:4]) > 1.0, bad[:4]


def test_cca_layer_backward_matches_central_differences():
    """The oracle's CCALayer training backward (the chain Theano's autodiff produces for layers/cca.py:91-203, written out)
    against central differences of the oracle's forward: a weighted sum of both outputs plus a term in corr."""
    from oracle import cca as occa
    rng = np.random.RandomState(0)
    m, d = 72, 32
    H1, H2 = (a.astype(np.float64) for a in occa.synth_latents(m, seed=3))
    W1, W2, gc = rng.normal(size=(m, d)), rng.normal(size=(m, d)), rng.normal(size=d)

    def loss(a, b):
        o = occa.cca_layer_train_forward(a, b)
        return (o["out"][:, :d] * W1).sum() + (o["out"][:, d:] * W2).sum() + (o["corr"] * gc).sum()

    g1, g2 = occa.cca_layer_train_backward(H1, H2, W1, W2, g_corr=gc)
    assert abs(g1.sum(0)).max() < 1e-9 * abs(g1).max() and abs(g2.sum(0)).max() < 1e-9 * abs(g2).max()   # shift invariance
    eps = 1e-6
    for _ in range(10):
        i, j, which = rng.randint(m), rng.randint(d), rng.randint(2)
        Hp, Hm = [H1.copy(), H2.copy()], [H1.copy(), H2.copy()]
        Hp[which][i, j] += eps
        Hm[which][i, j] -= eps
        fd = (loss(*Hp) - loss(*Hm)) / (2 * eps)
        an = (g1 if which == 0 else g2)[i, j]
        assert abs(fd - an) <= 2e-5 * max(1.0, abs(fd)), (which, i, j, fd, an)
