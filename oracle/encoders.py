"""Oracle: the two encoder branches + CCA projection + length norm (CPU, torch).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED for this
file: Theano/Lasagne cannot run here, the restatement follows documented
Lasagne semantics.

Follows
  asr/models/mutopia_ccal_cont_rsz.py:54-58   conv_bn = 3x3 pad-1 conv (no bias) -> BN -> ELU
  asr/models/mutopia_ccal_cont_rsz.py:64-145  layer graph (24,24,P,48,48,P,96,96,P,96,96,P, 1x1->32, BN, global mean)
  asr/models/mutopia_ccal_cont.py:64-145      same graph with 12/24/48/48 filters, full-res sheet input
  asr/models/mutopia_ccal_cont_rsz.py:170-190 prepare (/255, cv2.resize to half)
  asr/models/mutopia_ccal_cont.py:170-190     prepare (/255 only)
  asr/models/lasagne_extensions/layers/cca.py:184-203  CCALayer deterministic branch
  asr/models/lasagne_extensions/layers/cca.py:39-40    LengthNormLayer

Convolution semantics: the shipped weights were trained with cuDNN's
Conv2DDNNLayer (flip_filters=False => cross-correlation).  `flip_filters=True`
reproduces the reference's CPU fallback layer (true convolution) and is kept
only as a switch (hard part 4).
"""
import pickle

import numpy as np
import torch
import torch.nn.functional as F

MODEL_FILTERS = {
    "mutopia_ccal_cont_rsz": (24, 24, 48, 48, 96, 96, 96, 96),
    "mutopia_ccal_cont": (12, 12, 24, 24, 48, 48, 48, 48),
}
DIM_LATENT = 32
N_CONV = 9  # 8 3x3 layers + 1x1 head per view
BN_EPS = 1e-4


def load_param_list(path):
    """Flat list of 97 float32 arrays (asr/retrieval_wrapper.py:27-29; py2 pickle)."""
    with open(path, "rb") as fp:
        params = pickle.load(fp, encoding="latin1")
    if isinstance(params[0], list):  # "old redundant dump" asr/run_eval.py:76-80
        raise ValueError("redundant per-layer dumps are not handled by the oracle")
    return [np.asarray(p) for p in params]


def split_params(params):
    """Decode the 97-array layout (SURVEY.md 8b face 3)."""
    assert len(params) == 97, len(params)
    views = []
    for v in range(2):
        layers = []
        for l in range(N_CONV):
            W, beta, gamma, mean, inv_std = params[v * 45 + l * 5: v * 45 + l * 5 + 5]
            layers.append(dict(W=W, beta=beta, gamma=gamma, mean=mean, inv_std=inv_std))
        views.append(layers)
    cca = dict(U=params[90], V=params[91], mean1=params[92], mean2=params[93],
               S12=params[94], S11=params[95], S22=params[96])
    return views, cca


def prepare_rsz(x):
    """asr/models/mutopia_ccal_cont_rsz.py:170-190 (x only)."""
    import cv2
    x = x.astype(np.float32)
    x /= 255
    sheet_shape = [x.shape[2] // 2, x.shape[3] // 2]
    x_new = np.zeros([x.shape[0], x.shape[1]] + sheet_shape, np.float32)
    for i in range(len(x)):
        x_new[i, 0] = cv2.resize(x[i, 0], (sheet_shape[1], sheet_shape[0]))
    return x_new


def prepare_full(x):
    """asr/models/mutopia_ccal_cont.py:170-190 (x only)."""
    x = x.astype(np.float32)
    x /= 255
    return x


PREPARE = {"mutopia_ccal_cont_rsz": prepare_rsz, "mutopia_ccal_cont": prepare_full}


def _elu(x):
    # lasagne.nonlinearities.elu: x if x > 0 else expm1(x)
    return torch.where(x > 0, x, torch.expm1(x))


def encoder_latent(x, layers, dtype=torch.float32, flip_filters=False, quant=None):
    """One branch up to the pre-CCA latent (N,32).

    x: (N,1,H,W) already prepared.  `quant` optionally rounds activations and
    weights through a low-precision dtype (torch.bfloat16 / torch.float16)
    between layers, modelling a tensor-core path with fp32 accumulation; used
    only to study precision, never as "truth".
    """
    h = torch.as_tensor(np.ascontiguousarray(x)).to(dtype)

    def q(t):
        return t.to(quant).to(dtype) if quant is not None else t

    for li, L in enumerate(layers):
        W = torch.as_tensor(L["W"]).to(dtype)
        if flip_filters:
            W = torch.flip(W, dims=(2, 3))
        scale = torch.as_tensor(L["gamma"]).to(dtype) * torch.as_tensor(L["inv_std"]).to(dtype)
        mean = torch.as_tensor(L["mean"]).to(dtype)
        beta = torch.as_tensor(L["beta"]).to(dtype)
        last = li == len(layers) - 1
        if quant is not None and 0 < li < len(layers) - 1:
            # tensor-core path: BN scale folded into bf16 weights, bf16 activations
            Wf = q(W * scale.view(-1, 1, 1, 1))
            y = F.conv2d(q(h), Wf, padding=W.shape[2] // 2)
            y = y + (beta - mean * scale).view(1, -1, 1, 1)
        else:
            y = F.conv2d(h, W, padding=W.shape[2] // 2)
            y = (y - mean.view(1, -1, 1, 1)) * scale.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
        if not last:
            y = _elu(y)
            if li % 2 == 1:
                y = F.max_pool2d(y, 2)  # stride 2, ignore_border=True (floor)
        h = y
    return h.mean(dim=(2, 3))  # GlobalPoolLayer + Flatten


def cca_project(lat, mean, U, dtype=torch.float32):
    """CCALayer deterministic: (H - mean) . U   (cca.py:188-201)."""
    lat = lat.to(dtype)
    return (lat - torch.as_tensor(mean).to(dtype)) @ torch.as_tensor(U).to(dtype)


def length_norm(z):
    """LengthNormLayer (cca.py:39-40): x / ||x||_2 per row, no epsilon."""
    return z / z.norm(2, dim=1, keepdim=True)


class OracleNet(object):
    """Both branches with the reference's eval-mode semantics."""

    def __init__(self, model_name, params, flip_filters=False):
        self.model_name = model_name
        self.params = params
        self.views, self.cca = split_params(params)
        self.flip_filters = flip_filters
        filt = MODEL_FILTERS[model_name]
        for v in range(2):
            for l in range(8):
                assert self.views[v][l]["W"].shape[0] == filt[l], (v, l, self.views[v][l]["W"].shape)

    def prepare(self, x):
        return PREPARE[self.model_name](x)

    def latent(self, view, x, dtype=torch.float32, quant=None, batch=100):
        """Pre-CCA latent for prepared input x; batched like batch_compute1 (batch_iterators.py:17-62)."""
        outs = []
        with torch.no_grad():
            for s in range(0, x.shape[0], batch):
                outs.append(encoder_latent(x[s:s + batch], self.views[view - 1], dtype,
                                           self.flip_filters, quant))
        return torch.cat(outs, 0)

    def code(self, view, x, dtype=torch.float32, quant=None, batch=100):
        """Final 32-d unit-norm code for prepared input x."""
        lat = self.latent(view, x, dtype, quant, batch)
        if view == 1:
            z = cca_project(lat, self.cca["mean1"], self.cca["U"], dtype)
        else:
            z = cca_project(lat, self.cca["mean2"], self.cca["V"], dtype)
        return length_norm(z)

    def compute_view_1(self, X, dtype=torch.float32):
        """RetrievalWrapper.compute_view_1 with prepare_view_1=model.prepare (retrieval_wrapper.py:47-61)."""
        return self.code(1, self.prepare(X), dtype).numpy()

    def compute_view_2(self, Z, dtype=torch.float32):
        """RetrievalWrapper.compute_view_2 with prepare_view_2=None (retrieval_wrapper.py:63-77)."""
        return self.code(2, np.asarray(Z, np.float32), dtype).numpy()


# ----------------------------------------------------------------------------
# synthetic parameters for the full-resolution model (no pickle is shipped for it)
# ----------------------------------------------------------------------------
def synth_inputs(n, seed=23, sheet_hw=(160, 200), spec_hw=(92, 42), sheet_dtype=np.float32):
    """Synthetic sheet snippets (0..255, mostly white, dark staff lines + blobs) and
    log-spectrogram-like excerpts (>=0, sparse ridges, mean ~0.11).  SURVEY.md 8d config 1."""
    rng = np.random.RandomState(seed)
    H, W = sheet_hw
    X1 = np.full((n, 1, H, W), 255.0, np.float32)
    for i in range(n):
        y0 = rng.randint(20, 50)
        gap = rng.randint(6, 10)
        for s in range(2):
            for l in range(5):
                y = y0 + s * 70 + l * gap
                if y < H:
                    X1[i, 0, y:y + 1, :] = rng.randint(0, 60)
        nb = rng.randint(10, 40)
        ys = rng.randint(0, H - 8, nb)
        xs = rng.randint(0, W - 8, nb)
        for y, x in zip(ys, xs):
            h, w = rng.randint(3, 8), rng.randint(3, 8)
            X1[i, 0, y:y + h, x:x + w] = rng.randint(0, 100)
        nv = rng.randint(4, 16)
        for x in rng.randint(0, W - 1, nv):
            ya = rng.randint(0, H - 30)
            X1[i, 0, ya:ya + rng.randint(10, 30), x:x + 1] = rng.randint(0, 80)
    Hs, Ws = spec_hw
    X2 = np.zeros((n, 1, Hs, Ws), np.float32)
    for i in range(n):
        nn = rng.randint(3, 10)
        for _ in range(nn):
            f0 = rng.randint(5, 40)
            t0 = rng.randint(0, Ws - 4)
            dur = rng.randint(4, Ws)
            amp = rng.uniform(0.5, 2.0)
            t = np.arange(Ws)
            env = np.where(t >= t0, np.exp(-(t - t0) / float(dur)), 0.0) * (t < t0 + 2 * dur)
            for hnum in range(1, 6):
                f = int(f0 + 12 * np.log2(hnum))
                if f < Hs:
                    X2[i, 0, f, :] += amp / hnum * env
        X2[i, 0] += np.abs(rng.normal(0, 0.02, (Hs, Ws)))
    X2 = X2.astype(np.float32)
    if sheet_dtype != np.float32:
        X1 = X1.astype(sheet_dtype)
    return X1, X2


def synth_params(model_name, seed=23, calib_n=16):
    """97-array list for `model_name` with He-uniform weights and BN statistics
    calibrated on a synthetic batch so activations stay in range; CCA = svd refit
    on the calibration batch scaled like the shipped U/V (hard part 9)."""
    rng = np.random.RandomState(seed)
    filt = MODEL_FILTERS[model_name]
    X1, X2 = synth_inputs(calib_n, seed=seed + 1)
    X1 = PREPARE[model_name](X1)
    params = []
    lats = []
    for v, x in enumerate((X1, X2)):
        h = torch.as_tensor(x)
        cin = 1
        for li in range(N_CONV):
            cout = filt[li] if li < 8 else DIM_LATENT
            ks = 3 if li < 8 else 1
            fan_in = cin * ks * ks
            bound = np.sqrt(6.0 / fan_in)  # lasagne.init.HeUniform, gain 1
            W = rng.uniform(-bound, bound, (cout, cin, ks, ks)).astype(np.float32)
            with torch.no_grad():
                y = F.conv2d(h, torch.as_tensor(W), padding=ks // 2)
            mean = y.mean(dim=(0, 2, 3)).numpy().astype(np.float32)
            var = y.var(dim=(0, 2, 3), unbiased=False).numpy().astype(np.float32)
            inv_std = (1.0 / np.sqrt(var + BN_EPS)).astype(np.float32)
            gamma = rng.uniform(0.8, 1.2, cout).astype(np.float32)
            beta = rng.uniform(-0.2, 0.2, cout).astype(np.float32)
            if li == 8:  # shipped head: tiny latent variance (S11 diag ~4e-3)
                gamma = (gamma * 0.06).astype(np.float32)
                beta = (beta * 0.05).astype(np.float32)
            params += [W, beta, gamma, mean, inv_std]
            with torch.no_grad():
                y = (y - torch.as_tensor(mean).view(1, -1, 1, 1)) * torch.as_tensor(gamma * inv_std).view(1, -1, 1, 1) \
                    + torch.as_tensor(beta).view(1, -1, 1, 1)
                if li < 8:
                    y = _elu(y)
                    if li % 2 == 1:
                        y = F.max_pool2d(y, 2)
            h = y
            cin = cout
        lats.append(h.mean(dim=(2, 3)).numpy())
    # CCA part: random well-conditioned projections of the right magnitude
    m1 = lats[0].mean(0).astype(np.float32)
    m2 = lats[1].mean(0).astype(np.float32)

    def proj():
        Q, _ = np.linalg.qr(rng.normal(size=(32, 32)))
        s = np.linspace(10.0, 6.0, 32)
        return (Q * s).astype(np.float32)

    U, V = proj(), proj()
    S = (np.eye(32) * 4e-3).astype(np.float32)
    params += [U, V, m1, m2, (S * 0.5).astype(np.float32), S.copy(), S.copy()]
    assert len(params) == 97
    return params
