"""Oracle: CCA refit ('svd') and CCALayer train-statistics forward (NumPy, CPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  asr/utils/cca.py:25-53      CCA.fit common part (means, fp32 covariances, +r*I)
  asr/utils/cca.py:199-211    method == 'svd'
  asr/utils/cca.py:432-444    transform_V1 / transform_V2
  asr/refine_cca.py:100-107   how the fit is written back (float32 casts)
  asr/models/lasagne_extensions/layers/cca.py:91-182  CCALayer non-deterministic forward
"""
import numpy as np
from scipy.linalg import sqrtm


class CCA(object):
    """asr/utils/cca.py:6-23 (constructor), method 'svd' only."""

    def __init__(self, r1=1e-3, r2=1e-3, rT=1e-3, method="svd"):
        assert method == "svd", "only method='svd' is ever selected (refine_cca.py:100)"
        self.r1, self.r2, self.rT, self.method = r1, r2, rT, method
        self.m1 = self.m2 = self.U = self.V = None

    def fit(self, H1, H2, verbose=False):
        m = H1.shape[0]                                   # cca.py:29
        self.m1 = np.mean(H1, axis=0)                     # cca.py:32-33
        self.m2 = np.mean(H2, axis=0)
        H1bar = (H1 - self.m1).T                          # cca.py:36-41
        H2bar = (H2 - self.m2).T
        S12 = (1.0 / (m - 1)) * np.dot(H1bar, H2bar.T)    # cca.py:44 (fp32 sgemm when inputs are fp32)
        S11 = (1.0 / (m - 1)) * np.dot(H1bar, H1bar.T)    # cca.py:48-49
        S11 = S11 + self.r1 * np.identity(S11.shape[0])
        S22 = (1.0 / (m - 1)) * np.dot(H2bar, H2bar.T)    # cca.py:52-53
        S22 = S22 + self.r2 * np.identity(S22.shape[0])
        S11i = np.linalg.inv(sqrtm(S11))                  # cca.py:201-202
        S22i = np.linalg.inv(sqrtm(S22))
        Tnp = S11i.dot(S12).dot(S22i)                     # cca.py:204
        U, Values, V = np.linalg.svd(Tnp)                 # cca.py:206
        self.U = S11i.dot(U)                              # cca.py:210-211
        self.V = S22i.dot(V.T)
        self.S11, self.S22, self.S12 = S11, S22, S12
        return Values

    def transform_V1(self, X):
        return np.dot(X - self.m1, self.U)                # cca.py:432-439

    def transform_V2(self, Y):
        return np.dot(Y - self.m2, self.V)                # cca.py:441-444


def cca_layer_train_forward(H1, H2, r1=1e-3, r2=1e-3, rT=1e-3, alpha=1.0, state=None, dtype=np.float64):
    """CCALayer.get_output_for(deterministic=False), forward only (cca.py:91-203).

    Returns dict(out (m,64), U, V, mean1, mean2, S11, S22, S12, corr).  `state`
    carries the previous running statistics when alpha < 1.
    """
    H1 = np.asarray(H1, dtype)
    H2 = np.asarray(H2, dtype)
    m = dtype(H1.shape[0])
    st = state or {}
    z32 = np.zeros((32,), dtype)
    z = np.zeros((H1.shape[1], H1.shape[1]), dtype)
    mean1 = (1.0 - alpha) * st.get("mean1", z32) + alpha * H1.mean(0)        # cca.py:94-103
    mean2 = (1.0 - alpha) * st.get("mean2", z32) + alpha * H2.mean(0)
    H1bar = (H1 - mean1).T                                                    # cca.py:111-116
    H2bar = (H2 - mean2).T
    S12 = (1.0 / (m - 1)) * H1bar.dot(H2bar.T)                                # cca.py:119
    S11 = (1.0 / (m - 1)) * H1bar.dot(H1bar.T) + r1 * np.eye(H1.shape[1])     # cca.py:122-123
    S22 = (1.0 / (m - 1)) * H2bar.dot(H2bar.T) + r2 * np.eye(H2.shape[1])     # cca.py:126-127
    S12 = (1.0 - alpha) * st.get("S12", z) + alpha * S12                      # cca.py:130-143
    S11 = (1.0 - alpha) * st.get("S11", z) + alpha * S11
    S22 = (1.0 - alpha) * st.get("S22", z) + alpha * S22
    d, A = np.linalg.eigh(S11)                                                # cca.py:146-149
    S11si = (A * np.reciprocal(np.sqrt(d))).dot(A.T)
    d, A = np.linalg.eigh(S22)
    S22si = (A * np.reciprocal(np.sqrt(d))).dot(A.T)
    Tnp = S11si.dot(S12).dot(S22si)                                           # cca.py:152
    M1 = Tnp.dot(Tnp.T) + rT * np.eye(Tnp.shape[0])                           # cca.py:153-156
    M2 = Tnp.T.dot(Tnp) + rT * np.eye(Tnp.shape[0])
    E1, E = np.linalg.eigh(M1)                                                # cca.py:159-160 (ascending)
    _, Fm = np.linalg.eigh(M2)
    corr = np.sqrt(np.clip(E1, 1e-7, 1.0))                                    # cca.py:163-166
    U = S11si.dot(E)                                                          # cca.py:169-170
    V = S22si.dot(Fm)
    s = np.sign(U.T.dot(S12).dot(V).diagonal())                               # cca.py:174-175
    U = U * s
    out = np.hstack((H1bar.T.dot(U), H2bar.T.dot(V)))                         # cca.py:198-201
    return dict(out=out, U=U, V=V, mean1=mean1, mean2=mean2, S11=S11, S22=S22, S12=S12, corr=corr)


def synth_latents(n, seed=23, dim=32, scale=0.05):
    """Planted-correlation latents for config 3: H = Z.A + noise, var ~4e-3."""
    rng = np.random.RandomState(seed)
    Z = rng.normal(size=(n, dim))
    A1 = rng.normal(size=(dim, dim)) / np.sqrt(dim)
    A2 = rng.normal(size=(dim, dim)) / np.sqrt(dim)
    rho = np.linspace(0.95, 0.2, dim)
    N1 = rng.normal(size=(n, dim))
    N2 = rng.normal(size=(n, dim))
    H1 = (Z * rho + N1 * np.sqrt(1 - rho ** 2)).dot(A1) * scale + rng.normal(size=dim) * 0.01
    H2 = (Z * rho + N2 * np.sqrt(1 - rho ** 2)).dot(A2) * scale + rng.normal(size=dim) * 0.01
    return H1.astype(np.float32), H2.astype(np.float32)


def cca_layer_train_backward(H1, H2, G1, G2, r1=1e-3, r2=1e-3, rT=1e-3, g_corr=None):
    """Gradient of a scalar loss through CCALayer.get_output_for(deterministic=False) with ALPHA = 1 (cca.py:91-203):
    given G1 = dL/d lv1_cca, G2 = dL/d lv2_cca_fixed (each (m, d)) and optionally g_corr = dL/d corr (d,) -- the layer's own
    loss term -wl * mean(corr) -- returns (dL/dH1, dL/dH2).  The reference gets this from Theano's automatic
    differentiation (EighGrad for the four eigendecompositions, zero gradient through sgn and outside the clip range);
    this is the same chain written out in fp64 NumPy.  Pinned by central differences of cca_layer_train_forward
    (tests/test_oracle_golden.py), not by the reference (Theano is not installable)."""
    H1 = np.asarray(H1, np.float64); H2 = np.asarray(H2, np.float64)
    G1 = np.asarray(G1, np.float64); G2 = np.asarray(G2, np.float64)
    m, d = H1.shape
    Xc, Yc = H1 - H1.mean(0), H2 - H2.mean(0)
    S12 = Xc.T.dot(Yc) / (m - 1)
    S11 = Xc.T.dot(Xc) / (m - 1) + r1 * np.eye(d)
    S22 = Yc.T.dot(Yc) / (m - 1) + r2 * np.eye(d)
    d1, A1 = np.linalg.eigh(S11)
    d2, A2 = np.linalg.eigh(S22)
    S11si = (A1 / np.sqrt(d1)).dot(A1.T)
    S22si = (A2 / np.sqrt(d2)).dot(A2.T)
    T = S11si.dot(S12).dot(S22si)
    E1, E = np.linalg.eigh(T.dot(T.T) + rT * np.eye(d))
    E2, F = np.linalg.eigh(T.T.dot(T) + rT * np.eye(d))
    U0, V = S11si.dot(E), S22si.dot(F)
    s = np.sign(U0.T.dot(S12).dot(V).diagonal())
    U = U0 * s

    def eigvec_grad(lam, Q, Qbar, lam_bar=None):
        """A = Q diag(lam) Q' symmetric: dL/dA from dL/dQ (and dL/dlam), symmetrised."""
        diff = lam[None, :] - lam[:, None]                       # diff[i, j] = lam_j - lam_i
        K = np.where(diff != 0, 1.0 / np.where(diff != 0, diff, 1.0), 0.0)
        inner = K * Q.T.dot(Qbar)
        if lam_bar is not None:
            inner = inner + np.diag(lam_bar)
        Abar = Q.dot(inner).dot(Q.T)
        return 0.5 * (Abar + Abar.T)

    def inv_sqrt_grad(lam, Q, Gbar):
        """S = Q diag(lam) Q', f(S) = S^-1/2: dL/dS from dL/df(S) (Daleckii-Krein), symmetric."""
        f = lam ** -0.5
        diff = lam[:, None] - lam[None, :]
        Phi = np.where(np.abs(diff) > 1e-12 * lam.max(), (f[:, None] - f[None, :]) / np.where(diff != 0, diff, 1.0),
                       -0.5 * (0.5 * (lam[:, None] + lam[None, :])) ** -1.5)
        Gs = 0.5 * (Gbar + Gbar.T)
        return Q.dot(Phi * Q.T.dot(Gs).dot(Q)).dot(Q.T)

    dU, dV = Xc.T.dot(G1), Yc.T.dot(G2)
    dU0 = dU * s
    dS11si = dU0.dot(E.T)
    dS22si = dV.dot(F.T)
    dE, dF = S11si.dot(dU0), S22si.dot(dV)
    lam_bar = None
    if g_corr is not None:                                         # corr = sqrt(clip(E1, 1e-7, 1)): zero outside the range
        inside = (E1 > 1e-7) & (E1 < 1.0)
        lam_bar = np.where(inside, np.asarray(g_corr, np.float64) * 0.5 / np.sqrt(np.clip(E1, 1e-7, 1.0)), 0.0)
    dM1 = eigvec_grad(E1, E, dE, lam_bar)
    dM2 = eigvec_grad(E2, F, dF)
    dT = 2.0 * dM1.dot(T) + 2.0 * T.dot(dM2)
    dS11si = dS11si + dT.dot(S22si).dot(S12.T)
    dS22si = dS22si + S12.T.dot(S11si).dot(dT)
    dS12 = S11si.dot(dT).dot(S22si)
    dS11 = inv_sqrt_grad(d1, A1, dS11si)
    dS22 = inv_sqrt_grad(d2, A2, dS22si)
    dXc = G1.dot(U.T) + (2.0 * Xc.dot(dS11) + Yc.dot(dS12.T)) / (m - 1)
    dYc = G2.dot(V.T) + (2.0 * Yc.dot(dS22) + Xc.dot(dS12)) / (m - 1)
    return dXc - dXc.mean(0), dYc - dYc.mean(0)
