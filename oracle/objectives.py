"""CPU restatement of the reference's contrastive cosine loss and its gradient.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and nothing in the product path).

Follows audio_sheet_retrieval/models/objectives.py:30-69 (`get_contrastive_cos_loss`):
    D = lv1 . lv2^T, d = diag(D), L_ij = clip(gamma - d_i + D_ij, 0, 1000) for j != i,
    loss = weight * (mean(L) [+ the same on D^T when symmetric]).
PINNED: the forward value is checked against tests/golden/reference_contrastive_loss.npz, produced by
executing the reference's own function on NumPy arrays (tests/golden/make_golden_loss.py).
The gradient has no reference text (Theano autodiff); it is pinned by central differences of the pinned
forward in tests/test_oracle_golden.py.
"""
import numpy as np


def contrastive_cos_loss_ref(lv1, lv2, weight, gamma, symmetric=False):
    lv1 = np.asarray(lv1, np.float64)
    lv2 = np.asarray(lv2, np.float64)
    n = lv1.shape[0]
    D = lv1.dot(lv2.T)
    off = ~np.eye(n, dtype=bool)

    def direction(M):
        L = np.clip(gamma - np.diag(M)[:, None] + M, 0, 1000)
        return L[off].mean()

    loss = direction(D)
    if symmetric:
        loss += direction(D.T)
    return weight * loss


def contrastive_cos_grads_ref(lv1, lv2, weight, gamma, symmetric=False):
    """d loss / d lv1, d loss / d lv2 (clip passes the gradient on the closed interval [0, 1000])."""
    lv1 = np.asarray(lv1, np.float64)
    lv2 = np.asarray(lv2, np.float64)
    n = lv1.shape[0]
    D = lv1.dot(lv2.T)
    off = ~np.eye(n, dtype=bool)
    d = np.diag(D)[:, None]

    def active(M):
        X = gamma - d + M
        return ((X >= 0) & (X <= 1000) & off).astype(np.float64)

    G = active(D)
    G -= np.diag(G.sum(1))
    if symmetric:
        A2 = active(D.T)                       # row i: terms gamma - d_i + D_ji
        G += A2.T
        G -= np.diag(A2.sum(1))
    G *= weight / (n * (n - 1))
    return G.dot(lv2), G.T.dot(lv1)
