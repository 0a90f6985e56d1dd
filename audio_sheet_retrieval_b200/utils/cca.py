"""CCA(method='svd') with the reference's interface (audio_sheet_retrieval/utils/cca.py:6-53,
199-211, 432-444): fit -> canonical correlations; attributes m1, m2, U, V; transform_V1/V2.

fit runs on the device: a fused Gram kernel accumulates the sufficient statistics (column sums, second
moments, row count) in fp64 in one pass (the reference centres, then uses an fp32 sgemm), all-reduced as
ONE buffer over a process group for row-sharded inputs, then a single-CTA Jacobi kernel centres the
moments and computes S11^-1/2, S22^-1/2, T and its SVD.  torch only holds the buffers.
Only method 'svd' is supported: it is the only one the reference ever selects
(refine_cca.py:100, utils/train_dcca_pool.py:250).
"""
import numpy as np
import torch

from .. import _lib


def cca_sums_device(H1, H2, shift1=None, shift2=None, sums=None):
    """Accumulate this shard's statistics into `sums` (fp64, 3136)."""
    if sums is None:
        sums = torch.zeros(_lib.CCA_NSUMS, dtype=torch.float64, device=H1.device)
    _lib.check(_lib.lib.asr_cca_accumulate(_lib.dptr(H1), _lib.dptr(H2), int(H1.shape[0]), _lib.dptr(shift1),
                                           _lib.dptr(shift2), _lib.dptr(sums), _lib.stream_ptr()))
    return sums


def cca_solve_device(sums, n_total, shift1=None, shift2=None, r1=1e-3, r2=1e-3, rT=1e-3, mode=0):
    dev = sums.device
    m1 = torch.empty(32, dtype=torch.float64, device=dev)
    m2 = torch.empty(32, dtype=torch.float64, device=dev)
    U = torch.empty((32, 32), dtype=torch.float64, device=dev)
    V = torch.empty((32, 32), dtype=torch.float64, device=dev)
    sig = torch.empty(32, dtype=torch.float64, device=dev)
    _lib.check(_lib.lib.asr_cca_solve(_lib.dptr(sums), int(n_total), _lib.dptr(shift1), _lib.dptr(shift2),
                                      float(r1), float(r2), float(rT), int(mode), _lib.dptr(m1), _lib.dptr(m2),
                                      _lib.dptr(U), _lib.dptr(V), _lib.dptr(sig), _lib.stream_ptr()))
    return m1, m2, U, V, sig


def cca_layer_backward_device(H1, H2, G1, G2, r1=1e-3, r2=1e-3, rT=1e-3, g_corr=None):
    """Gradient through the CCALayer's training forward (batch statistics, ALPHA = 1; layers/cca.py:91-203).
    H1, H2: the layer's inputs (m,32); G1, G2: dL/d(lv1_cca), dL/d(lv2_cca_fixed) (m,32); g_corr: dL/d corr (32,) or None.
    -> (dL/dH1, dL/dH2) float32 CUDA tensors."""
    dev = torch.device("cuda", torch.cuda.current_device())
    t = [torch.as_tensor(a).to(dev, torch.float32).contiguous() for a in (H1, H2, G1, G2)]
    m = int(t[0].shape[0])
    if any(tuple(a.shape) != (m, 32) for a in t):
        raise ValueError("expected four (m,32) arrays")
    gc = None if g_corr is None else torch.as_tensor(np.asarray(g_corr, np.float64)).to(dev).contiguous()
    d1, d2 = torch.empty_like(t[0]), torch.empty_like(t[1])
    _lib.check(_lib.lib.asr_cca_layer_backward(_lib.dptr(t[0]), _lib.dptr(t[1]), _lib.dptr(t[2]), _lib.dptr(t[3]), m,
                                               float(r1), float(r2), float(rT), _lib.dptr(gc), _lib.dptr(d1), _lib.dptr(d2),
                                               _lib.stream_ptr()))
    return d1, d2


class CCA(object):
    """Cannonical correlation analysis"""

    def __init__(self, r1=1e-3, r2=1e-3, rT=1e-3, method='svd'):
        if method != 'svd':
            raise NotImplementedError("only method='svd' is provided (the one the reference uses)")
        self.r1, self.r2, self.rT, self.method = r1, r2, rT, method
        self.m1 = self.m2 = self.U = self.V = None

    def fit(self, H1, H2, verbose=False, group=None):
        """Compute projections into correlation space.  H1, H2: (m,32) NumPy arrays or CUDA tensors
        holding this rank's rows; with `group` the statistics are all-reduced over the ranks."""
        if not torch.cuda.is_available():
            raise _lib.AsrError("no CUDA device: CCA.fit has no CPU fallback")
        dev = H1.device if hasattr(H1, "is_cuda") and H1.is_cuda else torch.device("cuda", torch.cuda.current_device())
        H1 = torch.as_tensor(H1).to(dev, torch.float32).contiguous()
        H2 = torch.as_tensor(H2).to(dev, torch.float32).contiguous()
        if H1.shape[1] != 32 or H2.shape[1] != 32 or H1.shape[0] != H2.shape[0]:
            raise ValueError("expected two (m,32) arrays")
        # One pass, one collective: Gram + column sums + row count of this rank's rows in fp64 (products of fp32 values
        # are exact in fp64), all-reduced as ONE buffer; the solve centres the moments exactly (raw-moment correction
        # in fp64) and reads the total row count from the device -- no pass for the means, no host synchronisation.
        sums = torch.zeros(_lib.CCA_NSUMS + 1, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib.asr_cca_accumulate_counted(_lib.dptr(H1), _lib.dptr(H2), int(H1.shape[0]), None, None,
                                                       _lib.dptr(sums), _lib.stream_ptr()))
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(sums, group=group)
        m1, m2, U, V, sig = cca_solve_device(sums, _lib.CCA_COUNT_ON_DEVICE, None, None, self.r1, self.r2, self.rT, mode=0)
        self.m1, self.m2 = m1.cpu().numpy(), m2.cpu().numpy()
        self.U, self.V = U.cpu().numpy(), V.cpu().numpy()
        coeffs = sig.cpu().numpy()
        if verbose:
            print("\nCorrelation-Coeffs:  ", np.around(coeffs, 3))
            print("Canonical-Correlation:", np.sum(coeffs) / H1.shape[1])
        return coeffs

    def transform(self, X):
        """Project data into cca space"""
        return np.dot(X - self.m1, self.U)

    def transform_V1(self, X):
        return self.transform(X)

    def transform_V2(self, Y):
        return np.dot(Y - self.m2, self.V)
