"""CCA(method='svd') with the reference's interface (audio_sheet_retrieval/utils/cca.py:6-53,
199-211, 432-444): fit -> canonical correlations; attributes m1, m2, U, V; transform_V1/V2.

fit runs on the device: a fused Gram kernel accumulates the sufficient statistics in fp64 (the
reference uses an fp32 sgemm on centred data), optionally all-reduced over a process group for
row-sharded inputs, then a single-CTA Jacobi kernel computes S11^-1/2, S22^-1/2, T and its SVD.
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
        n = torch.tensor([H1.shape[0]], dtype=torch.float64, device=dev)
        # pass 1: means (as the reference, which centres before its sgemm)
        s0 = torch.zeros(64, dtype=torch.float64, device=dev)
        s0[:32] = H1.sum(dim=0, dtype=torch.float64)
        s0[32:] = H2.sum(dim=0, dtype=torch.float64)
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(n, group=group)
            dist.all_reduce(s0, group=group)
        n_total = int(n.item())
        shift = (s0 / n_total).to(torch.float32)
        sh1, sh2 = shift[:32].contiguous(), shift[32:].contiguous()
        # pass 2: centred second moments (fp64 accumulation)
        sums = cca_sums_device(H1, H2, sh1, sh2)
        if group is not None:
            dist.all_reduce(sums, group=group)
        m1, m2, U, V, sig = cca_solve_device(sums, n_total, sh1, sh2, self.r1, self.r2, self.rT, mode=0)
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
