"""dtw_by_dist with the reference's signature and return value
(audio_sheet_retrieval/utils/dtw_by_dist.py:6-34): the accumulated-cost matrix is filled by an
anti-diagonal wavefront kernel and the path traced back on the device (libasr_b200.so: asr_dtw)."""
import numpy as np
import torch

from .. import _lib


def dtw_device(dist):
    """dist: (r,c) float64 CUDA tensor -> (acc (r,c) float64 tensor, path_i, path_j NumPy int arrays)."""
    r, c = int(dist.shape[0]), int(dist.shape[1])
    acc = torch.empty_like(dist)
    pi = torch.empty(r + c - 1, dtype=torch.int32, device=dist.device)
    pj = torch.empty(r + c - 1, dtype=torch.int32, device=dist.device)
    n = torch.zeros(1, dtype=torch.int32, device=dist.device)
    _lib.check(_lib.lib.asr_dtw(_lib.dptr(dist), r, c, _lib.dptr(acc), _lib.dptr(pi), _lib.dptr(pj), _lib.dptr(n),
                                _lib.stream_ptr()))
    n = int(n.item())
    return acc, pi[:n].cpu().numpy().astype(np.int64), pj[:n].cpu().numpy().astype(np.int64)


def dtw_by_dist(dist):
    """
    Computes Dynamic Time Warping (DTW) on distance matrix
    :param dist: distance matrix
    Returns the minimum distance, the cost matrix, the accumulated cost matrix, and the wrap path.
    """
    if not torch.cuda.is_available():
        raise _lib.AsrError("no CUDA device: dtw_by_dist has no CPU fallback")
    dist = np.asarray(dist, dtype=np.float64)
    transposed = False
    if dist.shape[1] > dist.shape[0]:                       # :13-15
        dist = dist.T
        transposed = True
    C = np.ascontiguousarray(dist)
    acc, p, q = dtw_device(torch.as_tensor(C).cuda())
    D1 = acc.cpu().numpy()
    path = (p, q)
    if not transposed:                                      # :31-32 (the reference swaps in the NON-transposed case)
        path = (path[1], path[0])
    return D1[-1, -1] / sum(D1.shape), C, D1, path
