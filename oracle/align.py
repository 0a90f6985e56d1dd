"""Oracle: DTW alignment (NumPy, CPU).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  asr/utils/dtw_by_dist.py:6-34    dtw_by_dist (accumulated cost, transposition quirk, normalised distance)
  asr/utils/dtw_by_dist.py:69-83   _traceback (argmin of (diag, up, left), first minimum wins)
  asr/utils/alignment.py:113-174   align_baseline, align_pydtw, compute_alignment
Pinned: tests/golden/reference_numpy_paths.npz holds outputs of the reference's own source for these.
"""
import numpy as np


def _traceback(D):
    i, j = np.array(D.shape) - 2
    p, q = [i], [j]
    while (i > 0) or (j > 0):
        tb = np.argmin((D[i, j], D[i, j + 1], D[i + 1, j]))
        if tb == 0:
            i -= 1
            j -= 1
        elif tb == 1:
            i -= 1
        else:
            j -= 1
        p.insert(0, i)
        q.insert(0, j)
    return np.array(p), np.array(q)


def dtw_by_dist(dist):
    transposed = False
    if dist.shape[1] > dist.shape[0]:
        dist = dist.T
        transposed = True
    r, c = dist.shape
    D0 = np.zeros((r + 1, c + 1))
    D0[0, 1:] = np.inf
    D0[1:, 0] = np.inf
    D0[1:, 1:] = dist
    D1 = D0[1:, 1:]
    C = D1.copy()
    for i in range(r):
        for j in range(c):
            D1[i, j] += min(D0[i, j], D0[i, j + 1], D0[i + 1, j])
    path = _traceback(D0)
    if not transposed:
        path = (path[1], path[0])
    return D1[-1, -1] / sum(D1.shape), C, D1, path


def align_pydtw(dists):
    _, _, _, path = dtw_by_dist(dists)
    out = []
    for i in range(dists.shape[1]):
        sheet_idx = np.nonzero(path[0] == i)[0][0]
        out.append(path[1][sheet_idx])
    return np.array(out)


def compute_alignment(img_codes, spec_codes, sheet_idxs, spec_idxs, align_by):
    from scipy.interpolate import interp1d
    from scipy.spatial.distance import cdist
    dists = cdist(img_codes, spec_codes, metric="cosine")
    if align_by == "baseline":
        aligned = np.linspace(start=0, stop=dists.shape[0] - 1, num=dists.shape[1])
    else:
        aligned = align_pydtw(dists)
    aligned = np.round(aligned).astype(int)
    coords = sheet_idxs[aligned]
    keep = np.diff(np.concatenate((spec_idxs[0:1] - 1, spec_idxs))) > 0
    f_inter = interp1d(spec_idxs[keep], coords[keep])
    i_inter = np.arange(spec_idxs[0], spec_idxs[-1] + 1, 1)
    return dict(zip(i_inter, f_inter(i_inter))), dict(dists=dists, aligned_sheet_idxs=aligned)


def synth_alignment_problem(n_sheet=120, n_audio=90, seed=0):
    """Monotone ground-truth warp between sheet-window codes and audio-window codes + noise."""
    rng = np.random.RandomState(seed)
    base = rng.normal(size=(n_sheet, 32))
    base = np.cumsum(base, axis=0)                       # smooth trajectory: neighbours are similar
    warp = np.sort(rng.uniform(0, n_sheet - 1, n_audio))
    warp[0], warp[-1] = 0, n_sheet - 1
    idx = np.round(warp).astype(int)
    img = base + 0.05 * rng.normal(size=base.shape)
    spec = base[idx] + 0.05 * rng.normal(size=(n_audio, 32))
    return img.astype(np.float32), spec.astype(np.float32), idx
