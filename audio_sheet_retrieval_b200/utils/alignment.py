"""Audio-to-sheet alignment helpers with the reference's names
(audio_sheet_retrieval/utils/alignment.py:113-186): baseline / DTW alignment of a cosine-distance
matrix between sheet-window codes and audio-window codes, interpolation, pixel error."""
import numpy as np
import torch

from .. import _lib
from .dtw_by_dist import dtw_by_dist


def cosine_distances(img_codes, spec_codes):
    """cdist(img_codes, spec_codes, metric='cosine') as float64, computed on the device."""
    if not torch.cuda.is_available():
        raise _lib.AsrError("no CUDA device: cosine_distances has no CPU fallback")
    a = torch.as_tensor(np.ascontiguousarray(img_codes, np.float32)).cuda()
    b = torch.as_tensor(np.ascontiguousarray(spec_codes, np.float32)).cuda()
    if a.shape[1] != 32 or b.shape[1] != 32:
        raise ValueError("codes must be 32-dimensional")
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float64, device=a.device)
    _lib.check(_lib.lib.asr_cosine_distances(_lib.dptr(a), int(a.shape[0]), _lib.dptr(b), int(b.shape[0]), _lib.dptr(out),
                                             _lib.stream_ptr()))
    return out.cpu().numpy()


def align_baseline(dists):
    """ Compute alignment baseline by interpolation (:113-117) """
    i1_sheet = dists.shape[0]
    return np.linspace(start=0, stop=i1_sheet - 1, num=dists.shape[1])


def align_pydtw(dists):
    """ DTW alignment (:120-141): for every audio column the first path entry that reaches it """
    min_dist, C, C_acc, path = dtw_by_dist(dists)
    align_sheet_idxs = []
    for i in range(dists.shape[1]):
        sheet_idx = np.nonzero(path[0] == i)[0][0]
        align_sheet_idxs.append(path[1][sheet_idx])
    return np.array(align_sheet_idxs)


def compute_alignment(img_codes, spec_codes, sheet_idxs, spec_idxs, align_by):
    """ Evaluate Alignment (:143-174) """
    dists = cosine_distances(img_codes, spec_codes)
    if align_by == 'baseline':
        aligned_sheet_idxs = align_baseline(dists)
    elif align_by == 'pydtw':
        aligned_sheet_idxs = align_pydtw(dists)
    else:
        raise ValueError("align_by must be 'baseline' or 'pydtw'")
    aligned_sheet_idxs = np.round(aligned_sheet_idxs).astype(int)
    aligned_sheet_coords = sheet_idxs[aligned_sheet_idxs]
    filterd_idxs = np.diff(np.concatenate((spec_idxs[0:1] - 1, spec_idxs))) > 0
    i_inter = np.arange(spec_idxs[0], spec_idxs[-1] + 1, 1)
    a2s_alignment = np.interp(i_inter, spec_idxs[filterd_idxs], aligned_sheet_coords[filterd_idxs])
    a2s_mapping = dict(zip(i_inter, a2s_alignment))
    dtw_res = {"dists": dists, "aligned_sheet_idxs": aligned_sheet_idxs,
               "aligned_sheet_coords": aligned_sheet_coords, "i_inter": i_inter,
               "a2s_alignment": a2s_alignment, "spec_idxs": spec_idxs}
    return a2s_mapping, dtw_res


def estimate_alignment_error(true_coords, true_onsets, a2s_mapping):
    """ Compute alignment error measures (:177-186) """
    pxl_errors = np.zeros(len(true_onsets))
    for j, o in enumerate(true_onsets):
        if o in a2s_mapping:
            pxl_errors[j] = true_coords[j] - a2s_mapping[int(o)]
    return pxl_errors
