"""eval_retrieval with the reference's signature and return value
(audio_sheet_retrieval/utils/train_dcca_pool.py:28-82).

The reference materialises the N x N fp64 cosine-distance matrix and argsorts every row.  Here the
view-2 codes become a device-resident DB and one fused kernel counts, per query, how many items
rank before the best correct one (score desc, index asc on pinned-order fp32 cosine scores):
rank = 1 + count.  Hit rates, mean/median rank and MRR follow from the ranks.  The training loop
of the reference file is out of scope.
"""
import numpy as np
import torch

from ..retrieval import EmbeddingDB, _as_codes


def retrieval_ranks(lv1_cca, lv2_cca, group=None):
    """-> (ranks (n_v1,) int64, target cosine scores (n_v1,) float32) as NumPy arrays.

    With a process group every rank passes the same arrays and searches only its contiguous shard of the
    view-2 rows: the best correct item is agreed on with one all-gather + asr_rank_target_merge, the
    per-shard "ranked before it" counts are summed with one all-reduce."""
    n_v1, n_v2 = lv1_cca.shape[0], lv2_cca.shape[0]
    k = n_v2 // n_v1 if n_v2 > n_v1 else 1      # :35-36 (Python-2 integer division)
    h = n_v1 // n_v2 if n_v1 > n_v2 else 1
    lo, hi = 0, n_v2
    if group is not None:
        import torch.distributed as dist
        from ..dist import shard_bounds
        lo, hi = shard_bounds(n_v2, dist.get_rank(group), dist.get_world_size(group))
    db = EmbeddingDB(lv2_cca[lo:hi], idx_base=lo)
    try:
        q = _as_codes(lv1_cca, db.device)
        before, ts = db.ranks_device(q, kg=k, hg=h, normalise=True, group=group)
        return before.cpu().numpy() + 1, ts.cpu().numpy()
    finally:
        db.close()


def eval_retrieval(lv1_cca, lv2_cca, group=None):
    """Compute retrieval eval measures -> (mean_rank, median_rank, mean_dist, hit_rates, map)."""
    ranks, ts = retrieval_ranks(np.asarray(lv1_cca), np.asarray(lv2_cca), group=group)
    hit_rates = {1: 0, 5: 0, 10: 0, 25: 0}
    for key in hit_rates:
        hit_rates[key] = int((ranks <= key).sum())
    mean_rank = np.mean(ranks)
    median_rank = np.median(ranks)
    # the reference reports the mean of the distance-matrix diagonal (:79); with grouped views
    # that diagonal is not the correct pair, here it is always the best correct item
    mean_dist = float(np.mean(1.0 - ts.astype(np.float64)))
    map_ = np.mean(1.0 / ranks)
    return mean_rank, median_rank, mean_dist, hit_rates, map_
