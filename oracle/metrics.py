"""Oracle: ranking metrics (NumPy/SciPy, CPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  asr/utils/train_dcca_pool.py:28-82   eval_retrieval
  asr/run_eval.py:177-212              recall_at_k / YAML result schema

Two forms are provided:
  * eval_retrieval_ref  - the reference algorithm call for call (fp64 cdist,
    per-row argsort).  NumPy's default argsort leaves ties unspecified; this
    restatement uses kind='stable' so ties resolve by index.
  * ranks_pinned        - the deterministic definition the CUDA path is
    bit-exact against: fp32 pinned-order cosine scores (oracle/search.py),
    rank = 1 + #{j : (s_ij, -j) beats the best correct item}.
"""
import numpy as np

from .search import pinned_scores

HIT_KEYS = (1, 5, 10, 25)


def eval_retrieval_ref(lv1_cca, lv2_cca):
    """asr/utils/train_dcca_pool.py:28-82 (Python-3 integer division for k, h)."""
    from scipy.spatial.distance import cdist
    n_v1 = lv1_cca.shape[0]
    n_v2 = lv2_cca.shape[0]
    k = n_v2 // n_v1 if n_v2 > n_v1 else 1                 # :35-36
    h = n_v1 // n_v2 if n_v1 > n_v2 else 1
    dists = cdist(lv1_cca, lv2_cca, metric="cosine")       # :40
    ranks, aps = [], []
    hit_rates = {1: 0, 5: 0, 10: 0, 25: 0}
    for i in range(n_v1):
        i_fixed = np.floor_divide(i, h)                    # :49
        sorted_idx = np.argsort(dists[i], kind="stable")   # :52
        for key in hit_rates:
            top_k_results = np.floor_divide(sorted_idx[0:key], k)   # :57-60
            if i_fixed in top_k_results:
                hit_rates[key] += 1
        fixed_sorted_idx = np.floor_divide(sorted_idx, k)  # :67-68
        rank = np.min(np.nonzero(fixed_sorted_idx == i_fixed)[0]) + 1
        ranks.append(rank)
        aps.append(1.0 / rank)
    mean_rank = np.mean(ranks)
    median_rank = np.median(ranks)
    mean_dist = np.diag(dists).mean()
    mrr = np.mean(aps)
    return mean_rank, median_rank, mean_dist, hit_rates, mrr


def ranks_pinned(lv1, lv2, normalise=True):
    """Deterministic ranks: for query i (row of lv1) the correct items are the
    columns j with j // k == i // h.  rank_i = 1 + #{j : (s_ij > s*) or (s_ij == s* and j < j*)}
    where (s*, j*) is the best correct item under (score desc, index asc)."""
    n_v1, n_v2 = lv1.shape[0], lv2.shape[0]
    k = n_v2 // n_v1 if n_v2 > n_v1 else 1
    h = n_v1 // n_v2 if n_v1 > n_v2 else 1
    S = pinned_scores(lv1, lv2, normalise=normalise)      # (n_v1, n_v2) fp32
    ranks = np.zeros(n_v1, np.int64)
    tscore = np.zeros(n_v1, np.float32)
    cols = np.arange(n_v2)
    for i in range(n_v1):
        g = i // h
        correct = np.arange(g * k, min((g + 1) * k, n_v2))
        sc = S[i, correct]
        b = int(np.argmax(sc))                             # first max == smallest index among ties
        s_star, j_star = sc[b], correct[b]
        better = (S[i] > s_star) | ((S[i] == s_star) & (cols < j_star))
        ranks[i] = 1 + int(better.sum())
        tscore[i] = s_star
    return ranks, tscore


def metrics_from_ranks(ranks, target_scores=None):
    """(mean_rank, median_rank, mean_dist, hit_rates, mrr) in eval_retrieval's return order."""
    ranks = np.asarray(ranks)
    hit_rates = {key: int((ranks <= key).sum()) for key in HIT_KEYS}
    mean_dist = float(np.mean(1.0 - np.asarray(target_scores, np.float64))) if target_scores is not None else None
    return float(np.mean(ranks)), float(np.median(ranks)), mean_dist, hit_rates, float(np.mean(1.0 / ranks))
