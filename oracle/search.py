"""Oracle: brute-force cosine search, pinned-order fp32 scoring, per-piece vote (NumPy, CPU).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  asr/audio_sheet_server.py:530-551   _retrieve_sheet_snippet_ids  (cdist cosine, argsort[:n])
  asr/audio_sheet_server.py:553-563   _retrieve_perform_excerpt_ids
  asr/audio_sheet_server.py:213-253   detect_score  (100 linspace windows, per-window search, vote)
  asr/audio_sheet_server.py:255-300   detect_performance

The CUDA path is bit-exact against the *pinned-order fp32* definition below
(BASELINE.json north_star: "bit-exact against a NumPy fp32 ranking of the same
embeddings, with ties broken by index"):

  ss    = (((x0*x0) + x1*x1) + ...) + x31*x31     every * and + rounded to fp32 separately (no FMA)
  inv   = fp32(1) / sqrt(ss)                       IEEE fp32 sqrt and divide
  xn_k  = x_k * inv
  score = (((qn0*dn0) + qn1*dn1) + ...) + qn31*dn31   sequential k = 0..31, no FMA
  NaN scores (zero rows) rank as -inf
  order = score descending, then index ascending
"""
import numpy as np


def pinned_normalise(X):
    X = np.ascontiguousarray(X, np.float32)
    ss = np.zeros(X.shape[0], np.float32)
    for k in range(X.shape[1]):
        ss = ss + X[:, k] * X[:, k]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.float32(1.0) / np.sqrt(ss)
        return (X * inv[:, None]).astype(np.float32)


def pinned_scores(Q, D, normalise=True, chunk=1 << 16):
    """(nq, nd) fp32 score matrix with the pinned accumulation order."""
    Q = np.ascontiguousarray(Q, np.float32)
    D = np.ascontiguousarray(D, np.float32)
    if normalise:
        Q, D = pinned_normalise(Q), pinned_normalise(D)
    out = np.empty((Q.shape[0], D.shape[0]), np.float32)
    with np.errstate(invalid="ignore"):
        for s in range(0, D.shape[0], chunk):
            Dc = D[s:s + chunk]
            acc = np.zeros((Q.shape[0], Dc.shape[0]), np.float32)
            for k in range(Q.shape[1]):
                acc = acc + Q[:, k, None] * Dc[None, :, k]
            out[:, s:s + chunk] = acc
    out[np.isnan(out)] = -np.inf
    return out


def pinned_topk(Q, D, k, normalise=True, idx_base=0):
    """(scores (nq,k) fp32, indices (nq,k) int64), (score desc, index asc).  Rows
    beyond the DB size are filled with (-inf, -1)."""
    S = pinned_scores(Q, D, normalise=normalise)
    nq, nd = S.shape
    kk = min(k, nd)
    order = np.argsort(-S, axis=1, kind="stable")[:, :kk]   # stable on -S => ties by index ascending
    sc = np.take_along_axis(S, order, axis=1)
    scores = np.full((nq, k), -np.inf, np.float32)
    idx = np.full((nq, k), -1, np.int64)
    scores[:, :kk] = sc
    idx[:, :kk] = order + idx_base
    return scores, idx


def merge_topk(score_lists, idx_lists, k):
    """Merge per-shard (nq,k) candidate lists: what the all-gather + merge step computes."""
    S = np.concatenate(score_lists, axis=1)
    I = np.concatenate(idx_lists, axis=1)
    out_s = np.full((S.shape[0], k), -np.inf, np.float32)
    out_i = np.full((S.shape[0], k), -1, np.int64)
    for q in range(S.shape[0]):
        valid = I[q] >= 0
        s, i = S[q][valid], I[q][valid]
        order = np.lexsort((i, -s))[:k]
        out_s[q, :len(order)] = s[order]
        out_i[q, :len(order)] = i[order]
    return out_s, out_i


def retrieve_ids_ref(db_codes, db_ids, code, n_candidates=1):
    """asr/audio_sheet_server.py:530-551 call for call (fp64 cdist; argsort made stable)."""
    from scipy.spatial.distance import cdist
    dists = cdist(db_codes, code, metric="cosine").flatten()          # :534
    sorted_idx = np.argsort(dists, kind="stable")[:n_candidates]      # :537
    return db_ids[sorted_idx], sorted_idx                              # :551


def vote_ref(all_piece_ids, top_k):
    """asr/audio_sheet_server.py:237-240,249-251.  argsort made stable, so after the
    reversal ties in the vote count resolve to the LARGER piece id first."""
    unique, counts = np.unique(all_piece_ids, return_counts=True)     # :237
    sorted_count_idxs = np.argsort(counts, kind="stable")[::-1][:top_k]  # :240
    ids = unique[sorted_count_idxs]
    votes = counts[sorted_count_idxs]
    shares = np.asarray(votes, dtype=np.float64) / np.sum(votes)      # :250-251
    return ids, votes, shares


def window_starts(total, width, n_samples=100):
    """asr/audio_sheet_server.py:216-218 / 260-262."""
    return np.linspace(start=0, stop=total - width, num=n_samples).astype(int)


def detect_ref(query_codes, db_codes, db_ids, top_k=1, n_candidates=1):
    """detect_score / detect_performance after the embedding step (:229-253 / :278-300)."""
    all_piece_ids = np.zeros(0, dtype=np.int64)
    for i in range(len(query_codes)):
        piece_ids, _ = retrieve_ids_ref(db_codes, db_ids, query_codes[i:i + 1], n_candidates)
        all_piece_ids = np.concatenate((all_piece_ids, piece_ids))
    return vote_ref(all_piece_ids, top_k)


def detect_pinned(query_codes, db_codes, db_ids, top_k=1, n_candidates=1):
    """Same protocol with the pinned fp32 ranking (what the CUDA path reproduces exactly)."""
    _, idx = pinned_topk(query_codes, db_codes, n_candidates)
    ids = db_ids[idx[idx >= 0]]
    return vote_ref(ids, top_k)


def synth_piece_db(n_pieces, per_piece, n_rec, windows, seed_db=1, seed_q=2, sigma=0.12, dim=32):
    """Config-4 style data: unit-norm Gaussian DB, piece ids, noisy queries."""
    rng = np.random.RandomState(seed_db)
    db = rng.normal(size=(n_pieces * per_piece, dim)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    ids = np.repeat(np.arange(n_pieces, dtype=np.int32), per_piece)
    rq = np.random.RandomState(seed_q)
    true_piece = rq.randint(0, n_pieces, n_rec)
    rows = np.stack([p * per_piece + rq.randint(0, per_piece, windows) for p in true_piece])
    q = db[rows.reshape(-1)] + rq.normal(0, sigma, (n_rec * windows, dim)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return db, ids, q.astype(np.float32), true_piece
