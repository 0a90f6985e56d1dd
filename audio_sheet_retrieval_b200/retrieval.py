"""Device-resident embedding database: fused normalise + dot + top-k, rank-of-target, vote.

Host mirror of asr_db_* / asr_topk / asr_rank_of_target / asr_vote (include/asr_b200.h).  Replaces
the `cdist` + `argsort` calls of audio_sheet_retrieval/audio_sheet_server.py:530-563 and
utils/train_dcca_pool.py:39-74.  torch is used for device memory and (optionally) the NCCL
exchange of per-shard results; all arithmetic is in libasr_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _as_codes(x, device):
    """(n, d<=32) float32 -> (n,32) contiguous CUDA tensor (zero padded: zeros add exactly nothing
    to the pinned-order sums, so `--max_dim` clipping keeps bit-exact scores)."""
    t = torch.as_tensor(x)
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2 or t.shape[1] > _lib.DIM or t.shape[1] < 1:
        raise ValueError("codes must be (n, d) with 1 <= d <= 32, got %s" % (tuple(t.shape),))
    t = t.to(device, non_blocking=True)
    if t.shape[1] < _lib.DIM:
        full = torch.zeros((t.shape[0], _lib.DIM), dtype=torch.float32, device=device)
        full[:, :t.shape[1]].copy_(t)
        t = full
    return t.contiguous()


def chunk_layout(nq, k):
    """Byte layout of one rank's result in an all-gather: [scores (nq,k) f32 | pad to 8 | indices (nq,k) i64].
    -> (chunk_bytes, idx_offset_bytes); the layout asr_topk_merge_gathered / asr_rank_target_merge read."""
    idx_off = (nq * k * 4 + 7) // 8 * 8
    return idx_off + nq * k * 8, idx_off


def chunk_views(buf, nq, k, rank=0):
    """(scores (nq,k) f32, indices (nq,k) i64) views of chunk `rank` of a uint8 buffer laid out by chunk_layout."""
    chunk, idx_off = chunk_layout(nq, k)
    c = buf[rank * chunk:(rank + 1) * chunk]
    return c[:nq * k * 4].view(torch.float32).view(nq, k), c[idx_off:idx_off + nq * k * 8].view(torch.int64).view(nq, k)


class EmbeddingDB(object):
    """One shard of an embedding DB resident in HBM: rows [idx_base, idx_base + n)."""

    def __init__(self, codes, ids=None, idx_base=0, device=None, normalise_in_place=False, cosine_copy=True,
                 max_queries=16384):
        """normalise_in_place: donate `codes` (a CUDA tensor) -- its rows are overwritten with their pinned-normalised
        form and streamed from there, no second copy (a 1e8-row DB stays at 12.8 GB); cosine queries only.
        cosine_copy=False: no normalised rows at all (cosine queries normalise in-kernel, exact kernel only).
        max_queries: size of the query workspace (larger calls are processed in chunks)."""
        if not torch.cuda.is_available():
            raise _lib.AsrError("no CUDA device: the retrieval path has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.codes = _as_codes(codes, self.device)
        if self.codes.shape[0] == 0:
            raise ValueError("empty database")
        if self.codes.data_ptr() % 128:
            buf = torch.empty(self.codes.numel() + 32, dtype=torch.float32, device=self.device)
            off = (-buf.data_ptr() % 128) // 4
            buf[off:off + self.codes.numel()].copy_(self.codes.view(-1))
            self.codes = buf[off:off + self.codes.numel()].view(-1, _lib.DIM)
        self.n = int(self.codes.shape[0])
        self.idx_base = int(idx_base)
        self.ids = None
        if ids is not None:
            self.ids = torch.as_tensor(np.asarray(ids)).to(torch.int32).to(self.device).contiguous()
            if self.ids.numel() != self.n:
                raise ValueError("ids must have one entry per DB row")
        flags = (_lib.DB_NORMALISE_IN_PLACE if normalise_in_place else 0) | (0 if cosine_copy else _lib.DB_NO_COSINE_COPY)
        h = ctypes.c_void_p()
        _lib.check(_lib.lib.asr_db_create_ex(ctypes.byref(h), _lib.dptr(self.codes), self.n, self.idx_base, flags,
                                             int(max_queries)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib.asr_db_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- top-k ----------------------------------------------------------------------------
    def topk_device(self, q, k, normalise=True, out_scores=None, out_idx=None, stream=None):
        """q: (nq,32) CUDA float32.  Returns (scores (nq,k) f32, indices (nq,k) i64) CUDA tensors."""
        nq = int(q.shape[0])
        if out_scores is None:
            out_scores = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        if out_idx is None:
            out_idx = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        _lib.check(_lib.lib.asr_topk(self.handle, _lib.dptr(q), nq, int(k), int(bool(normalise)),
                                     _lib.dptr(out_scores), _lib.dptr(out_idx), _lib.stream_ptr(stream)))
        return out_scores, out_idx

    def topk(self, queries, k, normalise=True):
        """Host arrays in/out (the call a user of `_retrieve_*` makes)."""
        if not 1 <= k <= _lib.MAX_K:
            raise ValueError("k must be in [1, %d]" % _lib.MAX_K)
        q = _as_codes(queries, self.device)
        s, i = self.topk_device(q, k, normalise)
        return s.cpu().numpy(), i.cpu().numpy()

    # -- eval_retrieval ranks ----------------------------------------------------------------
    def ranks_device(self, q, kg=1, hg=1, q_base=0, normalise=True, group=None):
        """Per query: how many DB rows rank BEFORE its best correct item (rank = count + 1) and that item's
        score.  With a process group the DB is sharded over the ranks (queries replicated)."""
        nq = int(q.shape[0])
        ts = torch.empty(nq, dtype=torch.float32, device=self.device)
        ti = torch.empty(nq, dtype=torch.int64, device=self.device)
        better = torch.zeros(nq, dtype=torch.int64, device=self.device)
        st = _lib.stream_ptr()
        args = (self.handle, _lib.dptr(q), nq, int(q_base), int(kg), int(hg), int(bool(normalise)))
        _lib.check(_lib.lib.asr_rank_of_target(*args, 0, _lib.dptr(ts), _lib.dptr(ti), _lib.dptr(better), st))
        if group is not None:
            # one all-gather of [tscore | tidx] chunks, then one kernel picks the best correct item over the shards
            import torch.distributed as dist
            ws = dist.get_world_size(group)
            chunk, idx_off = chunk_layout(nq, 1)
            mine = torch.empty(chunk, dtype=torch.uint8, device=self.device)
            cs, ci = chunk_views(mine, nq, 1)
            cs.view(-1).copy_(ts); ci.view(-1).copy_(ti)
            gathered = torch.empty(ws * chunk, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(gathered, mine, group=group)
            _lib.check(_lib.lib.asr_rank_target_merge(_lib.dptr(gathered), chunk, idx_off, ws, nq, _lib.dptr(ts),
                                                      _lib.dptr(ti), st))
        _lib.check(_lib.lib.asr_rank_of_target(*args, 1, _lib.dptr(ts), _lib.dptr(ti), _lib.dptr(better), st))
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(better, group=group)
        return better, ts

    # -- vote -------------------------------------------------------------------------------
    def vote_device(self, cand_idx, top_k):
        """cand_idx (n_rec, m) int64 LOCAL row indices (-1 = empty) -> (piece ids, counts) (n_rec, top_k)."""
        if self.ids is None:
            raise ValueError("this DB was created without piece ids")
        n_rec, m = int(cand_idx.shape[0]), int(cand_idx.shape[1])
        out_ids = torch.empty((n_rec, top_k), dtype=torch.int32, device=self.device)
        out_cnt = torch.empty((n_rec, top_k), dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib.asr_vote(_lib.dptr(cand_idx.contiguous()), _lib.dptr(self.ids), self.n, n_rec, m,
                                     int(top_k), _lib.dptr(out_ids), _lib.dptr(out_cnt), _lib.stream_ptr()))
        return out_ids, out_cnt


def merge_topk_device(scores, idx, n_lists, k):
    """(nq, n_lists*k) candidate lists -> (nq,k): the step after an all-gather of per-GPU top-k."""
    nq = int(scores.shape[0])
    out_s = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=scores.device)
    _lib.check(_lib.lib.asr_topk_merge(_lib.dptr(scores.contiguous()), _lib.dptr(idx.contiguous()), nq, int(n_lists),
                                       int(k), _lib.dptr(out_s), _lib.dptr(out_i), _lib.stream_ptr()))
    return out_s, out_i


def merge_gathered_topk_device(gathered, nq, n_lists, k):
    """The result of ONE all-gather of chunk_layout chunks (rank-major, as all_gather_into_tensor leaves it)
    -> merged (nq,k) lists; the kernel reads the chunks where they are (no transpose copies)."""
    chunk, idx_off = chunk_layout(nq, k)
    out_s = torch.empty((nq, k), dtype=torch.float32, device=gathered.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=gathered.device)
    _lib.check(_lib.lib.asr_topk_merge_gathered(_lib.dptr(gathered), chunk, idx_off, nq, int(n_lists), int(k),
                                                _lib.dptr(out_s), _lib.dptr(out_i), _lib.stream_ptr()))
    return out_s, out_i


def vote_device(cand_idx, row_ids, top_k, out_ids=None, out_counts=None):
    """Vote over GLOBAL row indices with a replicated row->piece table (sharded DBs).  out_ids / out_counts: optional
    contiguous (n_rec, top_k) int32 CUDA tensors to write into (e.g. this rank's part of a gather buffer)."""
    n_rec, m = int(cand_idx.shape[0]), int(cand_idx.shape[1])
    out_ids = torch.empty((n_rec, top_k), dtype=torch.int32, device=cand_idx.device) if out_ids is None else out_ids
    out_cnt = torch.empty((n_rec, top_k), dtype=torch.int32, device=cand_idx.device) if out_counts is None else out_counts
    assert out_ids.is_contiguous() and out_cnt.is_contiguous() and out_ids.dtype == torch.int32 and out_cnt.dtype == torch.int32
    _lib.check(_lib.lib.asr_vote(_lib.dptr(cand_idx.contiguous()), _lib.dptr(row_ids), int(row_ids.numel()), n_rec, m,
                                 int(top_k), _lib.dptr(out_ids), _lib.dptr(out_cnt), _lib.stream_ptr()))
    return out_ids, out_cnt
