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
        t = torch.nn.functional.pad(t, (0, _lib.DIM - t.shape[1]))
    return t.contiguous()


class EmbeddingDB(object):
    """One shard of an embedding DB resident in HBM: rows [idx_base, idx_base + n)."""

    def __init__(self, codes, ids=None, idx_base=0, device=None):
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
        h = ctypes.c_void_p()
        _lib.check(_lib.lib.asr_db_create(ctypes.byref(h), _lib.dptr(self.codes), self.n, self.idx_base))
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
        """Rank (1-based) of the best correct item per query + its score.  With a process group
        the DB is sharded over the ranks (queries replicated)."""
        nq = int(q.shape[0])
        ts = torch.empty(nq, dtype=torch.float32, device=self.device)
        ti = torch.empty(nq, dtype=torch.int64, device=self.device)
        better = torch.zeros(nq, dtype=torch.int64, device=self.device)
        st = _lib.stream_ptr()
        args = (self.handle, _lib.dptr(q), nq, int(q_base), int(kg), int(hg), int(bool(normalise)))
        _lib.check(_lib.lib.asr_rank_of_target(*args, 0, _lib.dptr(ts), _lib.dptr(ti), _lib.dptr(better), st))
        if group is not None:
            import torch.distributed as dist
            ws = dist.get_world_size(group)
            all_s = [torch.empty_like(ts) for _ in range(ws)]
            all_i = [torch.empty_like(ti) for _ in range(ws)]
            dist.all_gather(all_s, ts, group=group)
            dist.all_gather(all_i, ti, group=group)
            S, I = torch.stack(all_s), torch.stack(all_i)
            I = torch.where(I < 0, torch.full_like(I, 2 ** 62), I)
            best = S.max(dim=0).values
            cand = torch.where(S == best, I, torch.full_like(I, 2 ** 62))
            ti = cand.min(dim=0).values.contiguous()
            ts = best.contiguous()
        _lib.check(_lib.lib.asr_rank_of_target(*args, 1, _lib.dptr(ts), _lib.dptr(ti), _lib.dptr(better), st))
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(better, group=group)
        return better + 1, ts

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


def vote_device(cand_idx, row_ids, top_k):
    """Vote over GLOBAL row indices with a replicated row->piece table (sharded DBs)."""
    n_rec, m = int(cand_idx.shape[0]), int(cand_idx.shape[1])
    out_ids = torch.empty((n_rec, top_k), dtype=torch.int32, device=cand_idx.device)
    out_cnt = torch.empty((n_rec, top_k), dtype=torch.int32, device=cand_idx.device)
    _lib.check(_lib.lib.asr_vote(_lib.dptr(cand_idx.contiguous()), _lib.dptr(row_ids), int(row_ids.numel()), n_rec, m,
                                 int(top_k), _lib.dptr(out_ids), _lib.dptr(out_cnt), _lib.stream_ptr()))
    return out_ids, out_cnt
