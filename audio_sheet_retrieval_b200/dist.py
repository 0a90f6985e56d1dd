"""Multi-GPU plumbing (one process per GPU, torch.distributed / NCCL).

The path shards in three places: snippet batches through the encoders (no
collective), CCA covariance sums (one all-reduce of [sums | row count], utils/cca.py), and the
embedding DB: each rank holds a contiguous row shard, writes its local top-k straight into one
contiguous chunk, ONE all-gather moves the chunks, and a merge kernel reads them where they are
(rank-major, no transpose copies); the piece vote runs on the merged candidates.

When the DB fits on every GPU (a 10^6-row DB is 128 MB, a 10^8-row one 12.8 GB of 180 GB) and there are many
recordings to identify, `ReplicatedDB` splits the RECORDINGS over the ranks instead: every rank runs
retrieval + vote for its recordings against the whole DB and only the (recordings x top_k) results are
gathered -- no candidate exchange, no merge, and per-rank work items that are as long as on one GPU.
"""
import os

import torch
import torch.distributed as dist

from .retrieval import EmbeddingDB, chunk_layout, chunk_views, merge_gathered_topk_device, vote_device


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"))
    return rank, world, local


def shard_bounds(n, rank, world):
    """Contiguous row range of `rank` in an n-row DB."""
    return n * rank // world, n * (rank + 1) // world


class ShardedDB(object):
    """A DB whose rows are split over the ranks of `group` (queries replicated)."""

    def __init__(self, local_codes, idx_base, row_ids_global=None, group=None, normalise_in_place=False):
        self.group = group
        self.world = dist.get_world_size(group) if (group is not None or dist.is_initialized()) else 1
        self.local = EmbeddingDB(local_codes, idx_base=idx_base, normalise_in_place=normalise_in_place)
        self.row_ids = None
        self._bufs = {}
        if row_ids_global is not None:
            self.row_ids = torch.as_tensor(row_ids_global).to(torch.int32).to(self.local.device).contiguous()

    def _buffers(self, nq, k):
        """(my chunk, gathered chunks) for (nq, k), kept across calls: the query loop allocates nothing."""
        key = (nq, k)
        if key not in self._bufs:
            chunk, _ = chunk_layout(nq, k)
            dev = self.local.device
            self._bufs[key] = (torch.empty(chunk, dtype=torch.uint8, device=dev),
                               torch.empty(self.world * chunk, dtype=torch.uint8, device=dev))
        return self._bufs[key]

    def topk_device(self, q, k, events=None):
        """Local top-k written straight into this rank's chunk -> ONE all-gather -> merge kernel over the gathered
        chunks.  events: optional list that receives four CUDA events (start, local top-k done, gathered, merged)."""
        nq = int(q.shape[0])
        if self.world == 1:
            return self.local.topk_device(q, k)
        mine, gathered = self._buffers(nq, k)
        s, i = chunk_views(mine, nq, k)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if events is not None else None
        if ev:
            ev[0].record()
        self.local.topk_device(q, k, out_scores=s, out_idx=i)
        if ev:
            ev[1].record()
        dist.all_gather_into_tensor(gathered, mine, group=self.group)
        if ev:
            ev[2].record()
        out = merge_gathered_topk_device(gathered, nq, self.world, k)
        if ev:
            ev[3].record()
            events[:] = ev
        return out

    def identify(self, q, n_recordings, top_k, n_candidates, events=None):
        _, idx = self.topk_device(q, n_candidates, events=events)
        return vote_device(idx.view(n_recordings, -1), self.row_ids, top_k)


def recording_layout(n_recordings, world):
    """Recordings split over the ranks in contiguous blocks; every rank contributes `per` rows to the gather buffer
    (ranks with fewer recordings pad).  -> (per, [(lo, hi) per rank], gathered-row index of every recording)."""
    per = (n_recordings + world - 1) // world
    bounds = [shard_bounds(n_recordings, r, world) for r in range(world)]
    rowmap = [r * per + i for r, (lo, hi) in enumerate(bounds) for i in range(hi - lo)]
    return per, bounds, rowmap


def gather_recording_results(mine, gathered, rowmap, group=None):
    """mine (2, per, top_k) int32 = this rank's [piece ids | counts]; gathered (world, 2, per, top_k).
    ONE all-gather, then the padding rows are dropped.  -> (ids, counts), each (n_recordings, top_k)."""
    dist.all_gather_into_tensor(gathered.view(-1), mine.view(-1), group=group)
    world, _, per, top_k = gathered.shape
    ids = gathered[:, 0].reshape(world * per, top_k).index_select(0, rowmap)
    cnt = gathered[:, 1].reshape(world * per, top_k).index_select(0, rowmap)
    return ids, cnt


class ReplicatedDB(object):
    """The whole DB on every rank; recordings (each a block of consecutive query rows) split over the ranks."""

    def __init__(self, codes, row_ids, group=None, normalise_in_place=False):
        self.group = group
        self.world = dist.get_world_size(group) if (group is not None or dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.local = EmbeddingDB(codes, idx_base=0, normalise_in_place=normalise_in_place)
        self.row_ids = torch.as_tensor(row_ids).to(torch.int32).to(self.local.device).contiguous()
        self._maps = {}

    def _layout(self, n_recordings, top_k):
        key = (n_recordings, top_k)
        if key not in self._maps:
            per, bounds, rowmap = recording_layout(n_recordings, self.world)
            dev = self.local.device
            self._maps[key] = (per, bounds, torch.as_tensor(rowmap, dtype=torch.int64, device=dev),
                               torch.full((2, per, top_k), -1, dtype=torch.int32, device=dev),          # [ids | counts] of this rank
                               torch.empty((self.world, 2, per, top_k), dtype=torch.int32, device=dev))
        return self._maps[key]

    def identify(self, q, n_recordings, top_k, n_candidates, events=None):
        """q: (n_recordings * windows, 32) on every rank.  -> (piece ids, counts), each (n_recordings, top_k) int32.
        events: optional list that receives three CUDA events (start, local retrieval + vote done, gathered)."""
        if self.world == 1:
            _, idx = self.local.topk_device(q, n_candidates)
            return vote_device(idx.view(n_recordings, -1), self.row_ids, top_k)
        per, bounds, rowmap, mine, gathered = self._layout(n_recordings, top_k)
        win = q.shape[0] // n_recordings
        lo, hi = bounds[self.rank]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if events is not None else None
        if ev:
            ev[0].record()
        if hi > lo:
            _, idx = self.local.topk_device(q[lo * win:hi * win], n_candidates)
            vote_device(idx.view(hi - lo, -1), self.row_ids, top_k, out_ids=mine[0, :hi - lo], out_counts=mine[1, :hi - lo])
        if ev:
            ev[1].record()
        out = gather_recording_results(mine, gathered, rowmap, self.group)
        if ev:
            ev[2].record()
            events[:] = ev
        return out


def gather_topk(scores, idx, group=None):
    """all-gather per-rank (nq,k) lists through the single-buffer chunk layout.  Returns the gathered uint8 buffer
    (world chunks, rank-major); chunk_views(buf, nq, k, rank) decodes one rank's lists."""
    world = dist.get_world_size(group)
    nq, k = scores.shape
    chunk, _ = chunk_layout(nq, k)
    mine = torch.empty(chunk, dtype=torch.uint8, device=scores.device)
    s, i = chunk_views(mine, nq, k)
    s.copy_(scores)
    i.copy_(idx)
    gathered = torch.empty(world * chunk, dtype=torch.uint8, device=scores.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    return gathered
