"""Multi-GPU plumbing (one process per GPU, torch.distributed / NCCL).

The path shards in three places: snippet batches through the encoders (no
collective), CCA covariance sums (one all-reduce of [sums | row count], utils/cca.py), and the
embedding DB: each rank holds a contiguous row shard, writes its local top-k straight into one
contiguous chunk, ONE all-gather moves the chunks, and a merge kernel reads them where they are
(rank-major, no transpose copies); the piece vote runs on the merged candidates.
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
