"""Multi-GPU plumbing (one process per GPU, torch.distributed / NCCL).

The path shards in three places: snippet batches through the encoders (no
collective), CCA covariance sums (one all-reduce, utils/cca.py), and the embedding DB: each rank
holds a contiguous row shard, computes its local top-k, the (score, index) lists are all-gathered
and merged on every rank, and the piece vote runs on the merged candidates.
"""
import os

import torch
import torch.distributed as dist

from .retrieval import EmbeddingDB, merge_topk_device, vote_device


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

    def __init__(self, local_codes, idx_base, row_ids_global=None, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if (group is not None or dist.is_initialized()) else 1
        self.local = EmbeddingDB(local_codes, idx_base=idx_base)
        self.row_ids = None
        if row_ids_global is not None:
            self.row_ids = torch.as_tensor(row_ids_global).to(torch.int32).to(self.local.device).contiguous()

    def topk_device(self, q, k):
        s, i = self.local.topk_device(q, k)
        if self.world == 1:
            return s, i
        gs, gi = gather_topk(s, i, self.group)
        return merge_topk_device(gs, gi, self.world, k)

    def identify(self, q, n_recordings, top_k, n_candidates):
        _, idx = self.topk_device(q, n_candidates)
        return vote_device(idx.view(n_recordings, -1), self.row_ids, top_k)


def gather_topk(scores, idx, group=None):
    """all-gather per-rank (nq,k) lists -> (nq, world*k) with lists laid out rank-major per query."""
    world = dist.get_world_size(group)
    nq, k = scores.shape
    gs = torch.empty((world * nq, k), dtype=scores.dtype, device=scores.device)
    gi = torch.empty((world * nq, k), dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(gs, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(gi, idx.contiguous(), group=group)
    gs, gi = gs.view(world, nq, k), gi.view(world, nq, k)
    return gs.permute(1, 0, 2).reshape(nq, world * k).contiguous(), gi.permute(1, 0, 2).reshape(nq, world * k).contiguous()
