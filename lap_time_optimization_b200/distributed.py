"""Multi-GPU: candidates are independent, so rows of alphas[B, n_alpha] are sharded contiguously over
ranks with no data-path collective; the only exchange is one all-gather of each rank's top-k
(lap f64, global index i64) -- 16*k bytes per rank -- after which every rank holds the identical
global top-k that seeds the COBYLA starts (trajectory_bayesian_nonlinear.py:253-257).

Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in CPU tests (the
merge itself is backend-agnostic tensor code)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Rows [lo, hi) of rank `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_topk(laps, idx, k):
    """Merge candidate lists by (lap, index) ascending -- stable, ties keep the lower global index.
    laps: float64 [M], idx: int64 [M]; entries with idx < 0 are padding."""
    laps = torch.where(idx < 0, torch.full_like(laps, float("inf")), laps)
    laps = torch.where(torch.isnan(laps), torch.full_like(laps, float("inf")), laps)
    # two stable sorts = lexicographic (lap, idx)
    o1 = torch.sort(idx, stable=True).indices
    o2 = torch.sort(laps[o1], stable=True).indices
    order = o1[o2][:k]
    return laps[order], idx[order]


def allgather_topk(local_laps, local_idx, k, group=None, merge=merge_topk):
    """All-gather each rank's k best and merge; every rank returns the same (laps[k], idx[k]).
    `merge(laps, idx, k)` defaults to the torch implementation above; on the GPU pass
    `LapTimeEvaluator.merge_topk_device` (ltk_topk_pairs kernel)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return merge(local_laps, local_idx, k)
    # one collective: pack the f64 laps' bits and the i64 indices into a single int64 buffer
    packed = torch.cat([local_laps.contiguous().view(torch.int64), local_idx.contiguous()])
    gathered = torch.empty(world * packed.numel(), dtype=torch.int64, device=packed.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    g = gathered.view(world, 2, -1)
    laps = g[:, 0, :].contiguous().view(torch.float64).reshape(-1)
    idx = g[:, 1, :].reshape(-1).contiguous()
    return merge(laps, idx, k)


class PackedTopkGather:
    """The cross-rank step as `finish` hook of `LapTimeEvaluator.stream_populations` / `run_resident`: ONE all-gather of the
    rank's packed top-k list (int64 [2k]: lap bit patterns, global indices -- written in place by the sweep kernel's
    epilogue) and ONE merge launch on the gathered buffer (`ltk_topk_gathered`).  The unpacked route (`allgather_topk`)
    needs a concatenation before and two re-packing copies after the collective: five small launches per population on
    the communication stream instead of two."""

    packed = True

    def __init__(self, evaluator, k, group=None):
        self.ev, self.k, self.group = evaluator, int(k), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def __call__(self, packed):
        if self.world == 1:
            return packed[:self.k].view(torch.float64), packed[self.k:]
        gathered = torch.empty(self.world * packed.numel(), dtype=torch.int64, device=packed.device)
        dist.all_gather_into_tensor(gathered, packed, group=self.group)
        return self.ev.merge_gathered_device(gathered, self.world, packed.numel() // 2, self.k)


def sharded_population_topk(evaluator, local_alphas, index_base, k, group=None):
    """Score this rank's shard on its GPU, local top-k with global indices, all-gather, merge.
    Returns (local_laps [B_local] CUDA, best_laps[k], best_idx[k])."""
    d_lap, _best, _idx, packed = evaluator.lap_times_topk_device(local_alphas, k=k, index_base=index_base, packed=True)
    g_best, g_idx = PackedTopkGather(evaluator, k, group)(packed)
    return d_lap, g_best, g_idx
