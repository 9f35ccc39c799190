"""Batch sharding of the loss path across ranks (one process per GPU).

The path shards by independent units: every warped pixel depends only on its own sample b, SSIM windows and
smoothness stencils never cross samples.  Rank r of k takes samples
``[r*B/k, (r+1)*B/k)`` with all S sources and all scales; because the shards are equal-sized, the average of the
per-rank mean losses (and of the gradients of the networks that produced depth and pose) is the global-mean result.
There is NO collective inside the loss.  ``mean_over_ranks`` is the optional all-reduce for logged scalars; the
all-reduce of the DispNetS / PoseNet gradients is the caller's (DDP).

Host-side logic only: works on any device / backend (the world_size-2 gloo test runs it on CPU).
"""
from __future__ import annotations

from typing import Tuple

import torch

from .synthetic import LossInputs


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """samples [lo, hi) of rank `rank`; the batch must divide evenly (equal shards keep mean-of-means exact)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    if B % world:
        raise ValueError(f"batch {B} does not divide over {world} ranks: unequal shards would change the mean")
    n = B // world
    return rank * n, (rank + 1) * n


def shard_inputs(inp: LossInputs, rank: int, world: int) -> LossInputs:
    """The rank's slice of every tensor, keeping the layouts of the path:
    src is source-major (n = src*B + b, models/baseline.py:772-775) and depth holds B target depths followed by
    S*B source depths (baseline.py:905-912), so both are sliced per source block."""
    B, S = inp.B, inp.S
    lo, hi = shard_range(B, rank, world)

    def src_major(t):        # (S*B, ...) -> (S*b, ...)
        return torch.cat([t[k * B + lo:k * B + hi] for k in range(S)], 0)

    def depth(t):            # (B*(1+S), ...) -> (b*(1+S), ...)
        return torch.cat([t[lo:hi]] + [t[B * (k + 1) + lo:B * (k + 1) + hi] for k in range(S)], 0)

    return LossInputs([t[lo:hi] for t in inp.tgt_pyr], [src_major(t) for t in inp.src_pyr], [depth(t) for t in inp.depth_pyr],
                      inp.poses[lo:hi], inp.K[lo:hi], inp.K_ms[lo:hi])


def unshard_depth_grads(parts, B: int, S: int):
    """Inverse of the depth slicing for gradients gathered from all ranks: parts[r] is rank r's (b*(1+S),h,w)."""
    world = len(parts)
    b = B // world
    blocks = []
    for blk in range(1 + S):
        blocks += [p[blk * b:(blk + 1) * b] for p in parts]
    return torch.cat(blocks, 0)


def mean_over_ranks(t: torch.Tensor) -> torch.Tensor:
    """Average a (loss) tensor over the ranks of the default process group; identity without one."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    out = t.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out / dist.get_world_size()
