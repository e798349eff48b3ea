"""Multi-GPU plumbing: frame-pairs are independent LM problems (the reference already solves every batch entry
separately: AtA[b] per b, utils.cu:368-380; per-pair lambda, bundlenet.py:243-249), so the batch is sharded
contiguously over ranks with NO data-path collective; the only exchange is one all-gather of the solved
(R,T,W) = [nb_local, 9+3+K] floats per rank after the last iteration.
One process per GPU, torch.distributed (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as td

Tensor = torch.Tensor


def shard_range(nb: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of nb pairs; the first nb % world ranks get one extra pair."""
    base, extra = divmod(nb, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_solution(R: Tensor, T: Tensor, W: Optional[Tensor]) -> Tensor:
    nb = R.shape[0]
    parts = [R.reshape(nb, 9), T.reshape(nb, 3)]
    if W is not None:
        parts.append(W.reshape(nb, -1))
    return torch.cat(parts, dim=1).contiguous()


def unpack_solution(x: Tensor, K: int):
    nb = x.shape[0]
    R = x[:, :9].reshape(nb, 3, 3); T = x[:, 9:12].reshape(nb, 3, 1)
    W = x[:, 12:12 + K].reshape(nb, K, 1) if K > 0 else None
    return R, T, W


def all_gather_solution(R: Tensor, T: Tensor, W: Optional[Tensor], group=None, counts: Optional[Sequence[int]] = None):
    """One collective: gather every rank's [nb_local, 12+K] block.  `counts` (pairs per rank) is needed only for
    ragged shards; equal shards use all_gather_into_tensor."""
    K = 0 if W is None else W.shape[1]
    mine = pack_solution(R, T, W)
    world = td.get_world_size(group)
    if counts is None or len(set(counts)) == 1:
        out = torch.empty(world * mine.shape[0], mine.shape[1], device=mine.device, dtype=mine.dtype)
        td.all_gather_into_tensor(out, mine, group=group)
    else:
        mx = max(counts)                       # ragged shards: pad to the largest, gather once, drop the padding
        padded = torch.zeros(mx, mine.shape[1], device=mine.device, dtype=mine.dtype)
        padded[:mine.shape[0]] = mine
        out = torch.empty(world * mx, mine.shape[1], device=mine.device, dtype=mine.dtype)
        td.all_gather_into_tensor(out, padded, group=group)
        out = torch.cat([out[r * mx: r * mx + c] for r, c in enumerate(counts)], dim=0)
    return unpack_solution(out, K)


def solve_sharded(solve_fn: Callable, shard_inputs: Callable[[int, int], tuple], nb: int, group=None):
    """Run `solve_fn(*shard_inputs(lo, hi)) -> (R,T,W)` on this rank's pairs [lo,hi) and all-gather the result.
    `solve_fn` is banet_b200.ops.lm_run-shaped in production; tests inject the CPU oracle over gloo."""
    rank, world = td.get_rank(group), td.get_world_size(group)
    lo, hi = shard_range(nb, rank, world)
    R, T, W = solve_fn(*shard_inputs(lo, hi))
    counts = [shard_range(nb, r, world)[1] - shard_range(nb, r, world)[0] for r in range(world)]
    return all_gather_solution(R, T, W, group=group, counts=counts)
