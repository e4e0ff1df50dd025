"""Multi-GPU plumbing: replicas only.

Independent replicas are the unit of parallelism: rank k owns a contiguous block of the global
replica range and steps it without talking to anyone.  The only collectives are *outside* the
step -- gathering per-replica statistics (rewards, KPIs) and agreeing on timings -- over
torch.distributed (NCCL over NVLink/NVSwitch on GPUs, gloo in CPU tests).  Counter-based draws are
keyed by the *global* replica index, so results do not depend on how replicas are sharded.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_replicas(total: int, world: int, rank: int):
    """Contiguous block of `total` replicas owned by `rank`: returns (count, first global index)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return count, first


def gather_replica_values(local: torch.Tensor, total: int = None) -> torch.Tensor:
    """All-gather a per-replica vector (or [R_local, k] matrix) into global replica order.
    Ranks may own different counts; every rank receives the full result."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.clone()
    world = dist.get_world_size()
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    counts[dist.get_rank()] = local.shape[0]
    dist.all_reduce(counts)
    width = int(counts.max())
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.cat([p[: int(c)] for p, c in zip(parts, counts.tolist())], dim=0)
    if total is not None and out.shape[0] != total:
        raise RuntimeError(f"gathered {out.shape[0]} replicas, expected {total}")
    return out


def max_over_ranks(values) -> list:
    """Element-wise maximum of a few floats over all ranks (device timings are reported as the
    slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
