"""Multi-GPU sharding of the front end: the path partitions by sequence (or frame) with NO exchange step
(SURVEY §8e), so every rank processes its own units and the only collective is an all_gather of a small
statistics vector at the end of a run (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import hashlib

import numpy as np


def shard_units(n_units, rank, world):
    """Unit (sequence or frame) s is owned by rank s mod world."""
    return list(range(rank, n_units, world))


def result_hash(*arrays):
    """Order-sensitive 64-bit digest of a frame's outputs, for device-count-invariance checks."""
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return int.from_bytes(h.digest()[:7], "little")          # 56 bits: exact in a float64 / int64 tensor


def gather_stats(vec):
    """all_gather of a 1-D float64 vector; returns [world, len] on every rank (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(vec, np.float64))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t.numpy()[None, :].copy()
    if dist.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


def agree_max(x):
    """max over ranks of a scalar — the same value on every rank (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    v = torch.tensor([float(x)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        v = v.cuda()
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v[0])


def region_count(first_region_seconds, min_seconds, lo=1, hi=60):
    """How many times a timed region is repeated so that it covers min_seconds.  The count comes from a rank's own clock, so
    it is agreed over the ranks (max of the first region's time) before use: every region contains barriers, and ranks that
    ran different numbers of regions would issue different numbers of collectives and hang."""
    t = agree_max(first_region_seconds)
    n = int(np.ceil(min_seconds / max(t, 1e-9))) if min_seconds > 0 else lo
    return int(min(max(lo, n), hi))
