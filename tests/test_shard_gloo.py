"""N>1 path on CPU: world_size-2 gloo run of the sharding + statistics gather (no collective on the data path).
The per-unit compute is the CPU oracle here (the CUDA path needs a device); the point is that the union of the
ranks' outputs is independent of the number of ranks."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import common

N_SEQ = 4


def unit_digest(seq):
    import orc
    from pysdyn import shard
    W, H, _, nf, ini, mn = common.CONFIGS["small"]
    E = orc.Extractor(nf, 1.2, 8, ini, mn)
    k, d = E(common.frame("small", 0, seq=seq))
    return shard.result_hash(k, d), len(k)


def worker(rank, world, port, q):
    import torch.distributed as dist
    from pysdyn import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_units(N_SEQ, rank, world)
    vec = np.zeros(2 * N_SEQ)
    for s in mine:
        h, n = unit_digest(s)
        vec[2 * s], vec[2 * s + 1] = h, n
    g = shard.gather_stats(vec)                 # [world, 2*N_SEQ]; every unit is non-zero on exactly one rank
    merged = g.sum(0)
    owners = (g[:, ::2] != 0).sum(0)
    if rank == 0:
        q.put((merged.tolist(), owners.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_two_rank_sharding_is_rank_count_invariant():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    merged, owners = q.get(timeout=240)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert owners == [1] * N_SEQ                             # every sequence processed by exactly one rank
    single = []
    for s in range(N_SEQ):
        h, n = unit_digest(s)
        single += [float(h), float(n)]
    assert merged == single                                  # same per-sequence digests as a 1-rank run


def test_shard_units_partition():
    from pysdyn import shard
    for world in (1, 2, 4, 8):
        seen = sorted(u for r in range(world) for u in shard.shard_units(13, r, world))
        assert seen == list(range(13))


def region_worker(rank, world, port, q):
    """bench.py's timed regions: each rank measures its own first region, the repetition count must still be the same."""
    import torch.distributed as dist
    from pysdyn import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = shard.region_count(0.011 * (1 + 3 * rank), 0.5, 5, 60)       # rank 1 is 4x slower than rank 0
    for _ in range(n):                                                  # a different n per rank would deadlock here
        dist.barrier()
    g = shard.gather_stats([float(n)])
    if rank == 0:
        q.put(g[:, 0].tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_region_count_is_agreed_over_ranks():
    from pysdyn import shard
    assert shard.region_count(0.011, 0.5, 5, 60) == 46 and shard.region_count(10.0, 0.5, 5, 60) == 5
    assert shard.region_count(1e-6, 0.5, 1, 60) == 60 and shard.region_count(0.1, 0.0, 1, 60) == 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=region_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    counts = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert counts == [12.0, 12.0]                                        # ceil(0.5 / 0.044), from the slower rank
