"""What the host side of a multi-GPU box delivers when every rank uploads at once (context for bench.py's e2e at N>1).

Launch like bench.py (torchrun, one rank per GPU).  Every rank copies CHUNK-byte pieces of a pinned pool to its GPU for about a
second, all ranks inside the same barrier-bracketed window; the pool size decides whether the DMA reads hit the CPU's
last-level cache (a pool of one chunk) or host DRAM (a pool far larger than the cache).  A second pass adds a concurrent
device-to-host stream of the size the hot path downloads.  Rank 0 prints one JSON line.
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    CHUNK = 32 << 20
    D2H = 10 << 20
    dst = torch.empty(CHUNK, dtype=torch.uint8, device="cuda")
    dsrc = torch.zeros(D2H, dtype=torch.uint8, device="cuda")
    hdst = torch.empty(D2H, dtype=torch.uint8).pin_memory()
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {}
    for pool_mb in (32, 128, 1024):
        pool = torch.zeros(pool_mb << 20, dtype=torch.uint8).pin_memory()
        pool[::4096] = 1                                   # touch every page
        nchunk = pool.numel() // CHUNK
        for down in (False, True):
            n = 64                                         # 2 GiB per rank per pass
            for it in range(2):                            # pass 0 warms up
                sync()
                t0 = time.perf_counter()
                for i in range(n):
                    j = i % nchunk
                    with torch.cuda.stream(s_up):
                        dst.copy_(pool[j * CHUNK:(j + 1) * CHUNK], non_blocking=True)
                    if down:
                        with torch.cuda.stream(s_dn):
                            hdst.copy_(dsrc, non_blocking=True)
                torch.cuda.synchronize()
                mine = time.perf_counter() - t0
                sync()
            gbs = n * CHUNK / mine / 1e9
            v = torch.tensor([gbs], dtype=torch.float64, device="cuda")
            allv = [torch.zeros_like(v) for _ in range(world)]
            if world > 1:
                dist.all_gather(allv, v)
            else:
                allv = [v]
            out["pool_%dMB%s" % (pool_mb, "_with_d2h" if down else "")] = [round(float(x[0]), 1) for x in allv]
        del pool
    if rank == 0:
        info = {"cpus_allowed": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}
        for p in ("/sys/fs/cgroup/cpu.max", "/sys/devices/system/node/online"):
            try:
                info[p] = open(p).read().strip()
            except Exception as e:
                info[p] = repr(e)
        try:
            lines = open("/proc/cpuinfo").read().split("\n")
            info["cpu_model"] = [l.split(":", 1)[1].strip() for l in lines if l.startswith("model name")][0]
            info["cache"] = [l.split(":", 1)[1].strip() for l in lines if l.startswith("cache size")][0]
        except Exception:
            pass
        try:
            info["mem_total_gb"] = round(int([l for l in open("/proc/meminfo") if l.startswith("MemTotal")][0].split()[1]) / 1e6, 1)
        except Exception:
            pass
        print(json.dumps({"n_gpus": world, "chunk_mb": CHUNK >> 20, "d2h_chunk_mb": D2H >> 20,
                          "h2d_gbs_per_rank_all_ranks_at_once": out, "host": info}))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
