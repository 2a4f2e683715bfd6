#!/usr/bin/env python3
"""Headline benchmark (BASELINE.json): frames/sec of the tracking front end — ORB extract + SearchByProjection
(frame and local-map) + dynamic mask — on synthetic KITTI-shaped 1241x376 frames, 2000 features, 8 levels,
and the fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl sdyn|reference]

One "step" = one pass of the hot path over one batch of synthetic frames per GPU.  Rank 0 prints ONE JSON
line.  DESIGN.md §Measurement defines value / e2e / roofline / cpu_baseline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "slam-dynamic_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (W, H, nrect, nfeatures, iniTh, minTh, config_id)
    "kitti": (1241, 376, 160, 2000, 12, 7, 0),      # Examples/Stereo/KITTI04-12.yaml
    "tum": (640, 480, 120, 1000, 20, 7, 1),         # Examples/RGB-D/TUM3.yaml
    "4k": (3840, 2160, 2800, 8000, 20, 7, 4),       # stress config
}
RGBD_CTOR = {"tum"}  # workloads tracked with RGB-D-constructor semantics (rgbd_split: static keypoints + re-admitted ones, Frame.cc:297-403)
POOL = 256          # distinct frames per GPU (frame_idx 0..255)
NLEVELS, SCALE = 8, 1.2
N_MAP = 3000        # local-map points per frame (2000-5000)
REF_STRIDE = 1024   # capacity for the reference frame's in-box keypoints
POPC_PEAK = 4.65e12  # __popc results per second: 148 SMs x 16 / clk x 1.965 GHz (replaced by the tools/popc_probe measurement when committed)
try:
    POPC_PEAK = json.load(open(os.path.join(ROOT, "profiles", "r02_popc_probe.json")))["popc_gops"] * 1e9
except Exception:
    pass
METRIC = "frames/sec ORB extract+match+dyn-mask @KITTI 1241x376 2k feats; % HBM roofline"


def csrc_sha():
    """Digest of the kernel sources: the committed ncu figures (profiles/*_ncu_traffic.json) are stamped with it and ignored
    when the kernels have changed since the capture."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "slam-dynamic_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cpp", ".h")):
            h.update(name.encode()); h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def csrc_file_sha():
    """Per-file digests of the kernel sources (the ncu figures carry them too: a kernel's figures stay valid while ITS files do)."""
    import hashlib
    d = os.path.join(ROOT, "slam-dynamic_b200", "csrc")
    return {name: hashlib.sha256(open(os.path.join(d, name), "rb").read()).hexdigest()[:16]
            for name in sorted(os.listdir(d)) if name.endswith((".cu", ".cpp", ".h"))}


# the source files a kernel's instruction count and traffic depend on (its own file + the headers it is built from)
KERNEL_FILES = {
    "k_fast": ["k_fast.cu", "sdyn_internal.h", "tma.h"], "k_blur": ["k_blur.cu", "sdyn_internal.h", "tma.h"],
    "k_resize": ["k_pyramid.cu", "sdyn_internal.h", "tma.h"], "k_level0": ["k_pyramid.cu", "sdyn_internal.h"],
    "k_octree": ["k_octree.cu", "sdyn_internal.h"], "k_orient_describe": ["k_describe.cu", "sdyn_internal.h", "tma.h"],
    "k_match_candidates": ["k_match.cu", "match_internal.h", "sdyn_internal.h"], "k_box_stage": ["k_track.cu", "match_internal.h", "sdyn_internal.h"],
}


def load_ncu_figures(kernel=None):
    """(figures per kernel, note).  Newest profiles/rNN_ncu_traffic.json whose csrc_sha matches the tree; failing that, the
    figures of `kernel` are still accepted when every source file that kernel is built from is unchanged since the capture."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")), reverse=True):
        try:
            tj = json.load(open(path))
        except Exception:
            continue
        if tj.get("csrc_sha") == csrc_sha():
            return tj, os.path.basename(path)
        then, now = tj.get("file_sha") or {}, csrc_file_sha()
        files = KERNEL_FILES.get(kernel)
        if files and all(then.get(f) is not None and then.get(f) == now.get(f) for f in files):
            return tj, "%s (captured for csrc %s; %s unchanged since)" % (os.path.basename(path), tj.get("csrc_sha"), ", ".join(files))
        return None, "%s is stale (captured for csrc %s, tree is %s)" % (os.path.basename(path), tj.get("csrc_sha"), csrc_sha())
    return None, "no ncu capture committed"


def level_sizes(W, H):
    s, out = np.float32(1.0), []
    for _ in range(NLEVELS):
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(W) * inv)), int(np.rint(np.float32(H) * inv))))
        s = np.float32(float(s) * float(np.float32(SCALE)))
    return out


def alg_bytes_extract(W, H, nkp):
    """input read once + every bordered level written once + keypoints and descriptors."""
    return W * H + sum((w + 38) * (h + 38) for w, h in level_sizes(W, H)) + nkp * (32 + 28)


def stage_alg_bytes(W, H, nkp, evals, queries):
    """Per-stage (unfused) algorithmic bytes per frame, SURVEY §8(d) secondary accounting."""
    lv = level_sizes(W, H)
    px = sum(w * h for w, h in lv)
    bordered = sum((w + 38) * (h + 38) for w, h in lv)
    return {
        "level0": W * H + (lv[0][0] + 38) * (lv[0][1] + 38),            # read input, write bordered level 0
        "pyramid": sum(w * h for w, h in lv[:-1]) + bordered - (lv[0][0] + 38) * (lv[0][1] + 38),   # resize chain
        "fast": px,                                                      # every level pixel read once
        "blur": 2 * px,
        "describe": nkp * (749 + 512 + 60),
        "octree": nkp * 8,
        "match": 32 * evals + 56 * queries + 32 * nkp,
        "candidates": 32 * evals + 56 * queries + 32 * nkp,
        "dynamic": 9 * nkp + 32 * 64,
    }


def seq_seed(cfg, rank):
    return 1000 * WORKLOADS[cfg][6] + 100000 * rank + 7


def pool_offsets(i):
    """Camera offset of pool frame i.  Periodic in POOL (frame POOL is frame 0 again), consecutive frames shift by <= 8 px:
    every batch slot can follow the pool round and round as ONE sequence whose LastFrame is always its previous frame."""
    j = i % POOL
    m = j % 32
    return 3 * (m if m <= 16 else 32 - m), (j * 5) % 8


def pool_time(i):
    m = (i % POOL) % 32
    return m if m <= 16 else 32 - m


def make_frames(cfg, rank, first, count, disparity=0):
    """Frames first..first+count-1 (mod POOL) of this rank's cyclic sequence; disparity > 0 renders the right view of a
    rectified stereo rig (content shifted, its own sensor noise)."""
    import pysdyn
    import scenario
    W, H, nrect, _, _, _, cid = WORKLOADS[cfg]
    frames = np.empty((count, H, W), np.uint8)
    nthreads = min(os.cpu_count() or 1, 16)

    def work(t):
        for j in range(t, count, nthreads):
            i = (first + j) % POOL
            ox, oy = pool_offsets(i)
            pysdyn.synth_frame(seq_seed(cfg, rank), 1000 * cid + 100000 * rank + i + (1 << 20) + (disparity << 22), W, H, nrect,
                               ox + disparity, oy, pool_time(i), out=frames[j])

    th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [t.start() for t in th]
    [t.join() for t in th]
    return frames


def bind_to_gpu_numa_node(local):
    """Pin this rank's host thread(s) and its future (pinned) allocations to the NUMA node its GPU hangs off: the H2D path
    of N ranks otherwise crosses the inter-socket link for half of them and they all draw on node 0's memory controllers
    (round 1: GPUs 0-3 saturated near 100 GB/s aggregate).  Best effort; returns what was done for the result line."""
    info = {"gpu_node": None, "cpus": None, "mempolicy": None}
    try:
        bus = subprocess.check_output(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                      text=True, stderr=subprocess.DEVNULL).strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        info["gpu_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = "%d cpus of node %d" % (len(allowed), node)
        else:
            info["cpus"] = "node %d has no cpu in this process's affinity mask (%d allowed)" % (node, len(os.sched_getaffinity(0)))
        # set_mempolicy(MPOL_PREFERRED, {node}): pinned buffers allocated from here on come from the GPU's node
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))
        info["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy failed (errno %d)" % ctypes.get_errno()
    except Exception as e:
        info["error"] = repr(e)
    return info


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 20 ms by a process started BEFORE the warm-up (nvidia-smi needs
    ~0.1 s to produce its first line) and killed after the last timed region; only samples whose host time stamp falls
    between the start of the device-timed region and the end of the end-to-end regions are reported (the GPU is under
    load throughout: device pass, per-stage pass, end-to-end passes)."""

    def __init__(self, gpu_index, period_ms=20):
        import threading
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        self.lines = []
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q,
                                       "--format=csv,noheader,nounits", "-lms", str(period_ms)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append((time.perf_counter(), line))

    def wait_started(self, timeout=3.0):
        t_end = time.perf_counter() + timeout
        while self.p and not self.lines and time.perf_counter() < t_end:
            time.sleep(0.005)

    def stop(self, t0, t1):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            pass
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in list(self.lines):
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons),
                "window": "device-timed region .. end of the end-to-end regions (GPU under load throughout)"}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU implementation of the path.  The reference itself cannot be built here (needs
# OpenCV C++ >= 3.4 + contrib, Eigen, Pangolin, PCL, Boost), so this runs its restatement in oracle/ ("port").
# ------------------------------------------------------------------------------------------------------------
def cpu_prepare(cfg, nframes, threads=1):
    """Inputs of the CPU arm: the first `nframes` frame inputs of rank 0's pool (frame i tracked against frame i-1),
    extracted once (frame-parallel) to build the track queries."""
    import orc
    import scenario
    W, H, nrect, nf, ini, mn, _ = WORKLOADS[cfg]
    frames = make_frames(cfg, 0, -1, nframes + 1)              # frame -1 (= POOL-1) is the LastFrame of frame 0
    kd = [None] * len(frames)

    def work(t):
        ex = orc.Extractor(nf, SCALE, NLEVELS, ini, mn)
        for i in range(t, len(frames), threads):
            kd[i] = ex(frames[i])

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    cap = nf + 200
    arrays = scenario.build_track_batch(kd, seq_seed(cfg, 0), 0, W, H, nrect, NLEVELS, cap, N_MAP, REF_STRIDE,
                                        n_map=N_MAP, seed=3, offsets=pool_offsets, time=pool_time, frustum=True,
                                        scale=orc.Extractor(nf, SCALE, NLEVELS, ini, mn).scale)
    if cfg in RGBD_CTOR:                   # LastFrame of input f = the list frame f-1 was tracked with (input f-1's own product)
        import oracle_track
        orders = [oracle_track.rgbd_order(kd[f + 1][0], kd[f + 1][1], arrays, f)[0] for f in range(len(kd) - 2)]
        for f, o in enumerate(orders):
            scenario.reorder_last(arrays, f + 1, o)
    return frames[1:], arrays, scenario.track_params(W, H), cap


def ref_available():
    """oracle/_ref/libref.so — the reference's OWN ORBextractor.cc / Frame.cc / ORBmatcher.cc compiled here — can be loaded."""
    try:
        import ref
        return ref.available()
    except Exception:
        return False


def ref_prebuild(cfg, frames_kd, arrays, params):
    """Per frame input of the pool, what persists across frames in the reference: LastFrame (a Frame with its MapPoints) and the
    local map (MapPoint objects with world position, normal, distances, descriptor).  Built once, outside the timed region."""
    import ref
    import scenario
    W, H, _, nf, ini, mn, _ = WORKLOADS[cfg]
    cam = scenario.KITTI_CAM
    camt = (cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["bf"], cam["bf"] / cam["fx"])
    RE = ref.Extractor(nf, SCALE, NLEVELS, ini, mn)
    pre = []
    for f in range(len(arrays["n_last"])):
        n0 = int(arrays["n_last"][f])
        lp = arrays["last_points"][f, :n0]
        last = ref.Frame.from_arrays(RE, arrays["last_keys"][f, :n0], np.zeros((n0, 32), np.uint8), (0.0, 0.0, float(W), float(H)), camt,
                                     tcw=arrays["poses"][f, 12:])
        lpts = ref.Points(lp["world"], lp["desc"], present=lp["has_mp"], nobs=lp["obs_positive"].astype(np.int32))
        last.set_points(lpts, lp["outlier"])
        nm = int(arrays["n_map"][f])
        mt, fl = arrays["map_table"][f, :nm], arrays["map_flags"][f, :nm]
        mpts = ref.Points(mt["world"], mt["desc"], present=((fl & 4) == 0).astype(np.uint8), normal=mt["normal"], min_dist=mt["min_distance"],
                          max_dist=mt["max_distance"], nobs=((fl & 2) != 0).astype(np.int32), bad=((fl & 1) != 0).astype(np.uint8))
        pre.append((last, lpts, mpts))
    return RE, pre, camt


def cpu_run(cfg, frames, arrays, params, cap, threads, seconds=None, count=None, first=0, pre=None):
    """Runs the full per-frame hot path (extract + 2 searches + dynamic mask) on the CPU, frame-parallel over `threads` host
    threads, for `seconds` or for `count` frames.  pre = ref_prebuild(...): extraction, grid, isInFrustum and both searches run
    THE REFERENCE'S OWN translation units (oracle/_ref/libref.so); otherwise the oracle port.  Returns (fps, frames_done)."""
    import orc
    import oracle_track
    W, H, _, nf, ini, mn, _ = WORKLOADS[cfg]
    split = cfg in RGBD_CTOR
    done = [0] * threads
    t0 = time.perf_counter()

    def work(t):
        if pre is not None:
            import ref
            RE = ref.Extractor(nf, SCALE, NLEVELS, ini, mn)
        else:
            ex = orc.Extractor(nf, SCALE, NLEVELS, ini, mn)
        i = t
        while True:
            if seconds is not None and time.perf_counter() - t0 >= seconds:
                break
            if count is not None and i >= count:
                break
            f = (first + i) % len(frames)
            if pre is not None:
                _, prebuilt, camt = pre
                last, lpts, mpts = prebuilt[f]
                k, d = RE(frames[f])                                                  # src/ORBextractor.cc
                if split:                 # the RGB-D constructor's list: firstSeparate + Separate + UpdateFrame (the port, small)
                    order, _, in_box, _ = oracle_track.rgbd_order(k, d, arrays, f)
                    cid = np.where(in_box, np.arange(len(k)), -1).astype(np.int32)
                    k = k[order].copy(); k["class_id"] = cid[order]; d = np.ascontiguousarray(d[order])
                cur = ref.Frame.from_arrays(RE, k, d, (0.0, 0.0, float(W), float(H)), camt, tcw=arrays["poses"][f, :12])   # Frame.cc:463-478
                ref.search_by_projection_frame(cur, last, params["th_frame"], bool(params["mono"]), 0.9, bool(params["check_orientation"]))
                ref.points_in_frustum(cur, mpts, 0.5)                                 # Frame.cc:677-733 for every local-map point
                ref.search_by_projection_map(cur, mpts, params["th_map"], params["nnratio_map"])
                if not split:
                    oracle_track.dyn_mask(k, d, arrays, f)                            # per-box BFMatcher + classifyF: the port (small)
                cur.close()
            else:
                k, d = ex(frames[f])
                (oracle_track.track_frame_rgbd if split else oracle_track.track_frame)(k, d, ex.scale, W, H, arrays, f, params, cap)
            done[t] += 1
            i += threads

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    return sum(done) / dt, sum(done)


REF_KIND_NOTE = ("extraction, grid, isInFrustum and both searches run the reference's own src/ORBextractor.cc, src/Frame.cc and "
                 "src/ORBmatcher.cc (oracle/_ref/libref.so, compiled from /root/reference over a stand-in OpenCV whose image primitives "
                 "are the cv2-pinned restatements); the per-box BFMatcher + classifyF stage runs the oracle port")


def workload_config(cfg, B, nctx=None):
    """The `config` object of the result line — identical in both arms (same workload string, pool, frames per step)."""
    W, H, _, nf, ini, mn, _ = WORKLOADS[cfg]
    c = {"workload": "%s %dx%d nfeatures=%d levels=%d scale=%.1f iniTh=%d minTh=%d" % (cfg, W, H, nf, NLEVELS, SCALE, ini, mn),
         "frames_per_step_per_gpu": B, "map_points_per_frame": N_MAP, "pool_frames": POOL,
         "stages": "extract + firstSeparate/tail split + Separate + UpdateFrame (RGB-D constructor) + SearchByProjection(cur,last) + "
                   "isInFrustum + SearchByProjection(F,map) on the static + re-admitted keypoints" if cfg in RGBD_CTOR else
                   "extract + SearchByProjection(cur,last) + isInFrustum + SearchByProjection(F,map) + dynamic mask",
         "sharding": "one set of sequences per rank, no data-path collective; NCCL all_gather of run statistics only"}
    return c


def run_reference(args):
    """The reference's CPU implementation of the path on all host threads: the C++ oracle ("port" — the reference itself
    needs OpenCV C++/Eigen/Pangolin/PCL to build; oracle/_ref holds its hot-path TUs compiled over a stand-in OpenCV and
    tests/test_oracle_ref.py shows the port equals them bit for bit).  Same pool, same frames per step as the sdyn arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.workload
    threads = os.cpu_count() or 1
    B = args.batch
    frames, arrays, params, cap = cpu_prepare(cfg, POOL, threads)
    pre = ref_prebuild(cfg, None, arrays, params) if ref_available() else None
    if args.warmup > 0:
        cpu_run(cfg, frames, arrays, params, cap, threads, count=threads, pre=pre)
    t0 = time.perf_counter()
    total = 0
    for s in range(args.steps):
        total += cpu_run(cfg, frames, arrays, params, cap, threads, count=B, first=(s * B) % POOL, pre=pre)[1]
    dt = time.perf_counter() - t0
    fps = total / dt
    emit(({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(cfg, B),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "reference" if pre is not None else "port",
                         "sample": "%d steps x %d frames of the %d-frame pool, frame-parallel; %s" % (
                             args.steps, B, POOL, REF_KIND_NOTE if pre is not None else
                             "C++ oracle (restated reference CPU path; equals the reference's own code bit for bit, tests/test_oracle_ref.py)")},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def big_vocabulary(k=10, L=6, seed=1):
    """ORBvoc-shaped synthetic tree (k = 10, L = 6: 1 111 111 nodes), generated level by level with numpy."""
    r = np.random.default_rng(seed)
    parents, leafs, descs, weights = [np.zeros(1, np.int32)], [np.zeros(1, np.uint8)], [np.zeros((1, 32), np.uint8)], [np.zeros(1)]
    first, count, prev = 0, 1, descs[0]
    for lvl in range(1, L + 1):
        n = count * k
        par = np.repeat(np.arange(first, first + count, dtype=np.int32), k)
        d = np.repeat(prev, k, axis=0)
        for _ in range(max(1, 48 >> lvl)):
            bit = r.integers(0, 256, n)
            d[np.arange(n), bit >> 3] ^= (1 << (bit & 7)).astype(np.uint8)
        parents.append(par); leafs.append(np.full(n, int(lvl == L), np.uint8)); descs.append(d)
        weights.append(r.uniform(0.5, 9.0, n) if lvl == L else np.zeros(n))
        first, count, prev = first + count, n, d
    return np.concatenate(parents), np.concatenate(leafs), np.concatenate(descs), np.concatenate(weights)


def next_rows(pysdyn, torch, cfg, W, H, nf, ini, mn, local, B, steps, rank, dptrs, strides, params, dev_frames, mtab, pitch, host=None):
    """Device timings of the SURVEY §8(f) rows built after the headline path (same parity bar, tests/test_gpu_*.py):
    ComputeStereoMatches per stereo pair and ComputeBoW per frame, with the oracle timed beside them.  Reported next
    to the headline, never part of it."""
    import orc
    import scenario
    out = {}
    Bs = min(B, 32)
    L = pysdyn.Extractor(nf, SCALE, NLEVELS, ini, mn, max_width=W, max_height=H, max_batch=Bs, device=local)
    R = pysdyn.Extractor(nf, SCALE, NLEVELS, ini, mn, max_width=W, max_height=H, max_batch=Bs, device=local)
    pairs = [scenario.stereo_pair(cfg, 300 + i, (4 + i % 7, 12, 28 - i % 5)) for i in range(Bs)]
    dl = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda(); dr = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
    cam = scenario.KITTI_CAM
    mb, mbf = cam["bf"] / cam["fx"], cam["bf"]

    def stereo_step():
        L.extract_batch_device(dl.data_ptr(), Bs, W * H, W, H, W)
        R.extract_batch_device(dr.data_ptr(), Bs, W * H, W, H, W)
        pysdyn.stereo_match_device(L, R, Bs, mb, mbf)

    for _ in range(3):
        stereo_step()
    L.sync(); R.sync()
    L.profile(True)
    t0 = time.perf_counter()
    for _ in range(steps):
        stereo_step()
    L.sync(); R.sync()
    dt = time.perf_counter() - t0
    st = L.profile_read(); L.profile(False)
    ur, dp, kept = pysdyn.stereo_fetch(L, Bs)
    oL, oR = orc.Extractor(nf, SCALE, NLEVELS, ini, mn), orc.Extractor(nf, SCALE, NLEVELS, ini, mn)
    kl, dsl = oL(pairs[0][0]); kr, dsr = oR(pairs[0][1])
    t1 = time.perf_counter()
    our, odp, okept = orc.stereo_matches(oL, oR, kl, dsl, kr, dsr, mb, mbf)
    cpu_ms = (time.perf_counter() - t1) * 1e3
    out["stereo"] = {"pairs_per_s_extract_x2_plus_stereo": Bs * steps / dt,
                     "stereo_match_ms_per_step": st["stereo"][0] / max(st["stereo"][1], 1), "pairs_per_step": Bs,
                     "stereo_points_per_pair": float(kept.mean()), "parity_frame0": bool(np.array_equal(ur[0, :len(our)], our)),
                     "cpu_oracle_ms_per_pair_stereo_match_only": cpu_ms,
                     "reference": "Frame::ComputeStereoMatches, src/Frame.cc:874-1048"}
    # BASELINE config 3: stereo pairs through the whole step — both extractions, ComputeStereoMatches, the two searches
    # with mvuRight gates, dynamic mask — on this rank's sequence (right views rendered with an 11 px disparity)
    if strides[0] <= L.cap:
        NS = 8                                    # the slots advance one frame per step, like the headline run (resident LastFrame)
        dright = torch.from_numpy(make_frames(cfg, rank, 0, Bs + NS, disparity=11)).cuda()
        sp = dict(params); sp["mono"] = 0
        kstep = [0]

        def stereo_track_step():
            i0 = kstep[0] % NS; kstep[0] += 1
            tin = pysdyn.track_inputs(dptrs, i0, strides, sp, map_table=mtab, frame_pitch=pitch)
            pysdyn.track_batch_stereo_device(L, R, Bs, dev_frames[i0].data_ptr(), dright[i0].data_ptr(), W * H, W, H, W, tin, mb, mbf)

        for _ in range(3):
            stereo_track_step()
        L.sync(); R.sync()
        smp = ClockSampler(local); smp.wait_started()
        tc0 = time.perf_counter()
        runs3 = []
        while len(runs3) < 3 or (time.perf_counter() - tc0 < 0.6 and len(runs3) < 40):     # >= 0.5 s under load for the clock record
            t0 = time.perf_counter()
            for _ in range(steps):
                stereo_track_step()
            L.sync(); R.sync()
            runs3.append(time.perf_counter() - t0)
        clk3 = smp.stop(tc0, time.perf_counter())
        clk3["window"] = "the stereo-pair regions"
        dt3 = float(np.median(runs3))
        while kstep[0] % NS in (0, 1):            # report the counts of a step whose LastFrame was the slot's previous frame
            stereo_track_step()
        L.sync(); R.sync()
        ur3, dp3, kept3 = pysdyn.stereo_fetch(L, Bs)
        a3, l3, m3, c3 = pysdyn.track_fetch(L, Bs)
        out["stereo_track_config3"] = {"pairs_per_s": Bs * steps / dt3, "ms_per_step": 1e3 * dt3 / steps, "pairs_per_step": Bs,
                                       "stereo_points_per_pair": float(kept3.mean()), "matches_frame": float(c3[:, 0].mean()),
                                       "matches_map": float(c3[:, 1].mean()), "clocks": clk3,
                                       "regions_ms": [round(1e3 * v, 3) for v in runs3],
                                       "stages": "extract L + extract R + ComputeStereoMatches + SearchByProjection(cur,last) + "
                                                 "SearchByProjection(F,map) + dynamic mask (one context pair, host clock)"}
    # ComputeBoW on the left frames with an ORBvoc-shaped vocabulary (k = 10, L = 6)
    parent, leaf, vdesc, weight = big_vocabulary()
    voc = pysdyn.Vocabulary(parent, leaf, vdesc, weight, 10, 6, device=local)
    for _ in range(3):
        pysdyn.bow_transform_device(L, voc, Bs, 4)
    L.sync(); L.profile(True)
    for _ in range(steps):
        pysdyn.bow_transform_device(L, voc, Bs, 4)
    L.sync()
    sb = L.profile_read(); L.profile(False)
    word, w, node = pysdyn.bow_fetch(L, Bs)
    kps, dsc, cnt = L.fetch(Bs)
    ovoc = orc.Vocabulary(parent, leaf, vdesc, weight, 10, 6)
    t1 = time.perf_counter()
    ref = ovoc.transform(dsc[0, :cnt[0]], 4)
    cpu_ms = (time.perf_counter() - t1) * 1e3
    ms = sb["bow"][0] / max(sb["bow"][1], 1)
    feats = float(cnt.sum())
    out["bow"] = {"ms_per_step": ms, "frames_per_step": Bs, "features_per_s": feats / (ms * 1e-3),
                  "vocabulary": "synthetic k=10 L=6, %d nodes, %d words (%.0f MB of descriptors)" % (voc.nnodes, voc.nwords, voc.nnodes * 32 / 1e6),
                  "alg_gbs_node_descriptors": feats * 60 * 32 / (ms * 1e-3) / 1e9,
                  "parity_frame0": bool(np.array_equal(word[0, :cnt[0]], ref["word"]) and np.array_equal(node[0, :cnt[0]], ref["node"])),
                  "cpu_oracle_ms_per_frame": cpu_ms,
                  "reference": "Frame::ComputeBoW -> DBoW2 transform, src/Frame.cc:803-810"}
    voc.close(); L.close(); R.close()
    # the heavy matcher case: monocular initialisation — the 2 x nFeatures extractor (Tracking.cc:128) and
    # SearchForInitialization with a 100-px window on level 0 (Tracking.cc:1461, ORBmatcher.cc:562-677)
    try:
        ini2 = pysdyn.Extractor(2 * nf, SCALE, NLEVELS, ini, mn, max_width=W, max_height=H, max_batch=1, device=local)
        ia, ib = pairs[0][0], pairs[1][0]
        ka, da = ini2(ia); kb, db = ini2(ib)
        sc = ini2.GetScaleFactors()
        F1 = scenario.frame_view(ka, da, sc, W, H); F2 = scenario.frame_view(kb, db, sc, W, H)
        prevm = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
        mi = pysdyn.Matcher(ini2, 0.9, True)
        for _ in range(3):
            mi.SearchForInitialization(F1, F2, prevm, 100)
        li = []
        for _ in range(20):
            t1 = time.perf_counter(); nmi, _, _ = mi.SearchForInitialization(F1, F2, prevm, 100); li.append(time.perf_counter() - t1)
        ev = mi.last_evals()
        t1 = time.perf_counter(); orc.match_init(F1, F2, prevm, 100, 0.9, True); cpu_i = time.perf_counter() - t1
        out["init_search"] = {"keypoints": [int(len(ka)), int(len(kb))], "level0_queries": int((ka["octave"] == 0).sum()), "matches": int(nmi),
                              "hamming_evals": ev, "ms_per_call": 1e3 * float(np.median(li)), "evals_per_s_through_the_call": ev / float(np.median(li)),
                              "cpu_oracle_ms": 1e3 * cpu_i,
                              "note": "one SearchForInitialization call through the C ABI with host arrays (H2D of both frames, grid, "
                                      "candidates, sequential resolve, D2H inside the call)"}
        ini2.close()
    except Exception as e:
        out["init_search"] = {"error": repr(e)}
    # the reference's own execution model: one frame per call through the drop-in entry points (host buffers in and out)
    try:
        import orc as _orc
        one = pysdyn.Extractor(nf, SCALE, NLEVELS, ini, mn, max_width=W, max_height=H, max_batch=1, device=local)
        img0, img1 = pairs[0][0], pairs[1][0]
        for _ in range(5):
            one(img0)
        lat = []
        for i in range(40):
            t1 = time.perf_counter(); k1_, d1_ = one(img1 if i & 1 else img0); lat.append(time.perf_counter() - t1)
        k0_, d0_ = one(img0); k1_, d1_ = one(img1)
        scale = one.GetScaleFactors()
        cur = scenario.frame_view(k1_, d1_, scale, W, H); last = scenario.frame_view(k0_, d0_, scale, W, H)
        lp = scenario.last_points(k0_, d0_, (0, 0), seed=7)
        mt = pysdyn.Matcher(one, 0.9, True)
        for _ in range(3):
            mt.SearchByProjectionFrame(cur, last, lp, 7.0, False)
        lm = []
        for _ in range(20):
            t1 = time.perf_counter(); mt.SearchByProjectionFrame(cur, last, lp, 7.0, False); lm.append(time.perf_counter() - t1)
        oe = _orc.Extractor(nf, SCALE, NLEVELS, ini, mn)
        t1 = time.perf_counter(); oe(img0); cpu_ex = time.perf_counter() - t1
        t1 = time.perf_counter(); _orc.match_projection_frame(cur, last, lp, 7.0, False, True); cpu_m = time.perf_counter() - t1
        full = None
        if host is not None:
            # the whole per-frame path as ONE call on a one-frame context: image + frame record up, keypoints, descriptors, match
            # indices, lock flags and mask down (the slot's LastFrame is the previous call's frame, resident)
            hptrs, pin_in = host
            cap1 = one.cap
            pinned1 = [pysdyn.PinnedArray(shape, dt) for shape, dt in
                       [((1, cap1), pysdyn.KP_DTYPE), ((1, cap1, 32), np.uint8), ((1,), np.int32), ((1, cap1), np.int32),
                        ((1, cap1), np.uint8), ((1, cap1), np.uint8), ((1, 4), np.int32)]]       # kept alive: they own the memory
            o1 = tuple(a.array for a in pinned1)
            lf = []
            for i in range(45):
                i0 = 20 + i
                tin = pysdyn.track_inputs(hptrs, i0, strides, params, map_table=mtab, frame_pitch=pitch)
                t1 = time.perf_counter(); pysdyn.track_batch_host(one, pin_in.array[i0:i0 + 1], tin, o1); lf.append(time.perf_counter() - t1)
            full = {"ms": 1e3 * float(np.median(lf[5:])), "matches_frame": int(o1[6][0, 0]), "matches_map": int(o1[6][0, 1]),
                    "note": "extract + SearchByProjection(cur,last) + isInFrustum + SearchByProjection(F,map) + dynamic mask for ONE frame "
                            "through sdyn_track_batch (host buffers in and out, synchronous)"}
        out["single_frame_latency"] = {"extract_ms": 1e3 * float(np.median(lat)), "search_by_projection_frame_ms": 1e3 * float(np.median(lm)),
                                       "full_path_one_call": full,
                                       "cpu_oracle_extract_ms": 1e3 * cpu_ex, "cpu_oracle_search_ms": 1e3 * cpu_m,
                                       "note": "ORBextractor::operator() / SearchByProjection(cur,last) one call at a time, host arrays in "
                                               "and out (H2D, kernels, D2H, synchronisation inside the call)"}
        one.close()
    except Exception as e:
        out["single_frame_latency"] = {"error": repr(e)}
    return out


_RESULT_FD = None


def emit(line):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def run_init(args):
    """--workload init: the heavy matcher case as its own line.  A step is one monocular-initialisation attempt of
    the reference (Tracking::MonocularInitialization, Tracking.cc:586-660): the 2 x nFeatures extractor (Tracking.cc:128) on two
    KITTI-shaped frames, then SearchForInitialization with a 100-px window on level 0 (Tracking.cc:1461, ORBmatcher.cc:562-677),
    every call through the C ABI with host arrays in and out — so `value` IS the end-to-end number (host clock around
    synchronous calls).  One GPU; under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    import pysdyn
    import scenario
    import orc
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sdyn path has no CPU fallback")
    torch.cuda.set_device(local)
    W, H, nrect, nf, ini, mn, _ = WORKLOADS["kitti"]
    K, Wm = args.steps, max(args.warmup, 3)
    npairs = 16
    frames = make_frames("kitti", 0, 0, npairs + 1)
    ex = pysdyn.Extractor(2 * nf, SCALE, NLEVELS, ini, mn, max_width=W, max_height=H, max_batch=1, device=local)
    mt = pysdyn.Matcher(ex, 0.9, True)
    sc = ex.GetScaleFactors()
    stats = {"evals": 0, "matches": 0, "queries": 0, "kp": 0}

    def step(i):
        a, b = frames[i % npairs], frames[i % npairs + 1]
        ka, da = ex(a); kb, db = ex(b)
        F1 = scenario.frame_view(ka, da, sc, W, H); F2 = scenario.frame_view(kb, db, sc, W, H)
        prevm = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)           # vbPrevMatched = F1's keypoints (Tracking.cc:606-608)
        n, _, _ = mt.SearchForInitialization(F1, F2, prevm, 100)
        stats["evals"] += mt.last_evals(); stats["matches"] += int(n); stats["queries"] += int((ka["octave"] == 0).sum())
        stats["kp"] += len(ka) + len(kb)

    sampler = ClockSampler(local); sampler.wait_started()
    for i in range(Wm):
        step(i)
    torch.cuda.synchronize()
    l0 = ex.launch_count()
    for k_ in stats:
        stats[k_] = 0
    t_clk0 = time.perf_counter()
    runs, nsteps = [], 0
    while len(runs) < 3 or (time.perf_counter() - t_clk0 < args.min_seconds and len(runs) < 60):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(K):
            step(nsteps + i)
        torch.cuda.synchronize(); runs.append(time.perf_counter() - t0); nsteps += K
    clocks = sampler.stop(t_clk0, time.perf_counter())
    launches = (ex.launch_count() - l0) // len(runs)
    dt = float(np.median(runs))
    fps = 2 * K / dt
    ev_step = stats["evals"] / nsteps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kp_frame = stats["kp"] / (2 * nsteps)
    alg = 2 * alg_bytes_extract(W, H, int(kp_frame)) + ev_step * 64          # two extractions + two 32-byte descriptors per evaluation
    cpu = None
    if args.cpu_seconds > 0:
        oe = orc.Extractor(2 * nf, SCALE, NLEVELS, ini, mn)
        t0 = time.perf_counter(); n_cpu = 0; t_search = 0.0
        while time.perf_counter() - t0 < args.cpu_seconds:
            a, b = frames[n_cpu % npairs], frames[n_cpu % npairs + 1]
            ka, da = oe(a); kb, db = oe(b)
            F1 = scenario.frame_view(ka, da, oe.scale, W, H); F2 = scenario.frame_view(kb, db, oe.scale, W, H)
            t1 = time.perf_counter()
            orc.match_init(F1, F2, np.stack([ka["x"], ka["y"]], 1).astype(np.float32), 100, 0.9, True)
            t_search += time.perf_counter() - t1
            n_cpu += 1
        cpu = {"value": 2 * n_cpu / (time.perf_counter() - t0), "unit": "frames/s", "cores": 1, "kind": "port",
               "search_ms_per_call": 1e3 * t_search / max(n_cpu, 1),
               "sample": "%d initialisation attempts (2 extractions with %d features + SearchForInitialization) on the C++ oracle, 1 thread"
                         % (n_cpu, 2 * nf)}
    h2d = 2 * W * H + int(kp_frame) * (2 * (28 + 32) + 8)
    emit({"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": K, "warmup": Wm, "ms_per_step": 1e3 * dt / K,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
          "config": {"workload": "init: kitti %dx%d, 2 x nFeatures = %d, levels=%d scale=%.1f, SearchForInitialization windowSize=100" % (W, H, 2 * nf, NLEVELS, SCALE),
                     "frames_per_step_per_gpu": 2, "stages": "extract(F1) + extract(F2) + SearchForInitialization, one call each through the C ABI",
                     "l2": "the two frames of a step change every step (16 pairs); calls are synchronous, so caches are cold in the sense the call pattern makes them"},
          "run": {"regions_ms": [round(1e3 * v, 3) for v in runs], "timing": "host clock around K synchronous steps, regions repeated for >= %.1f s; median region" % args.min_seconds},
          "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(2 * kp_frame * 60 + kp_frame * 4),
                  "note": "the calls take and return host arrays: value is already end to end"},
          "gpu_launches": int(launches), "clocks": clocks,
          "roofline": {"kernel": "whole step (latency-bound chain of one-frame kernels)", "bound": "hbm", "achieved": alg / (dt / K) / 1e9, "peak": peak,
                       "unit": "GB/s", "frac": alg / (dt / K) / 1e9 / peak, "traffic": None,
                       "alg_bytes_per_launch": alg},
          "matching": {"hamming_evals_per_step": ev_step, "level0_queries_per_step": stats["queries"] / nsteps, "matches_per_step": stats["matches"] / nsteps,
                       "evals_per_s": ev_step * K / dt, "popc_per_s": 8 * ev_step * K / dt, "popc_peak_per_s": POPC_PEAK,
                       "popc_pipe_frac": 8 * ev_step * K / dt / POPC_PEAK},
          "cpu_baseline": cpu})
    ex.close()


def main():
    # Libraries (NCCL's version banner, torchrun notices) write to fd 1; keep stdout for the result line only.
    global _RESULT_FD, POOL
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sdyn", choices=["sdyn", "reference"])
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS) + ["init"])
    ap.add_argument("--batch", type=int, default=64, help="frames per step per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample length (0 = skip)")
    ap.add_argument("--contexts", type=int, default=6, help="contexts (streams) per GPU taking steps round-robin")
    ap.add_argument("--min-seconds", type=float, default=0.5, help="repeat the K-step timed regions until each arm has run this long")
    ap.add_argument("--next-rows", type=int, default=1, help="also time the SURVEY 8(f) rows (stereo, BoW) on rank 0 at N=1")
    ap.add_argument("--pool", type=int, default=POOL, help="distinct frames per GPU (256 = the survey's; smaller only for smoke tests)")
    args = ap.parse_args()
    POOL = args.pool
    if args.workload == "init":
        if args.impl == "reference":
            raise SystemExit("--workload init has no reference arm: its line carries the CPU oracle as cpu_baseline")
        return run_init(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import pysdyn
    import scenario

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sdyn path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)                     # before any pinned allocation
    if world > 1:
        import datetime
        # a rank that dies (assert, CUDA error) must not leave the others waiting for the default 10 minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))

    cfg = args.workload
    W, H, nrect, nf, ini, mn, _ = WORKLOADS[cfg]
    split = cfg in RGBD_CTOR
    B = args.batch
    K, Wm = args.steps, max(args.warmup, 3)
    assert POOL % B == 0 or B <= POOL

    # ---- inputs: this rank's cyclic sequence (sharded by sequence: no data-path collective) ----------------------
    # Frame input i = (frame i as CurrentFrame, frame i-1 as LastFrame), i in 0..POOL-1, cyclic.  Every batch slot is one
    # sequence that walks the pool round and round, one frame per step, so the keypoints a slot extracted in its previous
    # step ARE its LastFrame (resident on the device, as the reference keeps mLastFrame), and what the reference keeps in
    # MapPoint objects lives in a device-resident table.  A step consumes, per frame: the image, LastFrame's MapPoint ids +
    # flag bytes, the local map's ids + projection records (Frame::isInFrustum's outputs), boxes + the reference frame's
    # in-box keypoints, F21 and the pose pair.
    frames = make_frames(cfg, rank, 0, POOL)
    ex = pysdyn.Extractor(nf, SCALE, NLEVELS, ini, mn, max_width=W, max_height=H, max_batch=B, device=local)
    cap = ex.cap
    kd = []
    for s in range(0, POOL, B):                                 # untimed pre-pass: LastFrame / map / box inputs
        chunk = frames[s:s + B]
        k, d, n = ex.extract_batch(chunk)
        kd += [(k[i, :n[i]].copy(), d[i, :n[i]].copy()) for i in range(len(chunk))]
    # per-keypoint arrays are sized by the data (largest keypoint count of the sequence, rounded up), like the reference's vectors
    last_stride = min(cap, (max(len(k) for k, _ in kd) + 63) // 64 * 64)
    arrays = scenario.build_track_batch([kd[-1]] + kd, seq_seed(cfg, rank), 0, W, H, nrect, NLEVELS, last_stride, N_MAP, REF_STRIDE,
                                        n_map=N_MAP, seed=3, offsets=pool_offsets, time=pool_time, frustum=True,
                                        scale=np.asarray(ex.GetScaleFactors(), np.float32))
    params = scenario.track_params(W, H)
    # the reference frame's in-box keypoint block is sized by the data (largest per-frame total, rounded up)
    ref_stride = min(REF_STRIDE, max(64, (int(arrays["ref_off"].reshape(len(arrays["ref_off"]), -1)[:, -1].max()) + 63) // 64 * 64))
    arrays["ref_desc"] = np.ascontiguousarray(arrays["ref_desc"].reshape(len(arrays["ref_desc"]), REF_STRIDE, 32)[:, :ref_stride])
    arrays["ref_xy"] = np.ascontiguousarray(arrays["ref_xy"].reshape(len(arrays["ref_xy"]), REF_STRIDE, 2)[:, :ref_stride])
    strides = (last_stride, N_MAP, ref_stride)
    table, res = scenario.resident_forms(arrays)
    # local map: ids + one state byte per point; Frame::isInFrustum (the producer of the projection records) runs on the device
    FORMS = pysdyn.FORM_RESIDENT_LAST | pysdyn.FORM_RESIDENT_MAP | pysdyn.FORM_DEVICE_FRUSTUM
    step_arrays = {k: arrays[k] for k in ("n_map", "boxes", "n_boxes", "ref_box", "ref_desc", "ref_xy", "ref_off", "fmat", "poses")}
    step_arrays.update({k: res[k] for k in ("last_ids", "last_flags", "map_ids", "map_flags")})
    # frame-major record pool (one record per frame input), cyclically extended by one batch: any batch is one contiguous run
    pin_pool = None
    layout, pitch = pysdyn.track_record_layout(strides, FORMS)
    pin_pool = pysdyn.PinnedArray((POOL + B, pitch), np.uint8)
    pysdyn.pack_records(step_arrays, strides, FORMS, out=pin_pool.array[:POOL])
    pin_pool.array[POOL:] = pin_pool.array[:B]
    pin_in = pysdyn.PinnedArray((POOL + B, H, W), np.uint8)
    pin_in.array[:POOL] = frames; pin_in.array[POOL:] = frames[:B]

    # a real (non-legacy) stream: libsdyn launches on it and the torch events below are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dev_frames = torch.from_numpy(pin_in.array).cuda()          # resident in HBM for the device-timed number
    dev_pool = torch.from_numpy(pin_pool.array).cuda()
    dptrs = {k: (dev_pool.data_ptr() + layout[k], 0) for k in step_arrays}
    hptrs = {k: (pin_pool.array.ctypes.data + layout[k], 0) for k in step_arrays}
    mtab = pysdyn.MapTable(len(table), device=local)
    mtab.update(0, table)
    torch.cuda.synchronize()
    if split:
        # RGB-D constructor semantics: a slot's LastFrame is the list its previous frame was TRACKED with (static keypoints +
        # re-admitted ones), so the per-keypoint id / flag rows of input j follow that list's order.  The order is a product of
        # frame j-1's own inputs (boxes, reference keypoints, F21): one untimed pass over the pool fetches it.
        orders = [None] * POOL
        for i0 in range(0, POOL, B):
            nb_ = min(B, POOL - i0)
            tin = pysdyn.track_inputs(dptrs, i0, strides, params, map_table=mtab, frame_pitch=pitch, rgbd_split=True)
            pysdyn.track_batch_device(ex, nb_, dev_frames[i0].data_ptr(), W * H, W, H, W, tin, stream.cuda_stream)
            torch.cuda.synchronize()
            o, n_all, _ = pysdyn.track_frame_order(ex, nb_)
            for b in range(nb_):
                orders[i0 + b] = o[b, :n_all[b]].copy()
        for name, empty in (("last_ids", -1), ("last_flags", 0)):
            src = step_arrays[name].copy()
            step_arrays[name][:] = empty
            for j in range(POOL):
                o = orders[(j - 1) % POOL]
                step_arrays[name][j, :len(o)] = src[j, o]
        pysdyn.pack_records(step_arrays, strides, FORMS, out=pin_pool.array[:POOL])
        pin_pool.array[POOL:] = pin_pool.array[:B]
        dev_pool.copy_(torch.from_numpy(pin_pool.array))
        torch.cuda.synchronize()

    NCTX = max(1, args.contexts)
    CTX_OFF = 37                                                # contexts start at different places of the pool

    def first_input(s, c):
        """Pool index of slot 0 of context c's step (s // NCTX): every step advances each slot's sequence by one frame."""
        return (c * CTX_OFF + s // NCTX) % POOL

    def step_device_on(s, c):
        i0 = first_input(s, c)
        tin = pysdyn.track_inputs(dptrs, i0, strides, params, map_table=mtab, frame_pitch=pitch, rgbd_split=split)
        pysdyn.track_batch_device(ctxs[c], B, dev_frames[i0].data_ptr(), W * H, W, H, W, tin, streams[c].cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") -------------------------------------------------------
    # NCTX contexts, each with its own stream, take the steps round-robin: the kernels of this path are latency-
    # rather than throughput-bound, so independent batches in flight fill each other's idle issue slots (the same
    # arrangement the end-to-end pass uses to overlap PCIe copies).
    ctxs = [ex] + [pysdyn.Extractor(nf, SCALE, NLEVELS, ini, mn, max_width=W, max_height=H, max_batch=B, device=local)
                   for _ in range(NCTX - 1)]
    streams = [stream] + [torch.cuda.Stream() for _ in range(NCTX - 1)]

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.wait_started()
    step_no = 0
    for s in range(Wm * NCTX):
        step_device_on(step_no, step_no % NCTX); step_no += 1
    barrier()
    launches0 = sum(c.launch_count() for c in ctxs)
    t_clk0 = time.perf_counter()
    # The K-step region is timed REPS times back to back (>= ~0.5 s of GPU work in total, so the driver's samplers see load);
    # value / ms_per_step are the median region, all regions are reported.
    e0 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in streams]

    def device_region():
        nonlocal step_no
        for c, st in enumerate(streams):
            e0[c].record(st)
        for s in range(K):
            step_device_on(step_no, step_no % NCTX); step_no += 1
        for c, st in enumerate(streams):
            e1[c].record(st)
        barrier()
        return max(e0[0].elapsed_time(e1[c]) for c in range(NCTX))     # first start .. last finish, on the device

    dev_runs = [device_region()]
    from pysdyn import shard
    reps = shard.region_count(dev_runs[0] * 1e-3, args.min_seconds, 1, 60)            # identical on every rank
    dev_runs += [device_region() for _ in range(reps - 1)]
    ms = float(np.median(dev_runs))
    launches = (sum(c.launch_count() for c in ctxs) - launches0) // len(dev_runs)
    launches_per_step = launches // max(K, 1)

    # per-stage device times: a separate single-stream pass on context 0 (stages of concurrent streams would overlap)
    ex.profile(True)
    for s in range(K):
        step_device_on(step_no, 0); step_no += NCTX
    barrier()
    stages = ex.profile_read()
    ex.profile(False)
    kps, desc, counts = ex.fetch(B)
    assign, locked, mask, cnt = pysdyn.track_fetch(ex, B)
    evals = pysdyn.track_stats(ex, B)
    mean_kp = float(counts.mean())

    # ---- end to end through the C ABI with pinned host buffers ("e2e") ------------------------------------
    # Per step the call uploads B frames and ONE run of B frame records (ids, flag bytes, projection records, boxes, the
    # reference frame's in-box keypoints, F21, poses) and downloads keypoints, descriptors, match indices, lock flags, mask and
    # counts.  LocalMapping's work on the map reaches the device as table updates: NEW_POINTS_PER_FRAME MapPoint records per
    # frame are uploaded with every step as well (the reference creates a few hundred MapPoints per keyframe).
    NEW_POINTS_PER_FRAME = 64
    # one staging buffer per context: a context's buffer is rewritten only after that context's previous step was waited for
    new_pts_set = [pysdyn.PinnedArray((B * NEW_POINTS_PER_FRAME,), pysdyn.MAP_POINT_DTYPE) for _ in range(NCTX)]
    new_pts = new_pts_set[0]
    out_sets = []
    for _ in ctxs:
        o = tuple(pysdyn.PinnedArray(shape, dt) for shape, dt in
                  [((B, cap), pysdyn.KP_DTYPE), ((B, cap, 32), np.uint8), ((B,), np.int32), ((B, cap), np.int32),
                   ((B, cap), np.uint8), ((B, cap), np.uint8), ((B, 4), np.int32)])
        out_sets.append((o, tuple(a.array for a in o)))

    def step_host_async(s):
        c = s % NCTX
        i0 = first_input(s, c)
        # the table rows rewritten here are rows whose content is identical (the synthetic map is static): the transfer is real
        np_c = new_pts_set[c].array
        np_c[:] = table[:len(np_c)]
        mtab.update(0, np_c, stream=ctxs[c].stream_handle())
        tin = pysdyn.track_inputs(hptrs, i0, strides, params, map_table=mtab, frame_pitch=pitch, rgbd_split=split)
        pysdyn.track_batch_host_async(ctxs[c], pin_in.array[i0:i0 + B], tin, out_sets[c][1])

    def run_host(first, count):
        for s in range(first, first + count):
            if s >= first + NCTX:
                pysdyn.track_wait(ctxs[s % NCTX])           # the step issued NCTX iterations ago on this context
            step_host_async(s)
        for s in range(max(first, first + count - NCTX), first + count):
            pysdyn.track_wait(ctxs[s % NCTX])

    run_host(step_no, 3 * NCTX); step_no += 3 * NCTX
    barrier()
    e2e_runs = []

    def e2e_region():
        nonlocal step_no
        barrier()
        t0 = time.perf_counter()
        run_host(step_no, K); step_no += K
        barrier()
        return time.perf_counter() - t0

    e2e_runs.append(e2e_region())
    ereps = shard.region_count(e2e_runs[0], args.min_seconds, 5, 60)                  # identical on every rank
    e2e_runs += [e2e_region() for _ in range(ereps - 1)]
    e2e_s = float(np.median(e2e_runs))
    clocks = sampler.stop(t_clk0, time.perf_counter()) if sampler else None

    # ---- one verification step, untimed: the same frame inputs through the device-pointer and the host-buffer entry points,
    # and a digest of this rank's results for them (rank r's digest must not depend on the number of GPUs) ----------------
    def verify_pair(host):
        outs = []
        for k in (0, 1):                    # step 0 primes the resident LastFrame, step 1 is compared
            i0 = (11 + k) % POOL
            if host:
                tin = pysdyn.track_inputs(hptrs, i0, strides, params, map_table=mtab, frame_pitch=pitch, rgbd_split=split)
                pysdyn.track_batch_host(ex, pin_in.array[i0:i0 + B], tin, out_sets[0][1])
                o = out_sets[0][1]
                outs = [o[2].copy(), o[6].copy(), o[3].copy(), o[5].copy()]
                lim = pysdyn.track_frame_order(ex, B)[1].copy() if split else outs[0]
            else:
                tin = pysdyn.track_inputs(dptrs, i0, strides, params, map_table=mtab, frame_pitch=pitch, rgbd_split=split)
                pysdyn.track_batch_device(ex, B, dev_frames[i0].data_ptr(), W * H, W, H, W, tin, stream.cuda_stream)
                torch.cuda.synchronize()                    # the fetches below run on the context's own stream
                kk, dd, nn = ex.fetch(B)
                aa, ll, mm, cc = pysdyn.track_fetch(ex, B)
                outs = [nn.copy(), cc.copy(), aa.copy(), mm.copy()]
                lim = pysdyn.track_frame_order(ex, B)[1].copy() if split else outs[0]
        return outs + [lim]                 # lim: how many entries of a frame's match list are specified (the tracked list's length)
    v_dev, v_host = verify_pair(False), verify_pair(True)
    digest_parts = []
    for name, a, b_ in zip(("n", "counts", "assign", "dyn_mask", "tracked_list_length"), v_dev, v_host):
        if a.ndim == 2 and a.shape[1] == cap:                # per-keypoint arrays: entries past a frame's list length are unspecified
            lim = (v_dev[4] if name == "assign" else v_dev[0])[:, None]
            a = np.where(np.arange(cap)[None, :] < lim, a, 0); b_ = np.where(np.arange(cap)[None, :] < lim, b_, 0)
        assert np.array_equal(a, b_), "e2e and device-resident results differ in %s: %d entries, first %s" % (
            name, int((a != b_).sum()), np.argwhere(a != b_)[:4].tolist())
        digest_parts.append(np.ascontiguousarray(a).tobytes())
    import hashlib
    # over the SPECIFIED entries only (the unspecified tails hold whatever earlier steps left there, which depends on how many
    # regions the run repeated): rank r's digest is then a function of rank r's inputs alone, the same at every N
    rank_digest = hashlib.sha256(b"".join(digest_parts)).hexdigest()[:16]

    h2d = B * W * H + B * pitch + new_pts.array.nbytes      # bytes actually copied per step (record padding included)
    # what the host link delivers for one plain pinned copy of a step's input volume (context for the e2e number)
    link = {"h2d_gbs_pinned_copy": 0.0, "h2d_bound_frames_per_s": 0.0, "note": "probe unavailable"}
    try:
        nb = int(h2d)
        # source = the bench's own pinned frame pool (cudaHostAlloc), so the probe sees the memory the end-to-end pass uploads from
        dst = torch.empty(nb, dtype=torch.uint8, device="cuda")
        nb = min(nb, pin_in.array.nbytes)
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        times = []
        for _ in range(26):
            ea.record(stream)
            rc = rt.cudaMemcpyAsync(dst.data_ptr(), pin_in.array.ctypes.data, nb, 1, ctypes.c_void_p(stream.cuda_stream))
            assert rc == 0, rc
            eb.record(stream); eb.synchronize()
            times.append(ea.elapsed_time(eb))
        times = times[2:]
        best = min(times)
        link = {"h2d_gbs_pinned_copy": nb / best / 1e6, "h2d_bound_frames_per_s": B / (best * 1e-3),
                "h2d_gbs_pinned_copy_median": nb / float(np.median(times)) / 1e6,
                "note": "best / median of 24 plain pinned copies of one step's input volume, this rank alone, right after the timed "
                        "regions; the host link is shared with whatever else runs on the box"}
        del dst
    except Exception as e:                  # a side measurement never costs the headline line
        link["note"] = "probe failed: %r" % (e,)
    d2h = B * (cap * (28 + 32) + 4 + cap * 6 + 16)

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)               # timing: max over ranks
    # NCCL all_gather of the per-rank run statistics: the only collective, off the hot path
    g = shard.gather_stats([float(B * K), mean_kp, float(cnt[:, 0].mean()), float(cnt[:, 1].mean()), float(cnt[:, 3].mean()),
                            float(int(rank_digest[:12], 16)), float(link["h2d_gbs_pinned_copy"]), float(B * K / e2e_s),
                            float(-1 if numa.get("gpu_node") is None else numa["gpu_node"])])
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    if world > 1:                       # every collective of the run is behind us: all ranks tear the communicator down together
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    if rank != 0:
        return

    total_frames = float(g[:, 0].sum())
    fps = total_frames / (ms_max * 1e-3)
    e2e_fps = total_frames / (e2e_ms_max * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    evals_per_frame = sum(evals) / B                       # Hamming evaluations per frame, counted on the device
    sab = stage_alg_bytes(W, H, int(round(mean_kp)), evals_per_frame, cap + N_MAP)
    tot_stage_ms = max(sum(v[0] for v in stages.values()), 1e-9)
    stage_report = {}
    for name, (sms, calls) in stages.items():
        if calls:
            per_step = sms / K
            ach = sab.get(name, 0) * B / (per_step * 1e-3) / 1e9
            stage_report[name] = {"ms_per_step": per_step, "share": sms / tot_stage_ms, "alg_gbs": ach, "hbm_frac": ach / peak}
    # the dominant KERNEL: stages that are one kernel (or one kernel launched per level); "match" and "dynamic" are
    # chains of small kernels and are represented by their largest member ("candidates")
    single = [n for n in stage_report if n not in ("match", "dynamic")]
    dom = max(single, key=lambda n: stage_report[n]["ms_per_step"]) if single else None
    roof = None
    # DRAM traffic per launch of the stage's main kernel, from the committed ncu --set full capture
    stage_kernel = {"fast": "k_fast", "match": "k_match_candidates", "candidates": "k_match_candidates", "describe": "k_orient_describe", "blur": "k_blur",
                    "pyramid": "k_resize", "octree": "k_octree", "level0": "k_level0", "dynamic": "k_box_stage"}
    traffic = None
    tj, ncu_note = load_ncu_figures(stage_kernel.get(dom))
    captured = cfg == "kitti" and B == 64               # the committed ncu capture is of the headline workload's launches
    if tj and dom and captured and stage_kernel.get(dom) in tj:
        traffic = tj[stage_kernel[dom]]["dram_bytes_per_launch"]
    if dom:
        r = stage_report[dom]
        roof = {"kernel": dom, "bound": "hbm", "achieved": r["alg_gbs"], "peak": peak, "unit": "GB/s",
                "frac": r["hbm_frac"], "traffic": traffic, "traffic_source": ncu_note, "peak_source": peak_src,
                "alg_bytes_per_launch": sab.get(dom, 0) * B,
                "note": "FAST scoring is integer-ALU bound, not HBM bound (DESIGN.md §Kernels)" if dom == "fast" else ""}
    balg = alg_bytes_extract(W, H, int(round(mean_kp)))
    # what actually bounds these kernels: warp-instruction issue (148 SMs x 4 schedulers x SM clock).  Instruction
    # counts per launch come from the committed ncu capture, the duration is the one measured live above.
    issue = None
    try:
        kname = stage_kernel.get(dom)
        if tj and dom and captured and kname in tj and tj[kname].get("warp_inst_per_launch") and dom in ("fast", "blur", "describe", "level0", "pyramid"):
            sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
            peak_i = 148 * 4 * sm_mhz * 1e6
            ach = tj[kname]["warp_inst_per_launch"] / (stage_report[dom]["ms_per_step"] * 1e-3)
            issue = {"kernel": kname, "warp_inst_per_launch": tj[kname]["warp_inst_per_launch"], "achieved_ginst_s": ach / 1e9,
                     "peak_ginst_s": peak_i / 1e9, "frac": ach / peak_i,
                     "note": "instruction-issue roofline of the dominant kernel: the path is issue bound, not HBM bound (DESIGN.md 4)"}
    except Exception:
        pass

    cpu = None
    if args.cpu_seconds > 0 and world == 1:             # the CPU baseline is a rank-0, N=1 measurement
        cframes, carrays, cparams, ccap = cpu_prepare(cfg, 8)
        cpre = ref_prebuild(cfg, None, carrays, cparams) if ref_available() else None
        cpu1, n1 = cpu_run(cfg, cframes, carrays, cparams, ccap, 1, seconds=args.cpu_seconds, pre=cpre)
        cpu = {"value": cpu1, "unit": "frames/s", "cores": 1, "kind": "reference" if cpre is not None else "port",
               "sample": "%d frames of the pool, full path (extract + 2 searches + dynamic mask), 1 thread (the reference's execution "
                         "model, Frame.cc:259,318,424); %s" % (n1, REF_KIND_NOTE if cpre is not None else "C++ oracle port")}
        if cpre is not None:               # the port beside it: the two agree bit for bit and should cost about the same
            cpu["port_frames_per_s_1_thread"] = cpu_run(cfg, cframes, carrays, cparams, ccap, 1, seconds=min(4.0, args.cpu_seconds))[0]
        # BASELINE.md §5 run (2): the stereo constructor's two extraction threads (Frame.cc:151-154) — pairs/s on 2 threads
        try:
            import orc
            t0 = time.perf_counter(); npairs = 0

            def one(img):
                orc.Extractor(nf, SCALE, NLEVELS, ini, mn)(img)
            while time.perf_counter() - t0 < min(4.0, args.cpu_seconds):
                th2 = [threading.Thread(target=one, args=(cframes[(npairs + j) % len(cframes)],)) for j in range(2)]
                [t.start() for t in th2]; [t.join() for t in th2]
                npairs += 1
            cpu["stereo_pairs_per_s_2_threads_extraction_only"] = npairs / (time.perf_counter() - t0)
        except Exception as e:
            cpu["stereo_pairs_error"] = repr(e)
        # run (4): the same primitives through OpenCV's own SIMD code (cv2), to show the restated primitives are not unfairly slow
        try:
            import cv2
            cv2.setNumThreads(1)
            img = cframes[0]
            oe = orc.Extractor(nf, SCALE, NLEVELS, ini, mn)
            t0 = time.perf_counter(); oe(img); t_orc = time.perf_counter() - t0
            sizes = level_sizes(W, H)
            t0 = time.perf_counter()
            lv = [img]
            for (w_, h_) in sizes[1:]:
                lv.append(cv2.resize(lv[-1], (w_, h_), interpolation=cv2.INTER_LINEAR))
            fast = cv2.FastFeatureDetector_create(mn, True)
            nk = 0
            for l in lv:
                b_ = cv2.copyMakeBorder(l, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
                nk += len(fast.detect(l))
                cv2.GaussianBlur(l, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
            t_cv = time.perf_counter() - t0
            cpu["cv2_primitives"] = {"ms_per_frame_pyramid_fast_blur_via_cv2_%s" % cv2.__version__: 1e3 * t_cv,
                                     "ms_per_frame_whole_extraction_oracle": 1e3 * t_orc,
                                     "note": "cv2: 8-level resize chain + borders + whole-level FAST(minTh, NMS) + 7x7 blur, 1 thread, SIMD; "
                                             "the oracle figure also contains the per-cell FAST calls, octree, orientation and descriptors"}
        except Exception as e:
            cpu["cv2_primitives"] = {"error": repr(e)}

    extras = None
    if args.next_rows and world == 1 and cfg == "kitti":
        try:
            extras = next_rows(pysdyn, torch, cfg, W, H, nf, ini, mn, local, B, max(K // 2, 5), rank, dptrs, strides, params, dev_frames,
                               mtab, pitch, host=(hptrs, pin_in))
        except Exception as e:          # never lose the headline line to a side measurement
            extras = {"error": repr(e)}

    config = workload_config(cfg, B)
    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": config,
        "run": {"contexts_per_gpu": NCTX, "sequences_per_gpu": NCTX * B,
                "inputs": "resident MapPoint table (%.0f MB) + resident LastFrame; per frame: image, MapPoint ids + state bytes of "
                          "LastFrame and of the local map (Frame::isInFrustum runs on the device), boxes, reference in-box keypoints, "
                          "F21, pose pair (one %d-byte record)" % (table.nbytes / 1e6, pitch),
                "l2": "inputs cycle through a %d-frame pool (%.0f MB of frames) and each step's working set "
                      "(~%.0f MB) exceeds the 126 MB L2" % (POOL, POOL * W * H / 1e6, B * 7.0),
                "device_regions_ms": [round(v, 3) for v in dev_runs],
                "timing": "K steps per region, regions repeated back to back for >= %.1f s per arm; value / e2e = median region" % args.min_seconds},
        "rank_digests": ["%012x" % int(v) for v in g[:, 5]],
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "runs_ms": [round(1e3 * v, 3) for v in e2e_runs], "timing": "median of %d back-to-back K-step regions (host clock, "
                "barrier + cudaDeviceSynchronize on both sides)" % len(e2e_runs),
                "map_updates_per_step": "%d MapPoint records (%d B)" % (len(new_pts.array), new_pts.array.nbytes),
                "h2d_gbs_achieved_per_gpu": h2d * (e2e_fps / world / B) / 1e9, "d2h_gbs_achieved_per_gpu": d2h * (e2e_fps / world / B) / 1e9},
        "host_link": link,
        "per_rank": {"h2d_gbs_pinned_copy_alone": [round(float(v), 1) for v in g[:, 6]], "e2e_frames_per_s": [round(float(v)) for v in g[:, 7]],
                     "gpu_numa_node": [int(v) for v in g[:, 8]], "numa_binding_rank0": numa},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "pipeline_roofline": {"alg_bytes_per_frame": balg, "achieved_gbs": balg * fps / world / 1e9, "peak": peak,
                              "frac": balg * fps / world / 1e9 / peak,
                              "note": "SURVEY §8(d) extraction bytes per frame x per-GPU frames/s"},
        "issue_roofline": issue,
        "stages": stage_report,
        "matching": {"hamming_evals_per_frame": evals_per_frame, "evals_per_s": evals_per_frame * fps / world,
                     "popc_per_s": 8 * evals_per_frame * fps / world, "match_stage_ms_per_step": stage_report.get("match", {}).get("ms_per_step"),
                     "evals_per_s_inside_the_match_stage": (evals_per_frame * B / (stage_report["match"]["ms_per_step"] * 1e-3)) if "match" in stage_report else None,
                     "popc_peak_per_s": POPC_PEAK, "popc_pipe_frac_inside_the_match_stage":
                         (8 * evals_per_frame * B / (stage_report["match"]["ms_per_step"] * 1e-3) / POPC_PEAK) if "match" in stage_report else None,
                     "note": "8 POPC.B32 per 256-bit distance (ORBmatcher.cc:1804-1820); peak = measured by tools/popc_probe (profiles/r02_popc_probe.json). "
                             "The searches are bound by their dependent gather chains, not by the popc pipe (DESIGN.md 4)"},
        "per_frame": {"hamming_evals": evals_per_frame, "keypoints": float(g[:, 1].mean()), "matches_frame": float(g[:, 2].mean()),
                      "matches_map": float(g[:, 3].mean()), "dyn_masked": float(g[:, 4].mean())},
        "cpu_baseline": cpu,
        "next_rows": extras,
    }
    emit(line)


if __name__ == "__main__":
    main()
