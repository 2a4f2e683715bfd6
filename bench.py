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
POOL = 256          # distinct frames per GPU (frame_idx 0..255)
NLEVELS, SCALE = 8, 1.2
N_MAP = 3000        # local-map points per frame (2000-5000)
REF_STRIDE = 1024   # capacity for the reference frame's in-box keypoints
METRIC = "frames/sec ORB extract+match+dyn-mask @KITTI 1241x376 2k feats; % HBM roofline"


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


def make_frames(cfg, rank, first, count, disparity=0):
    """Frames first..first+count-1 of this rank's sequence (consecutive frames shift by <= 8 px); disparity > 0 renders
    the right view of a rectified stereo rig (content shifted, its own sensor noise)."""
    import pysdyn
    import scenario
    W, H, nrect, _, _, _, cid = WORKLOADS[cfg]
    frames = np.empty((count, H, W), np.uint8)
    nthreads = min(os.cpu_count() or 1, 16)

    def work(t):
        for j in range(t, count, nthreads):
            i = first + j
            ox, oy = scenario.sequence_offsets(i)
            pysdyn.synth_frame(seq_seed(cfg, rank), 1000 * cid + 100000 * rank + i + (1 << 20) + (disparity << 22), W, H, nrect,
                               ox + disparity, oy, scenario.sequence_time(i), out=frames[j])

    th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [t.start() for t in th]
    [t.join() for t in th]
    return frames


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
def cpu_prepare(cfg, nframes):
    """Inputs of the CPU arm: a short sequence, extracted once to build the track queries."""
    import orc
    import scenario
    W, H, nrect, nf, ini, mn, _ = WORKLOADS[cfg]
    frames = make_frames(cfg, 0, 0, nframes + 1)
    ex = orc.Extractor(nf, SCALE, NLEVELS, ini, mn)
    kd = [ex(im) for im in frames]
    cap = nf + 200
    arrays = scenario.build_track_batch(kd, seq_seed(cfg, 0), 1, W, H, nrect, NLEVELS, cap, N_MAP, REF_STRIDE,
                                        n_map=N_MAP, seed=3)
    return frames[1:], arrays, scenario.track_params(W, H), cap


def cpu_run(cfg, frames, arrays, params, cap, threads, seconds=None, count=None):
    """Runs the full per-frame hot path (extract + 2 searches + dynamic mask) on the CPU oracle, frame-parallel
    over `threads` host threads, for `seconds` or for `count` frames.  Returns (fps, frames_done)."""
    import orc
    import oracle_track
    W, H, _, nf, ini, mn, _ = WORKLOADS[cfg]
    done = [0] * threads
    t0 = time.perf_counter()

    def work(t):
        ex = orc.Extractor(nf, SCALE, NLEVELS, ini, mn)
        i = t
        while True:
            if seconds is not None and time.perf_counter() - t0 >= seconds:
                break
            if count is not None and i >= count:
                break
            f = i % len(frames)
            k, d = ex(frames[f])
            oracle_track.track_frame(k, d, ex.scale, W, H, arrays, f, params, cap)
            done[t] += 1
            i += threads

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    return sum(done) / dt, sum(done)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.workload
    W, H = WORKLOADS[cfg][:2]
    threads = os.cpu_count() or 1
    frames, arrays, params, cap = cpu_prepare(cfg, 16)
    per_step = 2 * threads                      # bounded sample per step
    if args.warmup > 0:
        cpu_run(cfg, frames, arrays, params, cap, threads, count=threads)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        total += cpu_run(cfg, frames, arrays, params, cap, threads, count=per_step)[1]
    dt = time.perf_counter() - t0
    fps = total / dt
    emit(({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "%s %dx%d" % (cfg, W, H), "frames_per_step": per_step,
                   "stages": "extract + SearchByProjection(frame) + SearchByProjection(map) + dynamic mask"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d frames per step, frame-parallel C++ oracle (restated reference CPU path)" % per_step},
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


def next_rows(pysdyn, torch, cfg, W, H, nf, ini, mn, local, B, steps, rank, dptrs, strides, params, dev_frames):
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
        dright = torch.from_numpy(make_frames(cfg, rank, 1, Bs, disparity=11)).cuda()
        sp = dict(params); sp["mono"] = 0
        tin = pysdyn.track_inputs(dptrs, 0, strides, sp)

        def stereo_track_step():
            pysdyn.track_batch_stereo_device(L, R, Bs, dev_frames[0].data_ptr(), dright.data_ptr(), W * H, W, H, W, tin, mb, mbf)

        for _ in range(3):
            stereo_track_step()
        L.sync(); R.sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            stereo_track_step()
        L.sync(); R.sync()
        dt3 = time.perf_counter() - t0
        ur3, dp3, kept3 = pysdyn.stereo_fetch(L, Bs)
        a3, l3, m3, c3 = pysdyn.track_fetch(L, Bs)
        out["stereo_track_config3"] = {"pairs_per_s": Bs * steps / dt3, "ms_per_step": 1e3 * dt3 / steps, "pairs_per_step": Bs,
                                       "stereo_points_per_pair": float(kept3.mean()), "matches_frame": float(c3[:, 0].mean()),
                                       "matches_map": float(c3[:, 1].mean()),
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
        out["single_frame_latency"] = {"extract_ms": 1e3 * float(np.median(lat)), "search_by_projection_frame_ms": 1e3 * float(np.median(lm)),
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


def main():
    # Libraries (NCCL's version banner, torchrun notices) write to fd 1; keep stdout for the result line only.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sdyn", choices=["sdyn", "reference"])
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=64, help="frames per step per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample length (0 = skip)")
    ap.add_argument("--contexts", type=int, default=6, help="contexts (streams) per GPU taking steps round-robin")
    ap.add_argument("--next-rows", type=int, default=1, help="also time the SURVEY 8(f) rows (stereo, BoW) on rank 0 at N=1")
    args = ap.parse_args()
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    cfg = args.workload
    W, H, nrect, nf, ini, mn, _ = WORKLOADS[cfg]
    B = args.batch
    K, Wm = args.steps, max(args.warmup, 3)
    nsets = POOL // B
    assert nsets >= 1

    # ---- inputs: this rank's sequence (sharded by sequence: no data-path collective) ----------------------
    frames = make_frames(cfg, rank, 0, POOL + 1)                # frame 0 only serves as LastFrame of frame 1
    ex = pysdyn.Extractor(nf, SCALE, NLEVELS, ini, mn, max_width=W, max_height=H, max_batch=B, device=local)
    cap = ex.cap
    kd = []
    for s in range(0, POOL + 1, B):                             # untimed pre-pass: LastFrame / map / box inputs
        chunk = frames[s:s + B]
        k, d, n = ex.extract_batch(chunk)
        kd += [(k[i, :n[i]].copy(), d[i, :n[i]].copy()) for i in range(len(chunk))]
    # LastFrame arrays are sized by the data (largest keypoint count of the sequence, rounded up), like the reference's vectors
    last_stride = min(cap, (max(len(k) for k, _ in kd) + 63) // 64 * 64)
    arrays = scenario.build_track_batch(kd, seq_seed(cfg, rank), 1, W, H, nrect, NLEVELS, last_stride, N_MAP, REF_STRIDE,
                                        n_map=N_MAP, seed=3)
    params = scenario.track_params(W, H)
    # the reference frame's in-box keypoint block is sized by the data (largest per-frame total, rounded up)
    ref_stride = min(REF_STRIDE, max(64, (int(arrays["ref_off"].reshape(len(arrays["ref_off"]), -1)[:, -1].max()) + 63) // 64 * 64))
    arrays["ref_desc"] = np.ascontiguousarray(arrays["ref_desc"].reshape(len(arrays["ref_desc"]), REF_STRIDE, 32)[:, :ref_stride])
    arrays["ref_xy"] = np.ascontiguousarray(arrays["ref_xy"].reshape(len(arrays["ref_xy"]), REF_STRIDE, 2)[:, :ref_stride])
    strides = (last_stride, N_MAP, ref_stride)
    # undistorted camera: mvKeysUn is mvKeys (src/Frame.cc:814-818) -> pass the same array for both
    keys_un_alias = np.array_equal(arrays["last_keys"], arrays["last_keys_un"])
    cur_frames = frames[1:]

    # a real (non-legacy) stream: libsdyn launches on it and the torch events below are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dev_frames = torch.from_numpy(cur_frames).cuda()            # resident in HBM for the device-timed number
    dev = {k: torch.from_numpy(v.view(np.uint8).reshape(v.shape[0], -1)).cuda() for k, v in arrays.items()}
    dptrs = {k: (t.data_ptr(), t.shape[1]) for k, t in dev.items()}
    if keys_un_alias:
        dptrs["last_keys_un"] = dptrs["last_keys"]
    torch.cuda.synchronize()

    def step_device(s):
        base = (s % nsets) * B
        tin = pysdyn.track_inputs(dptrs, base, strides, params)
        pysdyn.track_batch_device(ex, B, dev_frames[base].data_ptr(), W * H, W, H, W, tin, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") -------------------------------------------------------
    # NCTX contexts, each with its own stream, take the steps round-robin: the kernels of this path are latency-
    # rather than throughput-bound, so independent batches in flight fill each other's idle issue slots (the same
    # arrangement the end-to-end pass uses to overlap PCIe copies).
    NCTX = max(1, args.contexts)
    ctxs = [ex] + [pysdyn.Extractor(nf, SCALE, NLEVELS, ini, mn, max_width=W, max_height=H, max_batch=B, device=local)
                   for _ in range(NCTX - 1)]
    streams = [stream] + [torch.cuda.Stream() for _ in range(NCTX - 1)]

    def step_device_on(s, c):
        base = (s % nsets) * B
        tin = pysdyn.track_inputs(dptrs, base, strides, params)
        pysdyn.track_batch_device(ctxs[c], B, dev_frames[base].data_ptr(), W * H, W, H, W, tin, streams[c].cuda_stream)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.wait_started()
    for s in range(Wm * NCTX):
        step_device_on(s, s % NCTX)
    barrier()
    launches0 = sum(c.launch_count() for c in ctxs)
    t_clk0 = time.perf_counter()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    for c, st in enumerate(streams):
        e0[c].record(st)
    for s in range(K):
        step_device_on(Wm + s, s % NCTX)
    for c, st in enumerate(streams):
        e1[c].record(st)
    barrier()
    ms = max(e0[0].elapsed_time(e1[c]) for c in range(NCTX))     # first start .. last finish, on the device
    launches = sum(c.launch_count() for c in ctxs) - launches0
    launches_per_step = launches // max(K, 1)

    # per-stage device times: a separate single-stream pass (stages of concurrent streams would overlap)
    ex.profile(True)
    for s in range(K):
        step_device(Wm + s)
    barrier()
    stages = ex.profile_read()
    ex.profile(False)
    kps, desc, counts = ex.fetch(B)
    assign, locked, mask, cnt = pysdyn.track_fetch(ex, B)
    evals = pysdyn.track_stats(ex, B)
    mean_kp = float(counts.mean())

    # ---- end to end through the C ABI with pinned host buffers ("e2e") ------------------------------------
    # the query arrays of a step live in ONE pinned block at sdyn_track_input_layout's offsets: libsdyn uploads such a block
    # with a single copy (a dozen small copies cost more PCIe time than their bytes)
    layout, block_bytes = pysdyn.track_input_layout(B, strides, not keys_un_alias)
    blocks, hptr_sets = [], []
    for sset in range(nsets):
        blk = pysdyn.PinnedArray((block_bytes,), np.uint8)
        blk.array[:] = 0
        ptrs = {}
        for k, v in arrays.items():
            rows = v.view(np.uint8).reshape(v.shape[0], -1)
            if not (keys_un_alias and k == "last_keys_un"):
                chunk = rows[sset * B:(sset + 1) * B].reshape(-1)
                blk.array[layout[k]:layout[k] + chunk.size] = chunk
            ptrs[k] = (blk.array.ctypes.data + layout[k], rows.shape[1])
        if keys_un_alias:
            ptrs["last_keys_un"] = ptrs["last_keys"]
        blocks.append(blk)
        hptr_sets.append(ptrs)
    pin_in = pysdyn.PinnedArray((POOL, H, W), np.uint8)
    pin_in.array[:] = cur_frames
    # the same NCTX contexts round-robin: a step's PCIe copies overlap the other contexts' kernels
    out_sets = []
    for _ in ctxs:
        o = tuple(pysdyn.PinnedArray(shape, dt) for shape, dt in
                  [((B, cap), pysdyn.KP_DTYPE), ((B, cap, 32), np.uint8), ((B,), np.int32), ((B, cap), np.int32),
                   ((B, cap), np.uint8), ((B, cap), np.uint8), ((B, 4), np.int32)])
        out_sets.append((o, tuple(a.array for a in o)))

    def step_host_async(s):
        base = (s % nsets) * B
        tin = pysdyn.track_inputs(hptr_sets[s % nsets], 0, strides, params)
        pysdyn.track_batch_host_async(ctxs[s % NCTX], pin_in.array[base:base + B], tin, out_sets[s % NCTX][1])

    def run_host(first, count):
        for s in range(first, first + count):
            if s >= first + NCTX:
                pysdyn.track_wait(ctxs[s % NCTX])           # the step issued NCTX iterations ago on this context
            step_host_async(s)
        for s in range(max(first, first + count - NCTX), first + count):
            pysdyn.track_wait(ctxs[s % NCTX])

    run_host(0, 2 * NCTX)
    barrier()
    # K steps take ~25 ms, short enough for a single host hiccup to move the number by 20 %: the K-step region is
    # timed five times back to back and the median is reported (all five are in e2e.runs_ms)
    e2e_runs = []
    for rep in range(5):
        barrier()
        t0 = time.perf_counter()
        run_host(Wm, K)
        barrier()
        e2e_runs.append(time.perf_counter() - t0)
    e2e_s = float(np.median(e2e_runs))
    clocks = sampler.stop(t_clk0, time.perf_counter()) if sampler else None
    # the e2e outputs of the last step must equal the device-resident run's results for the same frames
    last = Wm + K - 1
    if (last % nsets) == ((Wm + K - 1) % nsets):
        eo = out_sets[last % NCTX][1]
        assert np.array_equal(eo[2], counts) and np.array_equal(eo[6], cnt), "e2e and device-resident results differ"
    h2d = B * W * H + block_bytes                       # bytes actually copied per step (block padding included)
    # what the host link delivers for one plain pinned copy of a step's input volume (context for the e2e number)
    link = None
    if rank == 0:
        nb = int(h2d)
        src = torch.empty(nb, dtype=torch.uint8).pin_memory(); dst = torch.empty(nb, dtype=torch.uint8, device="cuda")
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            ea.record(stream); dst.copy_(src, non_blocking=True); eb.record(stream); eb.synchronize()
            best = min(best, ea.elapsed_time(eb))
        link = {"h2d_gbs_pinned_copy": nb / best / 1e6, "h2d_bound_frames_per_s": B / (best * 1e-3)}
        del src, dst
    d2h = B * (cap * (28 + 32) + 4 + cap * 6 + 16)

    from pysdyn import shard
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)               # timing: max over ranks
    # NCCL all_gather of the per-rank run statistics: the only collective, off the hot path
    g = shard.gather_stats([float(B * K), mean_kp, float(cnt[:, 0].mean()), float(cnt[:, 1].mean()), float(cnt[:, 3].mean()),
                            float(shard.result_hash(counts, cnt))])
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
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
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        if dom and B == 64 and stage_kernel.get(dom) in tj:
            traffic = tj[stage_kernel[dom]]["dram_bytes_per_launch"]
    except Exception:
        pass
    if dom:
        r = stage_report[dom]
        roof = {"kernel": dom, "bound": "hbm", "achieved": r["alg_gbs"], "peak": peak, "unit": "GB/s",
                "frac": r["hbm_frac"], "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": sab.get(dom, 0) * B,
                "note": "FAST scoring is integer-ALU bound, not HBM bound (DESIGN.md §Kernels)" if dom == "fast" else ""}
    balg = alg_bytes_extract(W, H, int(round(mean_kp)))
    # what actually bounds these kernels: warp-instruction issue (148 SMs x 4 schedulers x SM clock).  Instruction
    # counts per launch come from the committed ncu capture, the duration is the one measured live above.
    issue = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        kname = stage_kernel.get(dom)
        if dom and B == 64 and kname in tj and tj[kname].get("warp_inst_per_launch") and dom in ("fast", "blur", "describe", "level0"):
            sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
            peak_i = 148 * 4 * sm_mhz * 1e6
            ach = tj[kname]["warp_inst_per_launch"] / (stage_report[dom]["ms_per_step"] * 1e-3)
            issue = {"kernel": kname, "warp_inst_per_launch": tj[kname]["warp_inst_per_launch"], "achieved_ginst_s": ach / 1e9,
                     "peak_ginst_s": peak_i / 1e9, "frac": ach / peak_i,
                     "note": "instruction-issue roofline of the dominant kernel: the path is issue bound, not HBM bound (DESIGN.md 4)"}
    except Exception:
        pass

    cpu = None
    if args.cpu_seconds > 0:
        cframes, carrays, cparams, ccap = cpu_prepare(cfg, 8)
        cpu1, n1 = cpu_run(cfg, cframes, carrays, cparams, ccap, 1, seconds=args.cpu_seconds)
        cpu = {"value": cpu1, "unit": "frames/s", "cores": 1, "kind": "port",
               "sample": "%d KITTI frames, full path (extract + 2 searches + dynamic mask) on the C++ oracle, 1 thread "
                         "(the reference's execution model, Frame.cc:259,318,424)" % n1}

    extras = None
    if args.next_rows and world == 1 and cfg == "kitti":
        try:
            extras = next_rows(pysdyn, torch, cfg, W, H, nf, ini, mn, local, B, max(K // 2, 5), rank, dptrs, strides, params, dev_frames)
        except Exception as e:          # never lose the headline line to a side measurement
            extras = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "%s %dx%d nfeatures=%d levels=%d scale=%.1f iniTh=%d minTh=%d" % (cfg, W, H, nf, NLEVELS, SCALE, ini, mn),
                   "frames_per_step_per_gpu": B, "map_points_per_frame": N_MAP, "contexts_per_gpu": NCTX,
                   "stages": "extract + SearchByProjection(cur,last) + SearchByProjection(F,map) + dynamic mask",
                   "sharding": "one sequence per rank, no data-path collective; NCCL all_gather of run statistics only",
                   "l2": "inputs cycle through a %d-frame pool (%.0f MB of frames) and each step's working set "
                         "(~%.0f MB) exceeds the 126 MB L2" % (POOL, POOL * W * H / 1e6, B * 7.0)},
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "runs_ms": [round(1e3 * v, 3) for v in e2e_runs], "timing": "median of 5 back-to-back K-step regions (host clock, "
                "barrier + cudaDeviceSynchronize on both sides)"},
        "host_link": link,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "pipeline_roofline": {"alg_bytes_per_frame": balg, "achieved_gbs": balg * fps / world / 1e9, "peak": peak,
                              "frac": balg * fps / world / 1e9 / peak,
                              "note": "SURVEY §8(d) extraction bytes per frame x per-GPU frames/s"},
        "issue_roofline": issue,
        "stages": stage_report,
        "per_frame": {"hamming_evals": evals_per_frame, "keypoints": float(g[:, 1].mean()), "matches_frame": float(g[:, 2].mean()),
                      "matches_map": float(g[:, 3].mean()), "dyn_masked": float(g[:, 4].mean())},
        "cpu_baseline": cpu,
        "next_rows": extras,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
