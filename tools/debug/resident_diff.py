import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("slam-dynamic_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import common, orc, pysdyn, scenario, oracle_track
from test_gpu_track import _sequence, _dev
cfg, B = "tum", 3
W, H, nrect, nf, ini, mn, seq_seed, frames, cpu, kd = _sequence(cfg, B + 2)
gpu = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
last_stride, map_stride, ref_stride = gpu.cap, 1500, 512
strides = (last_stride, map_stride, ref_stride)
params = scenario.track_params(W, H)
a1 = scenario.build_track_batch(kd[:B + 1], seq_seed, 1, W, H, nrect, 8, *strides, n_map=1500, seed=3)
a2 = scenario.build_track_batch(kd[1:B + 2], seq_seed, 2, W, H, nrect, 8, *strides, n_map=1500, seed=3)
d1, p1 = _dev(a1); d2, p2 = _dev(a2)
f1 = torch.from_numpy(frames[1:B + 1]).cuda(); f2 = torch.from_numpy(frames[2:B + 2]).cuda()
pysdyn.track_batch_device(gpu, B, f2.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(p2, 0, strides, params))
want = [x.copy() for x in pysdyn.track_fetch(gpu, B)]
table, res = scenario.resident_forms(a2)
mt = pysdyn.MapTable(len(table)); mt.update(0, table)
torch.cuda.synchronize()
dr, pr = _dev(res)
for mode in ("both", "last_only", "map_only"):
    rp = dict(p2)
    if mode in ("both", "last_only"):
        for k in ("last_points", "last_keys", "last_keys_un", "n_last"): rp.pop(k)
        rp["last_ids"] = pr["last_ids"]; rp["last_flags"] = pr["last_flags"]
    if mode in ("both", "map_only"):
        rp.pop("map_points"); rp["map_ids"] = pr["map_ids"]; rp["map_proj"] = pr["map_proj"]
    pysdyn.track_batch_device(gpu, B, f1.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(p1, 0, strides, params))
    pysdyn.track_batch_device(gpu, B, f2.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(rp, 0, strides, params, map_table=mt))
    got = pysdyn.track_fetch(gpu, B)
    print(mode, [bool(np.array_equal(a, b)) for a, b in zip(got, want)], got[3].tolist(), want[3].tolist())
    for f in range(B):
        d = np.nonzero(got[1][f] != want[1][f])[0]
        if len(d):
            i = d[:8]
            print("  frame", f, "locked differs at", len(d), "kps", i.tolist(), "assign", got[0][f, i].tolist(), "got", got[1][f, i].tolist(), "want", want[1][f, i].tolist())
            for k in i[:4]:
                a = got[0][f, k]
                if a >= last_stride: print("     map q", a - last_stride, "obs", a2["map_points"]["obs_positive"][f, a - last_stride])
                elif a >= 0: print("     last q", a, "obs", a2["last_points"]["obs_positive"][f, a], "flags", res["last_flags"][f, a])
