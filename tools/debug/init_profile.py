import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("slam-dynamic_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import common, pysdyn, scenario
W, H, nrect, nf, ini, mn = common.CONFIGS["kitti"]
ex = pysdyn.Extractor(2 * nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=1)
ka, da = ex(common.frame("kitti", 0)); kb, db = ex(common.frame("kitti", 1, ox=4, oy=1, t=1))
sc = ex.GetScaleFactors()
F1 = scenario.frame_view(ka, da, sc, W, H); F2 = scenario.frame_view(kb, db, sc, W, H)
prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
m = pysdyn.Matcher(ex, 0.9, True)
for _ in range(3):
    m.SearchForInitialization(F1, F2, prev, 100)
t = []
for _ in range(10):
    t0 = time.perf_counter(); n, _, _ = m.SearchForInitialization(F1, F2, prev, 100); t.append(time.perf_counter() - t0)
print("init search: %.3f ms, %d matches, %d evals" % (1e3 * np.median(t), n, m.last_evals()))
