"""Independent Python twin of the reference extractor built on the REAL cv2 primitives.

TEST INFRASTRUCTURE ONLY.  Follows src/ORBextractor.cc:410-470, 539-853, 1043-1132 using cv2.resize /
copyMakeBorder / FastFeatureDetector / GaussianBlur / fastAtan2 (cv2 4.13.0), so it pins the C++ oracle
end-to-end to OpenCV's actual arithmetic.  Slow (pure Python loops) — used on a handful of frames and to
generate tests/golden/.  Tie-break of equal-size octree nodes: later-created first (SURVEY B-1).
"""
import math
import os
import re
import numpy as np
import cv2

EDGE = 19
HALF = 15
f32 = np.float32


def load_pattern():
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "sdyn_brief_pattern.inc")
    txt = re.sub(r"/\*.*?\*/", "", open(inc).read(), flags=re.S)
    return np.array([int(v) for v in re.findall(r"-?\d+", txt)], np.int32).reshape(256, 4)


PATTERN = load_pattern()


def rint(v):
    return int(np.rint(v))


class Twin:
    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th):
        self.nfeatures, self.nlevels, self.ini, self.min = nfeatures, nlevels, ini_th, min_th
        sf = float(f32(scale_factor))  # double member initialised from a float
        self.scale = [f32(1.0)]
        for i in range(1, nlevels):
            self.scale.append(f32(float(self.scale[-1]) * sf))
        self.inv = [f32(1.0) / s for s in self.scale]
        factor = f32(1.0 / sf)
        want = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
        self.quota, s = [], 0
        for _ in range(nlevels - 1):
            q = rint(want)
            self.quota.append(q)
            s += q
            want = f32(want * factor)
        self.quota.append(max(nfeatures - s, 0))
        umax = [0] * (HALF + 1)
        vmax = int(math.floor(float(f32(HALF) * np.sqrt(f32(2.0)) / f32(2) + f32(1))))
        vmin = int(math.ceil(float(f32(HALF) * np.sqrt(f32(2.0)) / f32(2))))
        for v in range(vmax + 1):
            umax[v] = rint(math.sqrt(HALF * HALF - v * v))
        v0 = 0
        for v in range(HALF, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0
            v0 += 1
        self.umax = umax

    def pyramid(self, img):
        h, w = img.shape
        self.pyr = []
        for l in range(self.nlevels):
            s = self.inv[l]
            sw, sh = rint(f32(w) * s), rint(f32(h) * s)
            if l == 0:
                cur = img
            else:
                prev = self.pyr[-1][EDGE:-EDGE, EDGE:-EDGE]
                cur = cv2.resize(prev, (sw, sh), interpolation=cv2.INTER_LINEAR)
            self.pyr.append(cv2.copyMakeBorder(cur, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101))
        return self.pyr

    def octree(self, keys, minX, maxX, minY, maxY, N):
        """keys: list of (x, y, resp). Returns kept keys in list order."""
        nIni = int(np.round(f32(maxX - minX) / f32(maxY - minY)) if True else 0)
        # C round(): half away from zero
        r = float(f32(maxX - minX) / f32(maxY - minY))
        nIni = int(math.floor(r + 0.5))
        hX = f32(maxX - minX) / f32(nIni)
        seq = [0]

        def node(ulx, urx, uly, bry):
            seq[0] += 1
            return {"k": [], "ulx": ulx, "urx": urx, "uly": uly, "bry": bry, "leaf": False, "seq": seq[0]}

        L = []
        for i in range(nIni):
            L.append(node(int(hX * f32(i)), int(hX * f32(i + 1)), 0, maxY - minY))
        for k in keys:
            L[int(f32(k[0]) / hX)]["k"].append(k)
        L = [n for n in L if n["k"]]
        for n in L:
            if len(n["k"]) == 1:
                n["leaf"] = True

        def divide(p):
            hx = int(math.ceil(float(f32(p["urx"] - p["ulx"]) / f32(2))))
            hy = int(math.ceil(float(f32(p["bry"] - p["uly"]) / f32(2))))
            mx, my = p["ulx"] + hx, p["uly"] + hy
            c = [node(p["ulx"], mx, p["uly"], my), node(mx, p["urx"], p["uly"], my),
                 node(p["ulx"], mx, my, p["bry"]), node(mx, p["urx"], my, p["bry"])]
            for k in p["k"]:
                if k[0] < mx:
                    c[0 if k[1] < my else 2]["k"].append(k)
                else:
                    c[1 if k[1] < my else 3]["k"].append(k)
            for n in c:
                if len(n["k"]) == 1:
                    n["leaf"] = True
            return c

        done = False
        while not done:
            prev_size = len(L)
            front, rest, expandable = [], [], []
            for n in L:
                if n["leaf"]:
                    rest.append(n)
                    continue
                for c in divide(n):
                    if c["k"]:
                        front.insert(0, c)
                        if len(c["k"]) > 1:
                            expandable.append(c)
            L = front + rest
            if len(L) >= N or len(L) == prev_size:
                done = True
            elif len(L) + 3 * len(expandable) > N:
                while not done:
                    prev_size = len(L)
                    prev = sorted(expandable, key=lambda n: (len(n["k"]), n["seq"]))
                    expandable = []
                    for p in reversed(prev):
                        for c in divide(p):
                            if c["k"]:
                                L.insert(0, c)
                                if len(c["k"]) > 1:
                                    expandable.append(c)
                        L.remove(p)
                        if len(L) >= N:
                            break
                    if len(L) >= N or len(L) == prev_size:
                        done = True
        out = []
        for n in L:
            best = n["k"][0]
            for k in n["k"][1:]:
                if k[2] > best[2]:
                    best = k
            out.append(best)
        return out

    def keypoints(self):
        res = []
        for l in range(self.nlevels):
            im = self.pyr[l][EDGE:-EDGE, EDGE:-EDGE]
            rows, cols = im.shape
            minBX = minBY = EDGE - 3
            maxBX, maxBY = cols - EDGE + 3, rows - EDGE + 3
            width, height = f32(maxBX - minBX), f32(maxBY - minBY)
            nCols, nRows = int(width / f32(30)), int(height / f32(30))
            wCell, hCell = int(math.ceil(float(width / f32(nCols)))), int(math.ceil(float(height / f32(nRows))))
            keys = []
            det = {th: cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True) for th in (self.ini, self.min)}
            for i in range(nRows):
                iniY = minBY + i * hCell
                maxY = iniY + hCell + 6
                if iniY >= maxBY - 3:
                    continue
                maxY = min(maxY, maxBY)
                for j in range(nCols):
                    iniX = minBX + j * wCell
                    maxX = iniX + wCell + 6
                    if iniX >= maxBX - 6:
                        continue
                    maxX = min(maxX, maxBX)
                    cell = np.ascontiguousarray(im[iniY:maxY, iniX:maxX])
                    kp = det[self.ini].detect(cell, None)
                    if not kp:
                        kp = det[self.min].detect(cell, None)
                    for p in kp:
                        keys.append((f32(p.pt[0] + j * wCell), f32(p.pt[1] + i * hCell), f32(p.response)))
            kept = self.octree(keys, minBX, maxBX, minBY, maxBY, self.quota[l])
            size = f32(int(f32(31) * self.scale[l]))
            res.append([(k[0] + f32(minBX), k[1] + f32(minBY), k[2], size) for k in kept])
        return res

    def ic_angle(self, im, x, y):
        m01 = m10 = 0
        for v in range(-HALF, HALF + 1):
            d = self.umax[abs(v)]
            row = im[y + v, x - d:x + d + 1].astype(np.int64)
            m10 += int((np.arange(-d, d + 1) * row).sum())
            m01 += v * int(row.sum())
        return f32(cv2.fastAtan2(float(m01), float(m10)))

    @staticmethod
    def descriptor(blur, x, y, angle):
        ang = f32(angle) * f32(np.pi / f32(180.0))
        a, b = f32(np.cos(ang, dtype=f32)), f32(np.sin(ang, dtype=f32))
        px0, py0, px1, py1 = (PATTERN[:, i].astype(f32) for i in range(4))

        def samp(px, py):
            yy = np.rint(px * b + py * a).astype(np.int64)
            xx = np.rint(px * a - py * b).astype(np.int64)
            return blur[y + yy, x + xx].astype(np.int32)

        bits = (samp(px0, py0) < samp(px1, py1)).astype(np.uint8)
        return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").reshape(32)

    def __call__(self, img):
        self.pyramid(img)
        per_level = self.keypoints()
        kps, descs = [], []
        for l, ks in enumerate(per_level):
            if not ks:
                continue
            im = np.ascontiguousarray(self.pyr[l][EDGE:-EDGE, EDGE:-EDGE])
            blur = cv2.GaussianBlur(im.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
            for (x, y, r, size) in ks:
                xi, yi = rint(x), rint(y)
                ang = self.ic_angle(im, xi, yi)
                descs.append(self.descriptor(blur, xi, yi, ang))
                s = self.scale[l]
                kps.append((x * s if l else x, y * s if l else y, size, ang, r, l, -1))
        return kps, (np.array(descs, np.uint8).reshape(-1, 32))
