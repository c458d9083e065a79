"""cv2-driven restatement of ORBextractor::operator() (TEST INFRASTRUCTURE ONLY).

Every third-party primitive the reference calls is executed by the real OpenCV (cv2) here —
cv2.resize, cv2.copyMakeBorder, cv2.FastFeatureDetector, cv2.GaussianBlur, cv2.fastAtan2 — and
cosf/sinf come from the host libm through ctypes, so this file is authoritative for the
third-party arithmetic; the glue follows /root/reference/Source/Libraries/ORB_SLAM2/src/
ORBextractor.cc line by line (citations inline).  It pins oracle_cpu.cpp (tests/
test_oracle_vs_cv2.py) and generates tests/golden/* (tests/golden/make_golden.py).
Pure-Python loops: use on single frames only.
"""
import ctypes
import ctypes.util
import math

import cv2
import numpy as np

cv2.setNumThreads(1)
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]
_libm.lrintf.restype = ctypes.c_long
_libm.lrintf.argtypes = [ctypes.c_float]

F = np.float32
EDGE = 19
HALF_PATCH = 15
PATCH = 31


def cv_round(v):
    """cvRound: round half to even (SSE cvtss2si)."""
    return int(np.rint(v))


def load_pattern():
    import os
    import re
    txt = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "orb_pattern.inc")).read()
    body = "\n".join(l for l in txt.splitlines() if not l.lstrip().startswith("//"))
    v = np.array([int(t) for t in re.findall(r"-?\d+", body)], np.int32)
    assert v.size == 1024
    return v.reshape(512, 2)


class Cv2Extractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sf = float(F(scale_factor))  # ctor takes float, member is double (ORBextractor.h:91)
        # ORBextractor.cc:411-425
        self.scale = [F(1.0)]
        for i in range(1, nlevels):
            self.scale.append(F(float(self.scale[-1]) * sf))
        self.inv_scale = [F(1.0) / s for s in self.scale]
        # ORBextractor.cc:429-441
        factor = F(1.0 / sf)
        want = F(nfeatures) * (F(1) - factor) / (F(1) - F(math.pow(float(factor), float(nlevels))))
        self.quota, s = [], 0
        for _ in range(nlevels - 1):
            self.quota.append(cv_round(want))
            s += self.quota[-1]
            want = F(want * factor)
        self.quota.append(max(nfeatures - s, 0))
        # ORBextractor.cc:449-463
        vmax = int(math.floor(float(F(HALF_PATCH) * np.sqrt(F(2)) / F(2) + F(1))))
        vmin = int(math.ceil(float(F(HALF_PATCH) * np.sqrt(F(2)) / F(2))))
        um = [0] * (HALF_PATCH + 1)
        for v in range(vmax + 1):
            um[v] = cv_round(math.sqrt(HALF_PATCH * HALF_PATCH - v * v))
        v0 = 0
        for v in range(HALF_PATCH, vmin - 1, -1):
            while um[v0] == um[v0 + 1]:
                v0 += 1
            um[v] = v0
            v0 += 1
        self.umax = um
        self.pattern = load_pattern()
        self.pyr = []      # padded planes
        self.trace = []

    # ORBextractor.cc:1041-1065
    def compute_pyramid(self, image):
        self.pyr = []
        h, w = image.shape
        for l in range(self.nlevels):
            s = self.inv_scale[l]
            lw, lh = cv_round(F(w) * s), cv_round(F(h) * s)
            if l == 0:
                lvl = image
            else:
                prev = self.pyr[l - 1][EDGE:-EDGE, EDGE:-EDGE]
                lvl = cv2.resize(prev, (lw, lh), interpolation=cv2.INTER_LINEAR)
            self.pyr.append(cv2.copyMakeBorder(lvl, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101))

    def level(self, l):
        return self.pyr[l][EDGE:-EDGE, EDGE:-EDGE]

    # ORBextractor.cc:733-799
    def detect_level(self, l):
        img = self.level(l)
        rows, cols = img.shape
        min_bx = min_by = EDGE - 3
        max_bx, max_by = cols - EDGE + 3, rows - EDGE + 3
        width, height = F(max_bx - min_bx), F(max_by - min_by)
        n_cols, n_rows = int(width / F(30)), int(height / F(30))
        w_cell, h_cell = int(math.ceil(width / F(n_cols))), int(math.ceil(height / F(n_rows)))
        det_ini = cv2.FastFeatureDetector_create(self.ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        det_min = cv2.FastFeatureDetector_create(self.min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        out = []
        for i in range(n_rows):
            ini_y = min_by + i * h_cell
            max_y = ini_y + h_cell + 6
            if ini_y >= max_by - 3:
                continue
            max_y = min(max_y, max_by)
            for j in range(n_cols):
                ini_x = min_bx + j * w_cell
                max_x = ini_x + w_cell + 6
                if ini_x >= max_bx - 6:
                    continue
                max_x = min(max_x, max_bx)
                cell = np.ascontiguousarray(img[ini_y:max_y, ini_x:max_x])
                kps = det_ini.detect(cell)
                if len(kps) == 0:
                    kps = det_min.detect(cell)
                for k in kps:
                    out.append((int(k.pt[0]) + j * w_cell, int(k.pt[1]) + i * h_cell, int(k.response)))
        return out, (min_bx, max_bx, min_by, max_by)

    # ORBextractor.cc:475-731; ties in (count, node*) resolved by creation order (later first)
    @staticmethod
    def distribute(cands, min_x, max_x, min_y, max_y, n_want):
        n_ini = int(np.round(F(max_x - min_x) / F(max_y - min_y)))  # numpy rounds half-even; see note below
        r = float(F(max_x - min_x) / F(max_y - min_y))
        n_ini = int(math.floor(r + 0.5)) if r >= 0 else -int(math.floor(-r + 0.5))  # C round(): half away
        hx = F(max_x - min_x) / F(n_ini)
        seq = [0]

        def new_node(x0, x1, y0, y1):
            seq[0] += 1
            return {"r": (x0, x1, y0, y1), "k": [], "one": False, "seq": seq[0], "dead": False}

        nodes = [new_node(int(hx * F(i)), int(hx * F(i + 1)), 0, max_y - min_y) for i in range(n_ini)]
        for c in cands:
            nodes[int(F(c[0]) / hx)]["k"].append(c)
        lst = [n for n in nodes if n["k"]]
        for n in lst:
            n["one"] = len(n["k"]) == 1

        def divide(n):
            x0, x1, y0, y1 = n["r"]
            hx_ = int(math.ceil(float(F(x1 - x0) / F(2))))
            hy_ = int(math.ceil(float(F(y1 - y0) / F(2))))
            mx, my = x0 + hx_, y0 + hy_
            ch = [new_node(x0, mx, y0, my), new_node(mx, x1, y0, my), new_node(x0, mx, my, y1), new_node(mx, x1, my, y1)]
            for c in n["k"]:
                q = (0 if c[1] < my else 2) if c[0] < mx else (1 if c[1] < my else 3)
                ch[q]["k"].append(c)
            ch = [c for c in ch if c["k"]]
            for c in ch:
                c["one"] = len(c["k"]) == 1
            return ch

        finish = False
        while not finish:
            prev_size = len(lst)
            expandable, front, keep = [], [], []
            for n in lst:
                if n["one"]:
                    keep.append(n)
                    continue
                for c in divide(n):
                    front.insert(0, c)
                    if len(c["k"]) > 1:
                        expandable.append(c)
            lst = front + keep
            n_to_expand = len(expandable)
            if len(lst) >= n_want or len(lst) == prev_size:
                finish = True
            elif len(lst) + 3 * n_to_expand > n_want:
                while not finish:
                    prev_size = len(lst)
                    prev = sorted(expandable, key=lambda n: (len(n["k"]), n["seq"]))
                    expandable = []
                    for n in reversed(prev):
                        ch = divide(n)
                        n["dead"] = True
                        lst = list(reversed(ch)) + [m for m in lst if not m["dead"]]
                        expandable += [c for c in ch if len(c["k"]) > 1]
                        if len(lst) >= n_want:
                            break
                    if len(lst) >= n_want or len(lst) == prev_size:
                        finish = True
        res = []
        for n in lst:
            best = n["k"][0]
            for c in n["k"][1:]:
                if c[2] > best[2]:
                    best = c
            res.append(best)
        return res

    # ORBextractor.cc:76-100
    def ic_angle(self, l, x, y):
        img = self.pyr[l]
        cy, cx = y + EDGE, x + EDGE
        m01 = m10 = 0
        row = img[cy].astype(np.int64)
        for u in range(-HALF_PATCH, HALF_PATCH + 1):
            m10 += u * int(row[cx + u])
        for v in range(1, HALF_PATCH + 1):
            d = self.umax[v]
            a = img[cy + v, cx - d:cx + d + 1].astype(np.int64)
            b = img[cy - v, cx - d:cx + d + 1].astype(np.int64)
            u = np.arange(-d, d + 1)
            m01 += v * int((a - b).sum())
            m10 += int((u * (a + b)).sum())
        return F(cv2.fastAtan2(float(m01), float(m10)))

    # ORBextractor.cc:102-146
    def describe(self, blurred_padded, x, y, angle):
        factor_pi = F(math.pi / float(F(180.0)))
        ang = F(angle * factor_pi)
        a, b = F(_libm.cosf(float(ang))), F(_libm.sinf(float(ang)))
        px = self.pattern[:, 0].astype(F)
        py = self.pattern[:, 1].astype(F)
        ry = np.rint((px * b).astype(F) + (py * a).astype(F)).astype(np.int64)
        rx = np.rint((px * a).astype(F) - (py * b).astype(F)).astype(np.int64)
        vals = blurred_padded[y + EDGE + ry, x + EDGE + rx].astype(np.int32)
        bits = (vals[0::2] < vals[1::2]).astype(np.uint8).reshape(32, 8)
        return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)

    # ORBextractor.cc:978-1039
    def __call__(self, image, mask=None):
        image = np.ascontiguousarray(image, np.uint8)
        self.compute_pyramid(image)
        self.trace = []
        kps, descs = [], []
        for l in range(self.nlevels):
            cands, (a, b, c, d) = self.detect_level(l)
            sel = self.distribute(cands, a, b, c, d, self.quota[l])
            self.trace.append((cands, sel))
            if not sel:
                continue
            lvl = np.ascontiguousarray(self.level(l))
            blurred = cv2.GaussianBlur(lvl, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            bp = cv2.copyMakeBorder(blurred, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_CONSTANT, value=0)
            patch = int(F(PATCH) * self.scale[l])
            for (x, y, resp) in sel:
                lx, ly = x + EDGE - 3, y + EDGE - 3
                ang = self.ic_angle(l, lx, ly)
                descs.append(self.describe(bp, lx, ly, ang))
                fx, fy = F(lx), F(ly)
                if l != 0:
                    fx, fy = F(fx * self.scale[l]), F(fy * self.scale[l])
                kps.append((fx, fy, F(patch), ang, F(resp), l, -1))
        from oracle import KP_DTYPE  # same record layout as the C++ oracle
        k = np.array(kps, dtype=KP_DTYPE) if kps else np.zeros(0, KP_DTYPE)
        d = np.array(descs, np.uint8).reshape(-1, 32)
        return k, d
