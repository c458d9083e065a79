"""TEST INFRASTRUCTURE — the eleven ORBmatcher searches driven as Tracking / LocalMapping / LoopClosing drive them
(SURVEY.md section 3.3), on a synthetic map, through one build of oracle/ref_shim/slam_api.cpp.

`run_all(kind, seed)` returns {scenario name: {output name: ndarray}}; running it for kind = "ref" (the reference's own
ORBmatcher.cc) and kind = "shim_cpu" / "shim_gpu" (shim/ORBmatcher.cc on the oracle backend / on liborbgpu.so) and
comparing the dictionaries is the parity test for the matcher row of SURVEY.md section 8a.
"""
import numpy as np

import slam_ref
from slam_ref import _pose, _ptr


def _k(cam):
    return np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]], np.float64)


def _f12(T1, T2, cam):
    """LocalMapping::ComputeF12 (LocalMapping.cc:538-556): F12 = K1^-T [t12]x R12 K2^-1."""
    R1, t1, R2, t2 = T1[:3, :3].astype(np.float64), T1[:3, 3].astype(np.float64), T2[:3, :3].astype(np.float64), T2[:3, 3].astype(np.float64)
    R12 = R1 @ R2.T
    t12 = -R1 @ R2.T @ t2 + t1
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    Ki = np.linalg.inv(_k(cam))
    return (Ki.T @ tx @ R12 @ Ki).astype(np.float32)


class Setup:
    """Frames f0..f3 (views of four poses), keyframes k0 = f0 and k1 = f1, MapPoints for 3/4 of the points seen in them."""

    def __init__(self, kind, seed, small=False):
        self.lib = slam_ref.SlamLib(kind)
        W = self.world = slam_ref.World(seed, npoints=500 if small else 1400, clutter=150 if small else 500)
        self.T = [_pose(0, 0, 0, [0, 0, 0]), _pose(0.004, -0.02, 0.002, [-0.55, 0.02, -0.45]), _pose(0.008, -0.035, 0.004, [-1.0, 0.03, -1.3]),
                  _pose(-0.006, 0.03, 0.0, [0.4, -0.02, 0.9])]
        self.V = [W.observe(T, roll_deg=2.0 * i) for i, T in enumerate(self.T)]
        S = self.S = slam_ref.Scene(self.lib, W)
        self.f = [S.add_frame(V) for V in self.V]
        self.k = [S.add_keyframe(self.f[0]), S.add_keyframe(self.f[1])]
        obs = slam_ref.observations_from_views(self.V[:2])
        rng = np.random.default_rng(seed + 77)
        keep = {p: o for p, o in obs.items() if rng.random() < 0.75}
        S.add_mappoints(keep)
        self.mp_of_point = dict(S.mp_of_point)
        for i, V in enumerate(self.V):
            S.set_featvec(False, self.f[i], V)
        for i in range(2):
            S.set_featvec(True, self.k[i], self.V[i])
        self.rng = rng

    def close(self):
        self.S.close()

    def associate(self, f, fraction, outlier_fraction=0.0):
        """Pre-existing associations of frame f with the MapPoints of its world points (as left by an earlier search)."""
        V, L, h = self.V[f], self.lib.L, self.S.h
        for idx, pid in enumerate(V.src):
            mp = self.mp_of_point.get(int(pid), -1)
            if mp >= 0 and self.rng.random() < fraction:
                L.slam_frame_assoc(h, self.f[f], idx, mp)
                if self.rng.random() < outlier_fraction:
                    L.slam_frame_outlier(h, self.f[f], idx, 1)


def run_all(kind, seed=3, small=False, only=None):
    out = {}

    def want(name):
        return only is None or name in only

    def fresh():
        return Setup(kind, seed, small)

    cam = slam_ref.KITTI
    if want("grid"):
        s = fresh()
        start, items = s.S.frame_grid(s.f[2])
        rng = np.random.default_rng(seed)
        area = []
        for _ in range(200):
            x, y, r = rng.uniform(-30, cam["w"] + 30), rng.uniform(-30, cam["h"] + 30), rng.uniform(1, 80)
            lo = int(rng.integers(-1, 7)); hi = int(rng.choice([-1, lo, lo + 1, lo + 2]))
            area.append(s.S.features_in_area(False, s.f[2], x, y, r, lo, hi))
            area.append(s.S.features_in_area(True, s.k[1], x, y, r))
        out["grid"] = {"start": start, "items": items, "area": np.concatenate(area), "area_len": np.array([len(a) for a in area])}
        s.close()

    for th in (1.0, 3.0):
        name = "local_points_th%g" % th
        if want(name):
            s = fresh()
            s.associate(2, 0.25)
            mps = np.arange(s.S.n_mps, dtype=np.int32)
            in_view = np.empty(len(mps), np.int32)
            n = s.lib.L.slam_search_local_points(s.S.h, s.f[2], _ptr(mps), len(mps), th, 0.8, 1, _ptr(in_view))
            out[name] = {"n": np.array(n), "frame": s.S.frame_mappoints(s.f[2]), "in_view": in_view}
            s.close()

    for cur, last, mono, th in ((2, 1, 0, 7.0), (3, 1, 0, 7.0), (2, 1, 1, 15.0), (1, 0, 0, 15.0)):
        name = "last_frame_c%d_l%d_m%d" % (cur, last, mono)
        if want(name):
            s = fresh()
            s.associate(last, 0.9, outlier_fraction=0.05)
            s.associate(cur, 0.1)
            n = s.lib.L.slam_search_last_frame(s.S.h, s.f[cur], s.f[last], th, mono, 0.9, 1)
            out[name] = {"n": np.array(n), "frame": s.S.frame_mappoints(s.f[cur])}
            s.close()

    for th, orb_dist in ((10.0, 100), (3.0, 64)):
        name = "reloc_th%g" % th
        if want(name):
            s = fresh()
            s.associate(2, 0.3)
            found = np.unique(s.S.frame_mappoints(s.f[2]))
            found = np.ascontiguousarray(found[found >= 0][::2], np.int32)
            n = s.lib.L.slam_search_reloc(s.S.h, s.f[2], s.k[0], _ptr(found), len(found), th, orb_dist, 0.9, 1)
            out[name] = {"n": np.array(n), "frame": s.S.frame_mappoints(s.f[2])}
            s.close()

    if want("sim3_projection"):
        s = fresh()
        Scw = s.T[1].copy()
        Scw[:3, :] *= np.float32(1.04)
        mps = np.arange(s.S.n_mps, dtype=np.int32)[::-1].copy()
        N = len(s.V[1].kp)
        matched = s.S.kf_mappoints(s.k[1], N)
        matched[::3] = -1
        matched = np.where(np.arange(N) % 2 == 0, matched, -1).astype(np.int32)
        n = s.lib.L.slam_search_sim3_projection(s.S.h, s.k[1], _ptr(np.ascontiguousarray(Scw)), _ptr(mps), len(mps), _ptr(matched), 10, 0.75)
        out["sim3_projection"] = {"n": np.array(n), "matched": matched}
        s.close()

    for ori in (1, 0):
        name = "bow_frame_ori%d" % ori
        if want(name):
            s = fresh()
            res = np.empty(len(s.V[2].kp), np.int32)
            n = s.lib.L.slam_search_bow_frame(s.S.h, s.k[0], s.f[2], _ptr(res), 0.7, ori)
            out[name] = {"n": np.array(n), "matches": res}
            s.close()

    if want("bow_kf"):
        s = fresh()
        res = np.empty(len(s.V[0].kp), np.int32)
        n = s.lib.L.slam_search_bow_kf(s.S.h, s.k[0], s.k[1], _ptr(res), 0.75, 1)
        out["bow_kf"] = {"n": np.array(n), "matches": res}
        s.close()

    for window in (100, 30):
        name = "initialization_w%d" % window
        if want(name):
            s = fresh()
            prev = np.ascontiguousarray(np.stack([s.V[0].kp["x"], s.V[0].kp["y"]], 1), np.float32)
            m12 = np.empty(len(s.V[0].kp), np.int32)
            n = s.lib.L.slam_search_initialization(s.S.h, s.f[0], s.f[1], _ptr(prev), _ptr(m12), window, 0.9, 1)
            out[name] = {"n": np.array(n), "matches12": m12, "prev": prev}
            s.close()

    for only_stereo, ori in ((0, 0), (1, 0), (0, 1)):
        name = "triangulation_s%d_o%d" % (only_stereo, ori)
        if want(name):
            s = fresh()
            F12 = np.ascontiguousarray(_f12(s.T[0], s.T[1], cam))
            pairs = np.full((4096, 2), -1, np.int32)
            n = s.lib.L.slam_search_triangulation(s.S.h, s.k[0], s.k[1], _ptr(F12), only_stereo, _ptr(pairs), len(pairs), 0.6, ori)
            out[name] = {"n": np.array(n), "pairs": pairs[:max(n, 0)].copy()}
            s.close()

    if want("fuse"):
        s = fresh()
        # extra MapPoints seen only from a third keyframe, so that Fuse into k1 finds both empty and occupied slots
        k2 = s.S.add_keyframe(s.f[2])
        obs = {p: [(k2, i)] for i, p in enumerate(s.V[2].src) if p >= 0 and i % 2 == 0}
        first_new = s.S.n_mps
        # world points already mapped get a SECOND MapPoint here: the duplicates Fuse exists to merge
        for p in sorted(obs):
            pos = np.ascontiguousarray(s.world.P[p], np.float32)
            mp = s.lib.L.slam_add_mappoint(s.S.h, _ptr(pos), k2)
            s.lib.L.slam_observe(s.S.h, mp, k2, obs[p][0][1])
            s.lib.L.slam_mappoint_update(s.S.h, mp)
            s.S.n_mps = mp + 1
        mps = np.concatenate([np.arange(first_new, s.S.n_mps), [-1], np.arange(0, first_new, 5)]).astype(np.int32)
        n = s.lib.L.slam_fuse(s.S.h, s.k[1], _ptr(mps), len(mps), 3.0, 0.6)
        out["fuse"] = {"n": np.array(n), "kf": s.S.kf_mappoints(s.k[1], len(s.V[1].kp)), "states": s.S.mappoint_states()}
        s.close()

    if want("fuse_sim3"):
        s = fresh()
        k2 = s.S.add_keyframe(s.f[2])
        first_new = s.S.n_mps
        for i, p in enumerate(s.V[2].src):
            if p >= 0 and i % 2 == 1:
                pos = np.ascontiguousarray(s.world.P[p], np.float32)
                mp = s.lib.L.slam_add_mappoint(s.S.h, _ptr(pos), k2)
                s.lib.L.slam_observe(s.S.h, mp, k2, i)
                s.lib.L.slam_mappoint_update(s.S.h, mp)
                s.S.n_mps = mp + 1
        Scw = s.T[0].copy()
        Scw[:3, :] *= np.float32(0.97)
        mps = np.arange(first_new, s.S.n_mps, dtype=np.int32)
        rep = np.empty(len(mps), np.int32)
        n = s.lib.L.slam_fuse_sim3(s.S.h, s.k[0], _ptr(np.ascontiguousarray(Scw)), _ptr(mps), len(mps), 4.0, _ptr(rep), 0.8)
        out["fuse_sim3"] = {"n": np.array(n), "replace": rep, "kf": s.S.kf_mappoints(s.k[0], len(s.V[0].kp)), "states": s.S.mappoint_states()}
        s.close()

    if want("by_sim3"):
        s = fresh()
        T0, T1 = s.T[0].astype(np.float64), s.T[1].astype(np.float64)
        T12 = T0 @ np.linalg.inv(T1)
        R12 = np.ascontiguousarray(T12[:3, :3], np.float32)
        t12 = np.ascontiguousarray(T12[:3, 3], np.float32)
        m12 = np.full(len(s.V[0].kp), -1, np.int32)
        seen = s.S.kf_mappoints(s.k[0], len(s.V[0].kp))
        m12[::7] = np.where(seen[::7] >= 0, seen[::7], -1)          # "already matched" entries (vpMatches12 on entry)
        n = s.lib.L.slam_search_by_sim3(s.S.h, s.k[0], s.k[1], _ptr(m12), 1.0, _ptr(R12), _ptr(t12), 7.5, 0.75)
        out["by_sim3"] = {"n": np.array(n), "matches12": m12}
        s.close()
    return out


def compare(a, b):
    """Returns a list of 'scenario.output' names that differ."""
    bad = []
    for name in a:
        for key in a[name]:
            x, y = a[name][key], b[name][key]
            if x.shape != y.shape or not np.array_equal(x, y):
                bad.append("%s.%s" % (name, key))
    return bad
