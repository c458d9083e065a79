"""TEST INFRASTRUCTURE — ctypes view of oracle/_ref/lib{ref,shim}_slam*.so (oracle/ref_shim/slam_api.cpp) and a
synthetic map generator for the ORBmatcher parity tests.

`SlamLib(path)` wraps one build: `ref` = the reference's own ORBmatcher.cc, `shim_cpu` = shim/ORBmatcher.cc on the CPU
oracle backend, `shim_gpu` = shim/ORBmatcher.cc on liborbgpu.so.  `World` draws a deterministic scene (3-D points with
base descriptors, cameras, noisy observations, clutter, vocabulary-node labels) and `World.build(lib)` replays it into
a library, so that two builds see byte-identical inputs.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])

_f, _i, _p = C.c_float, C.c_int, C.c_void_p


def lib_path(kind):
    return os.path.join(REF_DIR, {"ref": "libref_slam.so", "shim_cpu": "libshim_slam_cpu.so", "shim_gpu": "libshim_slam_gpu.so"}[kind])


def available(kind):
    return os.path.exists(lib_path(kind))


def _ptr(a):
    return a.ctypes.data_as(_p) if a is not None else None


class SlamLib:
    def __init__(self, kind):
        self.kind = kind
        self.L = C.CDLL(lib_path(kind))
        L = self.L
        L.slam_scene_create.restype = _p
        L.slam_scene_create.argtypes = [_f] * 8 + [_i, _f]
        L.slam_scene_destroy.argtypes = [_p]
        L.slam_add_frame.argtypes = [_p, _i, _p, _p, _p, _p, _f, _f, _p]
        L.slam_frame_set_pose.argtypes = [_p, _i, _p]
        L.slam_add_keyframe.argtypes = [_p, _i]
        L.slam_add_mappoint.argtypes = [_p, _p, _i]
        L.slam_observe.argtypes = [_p, _i, _i, _i]
        L.slam_mappoint_update.argtypes = [_p, _i]
        L.slam_mappoint_set_bad.argtypes = [_p, _i]
        L.slam_mappoint_set_descriptor.argtypes = [_p, _i, _p]
        L.slam_mappoint_distinctive.argtypes = [_p, _i, _p, _i, _p]
        L.slam_mappoint_descriptor.argtypes = [_p, _i, _p]
        L.slam_frame_assoc.argtypes = [_p, _i, _i, _i]
        L.slam_frame_outlier.argtypes = [_p, _i, _i, _i]
        L.slam_set_featvec.argtypes = [_p, _i, _i, _i, _p, _p, _p]
        L.slam_frame_mappoints.argtypes = [_p, _i, _p]
        L.slam_kf_mappoints.argtypes = [_p, _i, _p]
        L.slam_mappoint_state.argtypes = [_p, _i, _p]
        L.slam_mappoint_index_in_kf.argtypes = [_p, _i, _i]
        L.slam_frame_grid.argtypes = [_p, _i, _p, _p]
        L.slam_features_in_area.argtypes = [_p, _i, _i, _f, _f, _f, _i, _i, _p, _i]
        L.slam_search_local_points.argtypes = [_p, _i, _p, _i, _f, _f, _i, _p]
        L.slam_search_last_frame.argtypes = [_p, _i, _i, _f, _i, _f, _i]
        L.slam_search_reloc.argtypes = [_p, _i, _i, _p, _i, _f, _i, _f, _i]
        L.slam_search_sim3_projection.argtypes = [_p, _i, _p, _p, _i, _p, _i, _f]
        L.slam_search_bow_frame.argtypes = [_p, _i, _i, _p, _f, _i]
        L.slam_search_bow_kf.argtypes = [_p, _i, _i, _p, _f, _i]
        L.slam_search_initialization.argtypes = [_p, _i, _i, _p, _p, _i, _f, _i]
        L.slam_search_triangulation.argtypes = [_p, _i, _i, _p, _i, _p, _i, _f, _i]
        L.slam_fuse.argtypes = [_p, _i, _p, _i, _f, _f]
        L.slam_fuse_sim3.argtypes = [_p, _i, _p, _p, _i, _f, _p, _f]
        L.slam_search_by_sim3.argtypes = [_p, _i, _i, _p, _f, _p, _p, _f, _f]
        L.slam_descriptor_distance.argtypes = [_p, _p]
        if hasattr(L, "slam_compute_stereo"):       # absent from the build whose ORBextractor is the drop-in (no ComputePyramid member)
            L.slam_compute_stereo.argtypes = [_i, _f, _i, _i, _i, _p, _p, _i, _i, _p, _p, _i, _p, _p, _i, _f, _f, _p, _p]

    def stereo_frame(self, params, left, right, fx, fy, cx, cy, bf, th_depth=35.0, cap=20000):
        """The reference's stereo Frame constructor (Frame.cc:66-127) on this build's ORBextractor / ORBmatcher."""
        nf, sf, nl, ini, mn = params
        left, right = np.ascontiguousarray(left), np.ascontiguousarray(right)
        keys, desc = np.zeros(cap, KP), np.zeros((cap, 32), np.uint8)
        u, z, u2, z2 = (np.full(cap, np.nan, np.float32) for _ in range(4))
        start, items = np.zeros(64 * 48 + 1, np.int32), np.zeros(cap, np.int32)
        self.L.slam_stereo_frame.argtypes = [_i, _f, _i, _i, _i, _p, _p, _i, _i, _f, _f, _f, _f, _f, _f, _i] + [_p] * 8
        n = self.L.slam_stereo_frame(nf, sf, nl, ini, mn, _ptr(left), _ptr(right), left.shape[1], left.shape[0], fx, fy, cx, cy, bf, th_depth, cap,
                                     _ptr(keys), _ptr(desc), _ptr(u), _ptr(z), _ptr(start), _ptr(items), _ptr(u2), _ptr(z2))
        assert 0 <= n <= cap
        return dict(keys=keys[:n], desc=desc[:n], u_right=u[:n], depth=z[:n], cell_start=start, cell_items=items[:start[-1]],
                    u_right_device=u2[:n], depth_device=z2[:n])

    def compute_stereo(self, params, left, right, kL, dL, kR, dR, mbf, mb):
        nf, sf, nl, ini, mn = params
        left, right = np.ascontiguousarray(left), np.ascontiguousarray(right)
        kL, kR = np.ascontiguousarray(kL), np.ascontiguousarray(kR)
        dL, dR = np.ascontiguousarray(dL), np.ascontiguousarray(dR)
        u, z = np.empty(len(kL), np.float32), np.empty(len(kL), np.float32)
        n = self.L.slam_compute_stereo(nf, sf, nl, ini, mn, _ptr(left), _ptr(right), left.shape[1], left.shape[0], _ptr(kL), _ptr(dL), len(kL),
                                       _ptr(kR), _ptr(dR), len(kR), mbf, mb, _ptr(u), _ptr(z))
        return u, z, n


KITTI = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, w=1241, h=376, bf=386.1448)


def _pose(rx, ry, rz, t):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T.astype(np.float32)


class View:
    """One camera's observation of the world: keypoints, descriptors, right coordinates, node labels."""
    pass


class World:
    def __init__(self, seed, npoints=1400, clutter=500, cam=KITTI, nlevels=8, scale_factor=1.2, stereo_fraction=0.7, flips=40, nodes=90):
        self.rng = np.random.default_rng(seed)
        self.cam, self.nlevels, self.sf = cam, nlevels, scale_factor
        rng = self.rng
        self.P = np.stack([rng.uniform(-22, 22, npoints), rng.uniform(-5, 4, npoints), rng.uniform(4, 45, npoints)], 1).astype(np.float32)
        self.base_desc = rng.integers(0, 256, (npoints, 32), dtype=np.uint8)
        self.base_angle = rng.uniform(0, 360, npoints).astype(np.float32)
        self.node_of_point = rng.integers(0, nodes, npoints)
        self.nodes, self.clutter, self.stereo_fraction, self.flips = nodes, clutter, stereo_fraction, flips
        self.scales = np.float32(scale_factor) ** np.arange(nlevels, dtype=np.float32)

    def observe(self, Tcw, roll_deg=0.0, angle_outliers=0.08, noise_px=0.7):
        """Returns a View of the world from pose Tcw (4x4 float32)."""
        rng, cam = self.rng, self.cam
        Pc = (Tcw[:3, :3].astype(np.float64) @ self.P.T.astype(np.float64)).T + Tcw[:3, 3].astype(np.float64)
        z = Pc[:, 2]
        ok = z > 1.0
        u = cam["fx"] * Pc[:, 0] / np.where(ok, z, 1) + cam["cx"]
        v = cam["fy"] * Pc[:, 1] / np.where(ok, z, 1) + cam["cy"]
        ok &= (u > 20) & (u < cam["w"] - 20) & (v > 20) & (v < cam["h"] - 20)
        ids = np.nonzero(ok)[0]
        ids = ids[rng.random(len(ids)) < 0.9]
        n_pts, n = len(ids), len(ids) + self.clutter
        dist = np.linalg.norm(Pc[ids], axis=1)
        level = np.clip(np.round(np.log(dist / 5.0) / np.log(self.sf) + rng.normal(0, 0.4, n_pts)), 0, self.nlevels - 1).astype(np.int32)
        level = (self.nlevels - 1 - level).clip(0, self.nlevels - 1)            # near points are seen at coarse levels
        kp = np.zeros(n, KP)
        desc = np.empty((n, 32), np.uint8)
        src = np.full(n, -1, np.int64)
        uR = np.full(n, -1, np.float32)
        depth = np.full(n, -1, np.float32)
        node = np.empty(n, np.int64)
        # world points
        s = self.scales[level]
        x = np.round((u[ids] + rng.normal(0, noise_px, n_pts)) / s) * s          # ORB keypoints sit on level pixels
        y = np.round((v[ids] + rng.normal(0, noise_px, n_pts)) / s) * s
        kp["x"][:n_pts], kp["y"][:n_pts] = x, y
        kp["octave"][:n_pts] = level
        ang = self.base_angle[ids] + roll_deg + rng.normal(0, 4, n_pts)
        wrong = rng.random(n_pts) < angle_outliers
        ang[wrong] = rng.uniform(0, 360, wrong.sum())
        kp["angle"][:n_pts] = np.mod(ang, 360)
        d = self.base_desc[ids].copy()
        bits = np.unpackbits(d, axis=1)
        nflip = rng.integers(0, self.flips + 1, n_pts)
        for i in range(n_pts):
            bits[i, rng.choice(256, nflip[i], replace=False)] ^= 1
        desc[:n_pts] = np.packbits(bits, axis=1)
        src[:n_pts] = ids
        stereo = rng.random(n_pts) < self.stereo_fraction
        disp = cam["bf"] / z[ids]
        uR[:n_pts] = np.where(stereo, x - disp + rng.normal(0, 0.3, n_pts), -1).astype(np.float32)
        depth[:n_pts] = np.where(stereo, z[ids], -1).astype(np.float32)
        node[:n_pts] = np.where(rng.random(n_pts) < 0.9, self.node_of_point[ids], rng.integers(0, self.nodes, n_pts))
        # clutter
        c = slice(n_pts, n)
        lv = rng.integers(0, self.nlevels, self.clutter).astype(np.int32)
        sc = self.scales[lv]
        kp["x"][c] = np.round(rng.uniform(20, cam["w"] - 20, self.clutter) / sc) * sc
        kp["y"][c] = np.round(rng.uniform(20, cam["h"] - 20, self.clutter) / sc) * sc
        kp["octave"][c] = lv
        kp["angle"][c] = rng.uniform(0, 360, self.clutter)
        desc[c] = rng.integers(0, 256, (self.clutter, 32), dtype=np.uint8)
        node[c] = rng.integers(0, self.nodes, self.clutter)
        kp["size"] = 31 * self.scales[kp["octave"]]
        kp["response"] = rng.integers(7, 120, n)
        kp["class_id"] = -1
        perm = rng.permutation(n)
        V = View()
        V.kp, V.desc, V.src, V.uR, V.depth, V.node, V.Tcw = kp[perm], np.ascontiguousarray(desc[perm]), src[perm], uR[perm], depth[perm], node[perm], Tcw
        return V


class Scene:
    """A World replayed into one SlamLib: frames, keyframes and MapPoints with the same ids in every build."""

    def __init__(self, lib, world):
        self.lib, self.L, self.world = lib, lib.L, world
        cam = world.cam
        self.h = self.L.slam_scene_create(cam["fx"], cam["fy"], cam["cx"], cam["cy"], 0.0, float(cam["w"]), 0.0, float(cam["h"]), world.nlevels, world.sf)
        self.views, self.kf_of_frame, self.n_mps = [], {}, 0
        self.mp_of_point = {}

    def close(self):
        if self.h:
            self.L.slam_scene_destroy(self.h)
            self.h = None

    def add_frame(self, V, with_pose=True):
        Tcw = np.ascontiguousarray(V.Tcw, np.float32)
        f = self.L.slam_add_frame(self.h, len(V.kp), _ptr(V.kp), _ptr(V.uR), _ptr(V.depth), _ptr(V.desc), self.world.cam["bf"], 35.0,
                                  _ptr(Tcw) if with_pose else None)
        self.views.append(V)
        return f

    def add_keyframe(self, f):
        k = self.L.slam_add_keyframe(self.h, f)
        self.kf_of_frame[f] = k
        return k

    def set_featvec(self, is_kf, idx, V):
        order = np.argsort(V.node, kind="stable")
        nodes, counts = np.unique(V.node, return_counts=True)
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        nodes = nodes.astype(np.int32)
        indices = order.astype(np.int32)
        self.L.slam_set_featvec(self.h, int(is_kf), idx, len(nodes), _ptr(nodes), _ptr(offsets), _ptr(indices))

    def add_mappoints(self, observations, update=True):
        """observations: dict world point id -> list of (kf, keypoint idx); the first keyframe is the reference one."""
        for pid in sorted(observations):
            obs = observations[pid]
            pos = np.ascontiguousarray(self.world.P[pid], np.float32)
            mp = self.L.slam_add_mappoint(self.h, _ptr(pos), obs[0][0])
            for kf, idx in obs:
                self.L.slam_observe(self.h, mp, kf, idx)
            if update:
                self.L.slam_mappoint_update(self.h, mp)
            # the reference picks between equally good descriptors by KeyFrame address: pin the reference observation's
            d = np.ascontiguousarray(self.views[obs[0][0]].desc[obs[0][1]])
            self.L.slam_mappoint_set_descriptor(self.h, mp, _ptr(d))
            self.mp_of_point[pid] = mp
            self.n_mps = mp + 1

    # ---- dumps
    def frame_mappoints(self, f):
        out = np.empty(len(self.views[f].kp), np.int32)
        self.L.slam_frame_mappoints(self.h, f, _ptr(out))
        return out

    def kf_mappoints(self, k, n):
        out = np.empty(n, np.int32)
        self.L.slam_kf_mappoints(self.h, k, _ptr(out))
        return out

    def mappoint_states(self):
        out = np.empty((self.n_mps, 3), np.int32)
        for i in range(self.n_mps):
            self.L.slam_mappoint_state(self.h, i, _ptr(out[i]))
        return out

    def mappoint_descriptors(self):
        out = np.empty((self.n_mps, 32), np.uint8)
        for i in range(self.n_mps):
            self.L.slam_mappoint_descriptor(self.h, i, _ptr(out[i]))
        return out

    def frame_grid(self, f):
        start = np.empty(64 * 48 + 1, np.int32)
        items = np.empty(len(self.views[f].kp), np.int32)
        self.L.slam_frame_grid(self.h, f, _ptr(start), _ptr(items))
        return start, items[:start[-1]]

    def features_in_area(self, is_kf, idx, x, y, r, lo=-1, hi=-1):
        out = np.empty(4096, np.int32)
        n = self.L.slam_features_in_area(self.h, int(is_kf), idx, x, y, r, lo, hi, _ptr(out), len(out))
        return out[:n].copy()


def observations_from_views(views_kf, min_views=1):
    """Common world points of the keyframe views -> {point: [(kf, idx), ...]} (kf numbering = position in views_kf)."""
    obs = {}
    for k, V in enumerate(views_kf):
        for idx, pid in enumerate(V.src):
            if pid >= 0:
                obs.setdefault(int(pid), []).append((k, idx))
    return {p: o for p, o in obs.items() if len(o) >= min_views}
