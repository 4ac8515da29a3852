"""Match production feeding the SfT solve: projection search of the last frame's template points
(DefORBmatcher::SearchByProjection, Modules/Matching/DefORBmatcher.cc:296-451) through the C ABI,
plus a synthetic two-frame generator for tests and timing."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from .sft import DefslamError


@dataclass
class ProjSearchCase:
    last_state: np.ndarray      # u8 [n_last]
    last_has_obs: np.ndarray    # u8 [n_last]
    last_world_xyz: np.ndarray  # f32 [n_last,3]
    last_desc: np.ndarray       # u8 [n_last,32]
    last_octave: np.ndarray     # i32
    last_angle: np.ndarray      # f32
    cur_xy: np.ndarray          # f32 [n_cur,2]
    cur_octave: np.ndarray
    cur_angle: np.ndarray
    cur_desc: np.ndarray
    cur_uright: np.ndarray
    cur_taken: np.ndarray
    scale_factors: np.ndarray
    T_cw: np.ndarray
    T_lw: np.ndarray
    fx: float = 435.2047
    fy: float = 435.2047
    cx: float = 367.4517
    cy: float = 252.2009
    mb: float = 0.0
    mbf: float = 0.0
    width: int = 640
    height: int = 480
    th: float = 15.0
    mono: int = 1
    th_high: int = 75
    check_orientation: int = 1
    truth: np.ndarray = field(default=None)  # generator only: keypoint j observes last-frame point truth[j] (or -1)

    def problem(self) -> _capi.ProjSearchProblem:
        p = _capi.ProjSearchProblem()
        p.n_last, p.n_cur, p.n_levels = len(self.last_state), len(self.cur_octave), len(self.scale_factors)
        for name, ct in (("last_state", C.c_uint8), ("last_has_obs", C.c_uint8), ("last_world_xyz", C.c_float),
                         ("last_desc", C.c_uint8), ("last_octave", C.c_int32), ("last_angle", C.c_float),
                         ("cur_xy", C.c_float), ("cur_octave", C.c_int32), ("cur_angle", C.c_float),
                         ("cur_desc", C.c_uint8), ("cur_uright", C.c_float), ("cur_taken", C.c_uint8),
                         ("scale_factors", C.c_float)):
            setattr(p, name, _capi.as_ptr(getattr(self, name), ct))
        for k, v in enumerate(np.asarray(self.T_cw, np.float32).reshape(-1)):
            p.T_cw[k] = float(v)
        for k, v in enumerate(np.asarray(self.T_lw, np.float32).reshape(-1)):
            p.T_lw[k] = float(v)
        p.fx, p.fy, p.cx, p.cy, p.mb, p.mbf = self.fx, self.fy, self.cx, self.cy, self.mb, self.mbf
        # undistorted image bounds and the 64 x 48 grid (Frame.cc: ComputeImageBounds, mfGridElement*Inv)
        p.min_x, p.max_x, p.min_y, p.max_y = 0.0, float(self.width), 0.0, float(self.height)
        p.grid_width_inv = float(np.float32(64) / (np.float32(self.width) - np.float32(0)))
        p.grid_height_inv = float(np.float32(48) / (np.float32(self.height) - np.float32(0)))
        p.th, p.mono, p.th_high, p.check_orientation = self.th, self.mono, self.th_high, self.check_orientation
        return p


def search_by_projection(case: ProjSearchCase, lib=None, prefix: str = "defslam_"):
    """returns (match[n_cur] -> last-frame keypoint index or -1, nmatches)"""
    lib = lib if lib is not None else _capi.load()
    f = getattr(lib, prefix + "search_by_projection")
    if prefix != "defslam_":
        f.restype, f.argtypes = _capi.PROTOTYPES["defslam_search_by_projection"]
    n_cur = len(case.cur_octave)
    match = np.full(max(n_cur, 1), -1, np.int32)
    nm = C.c_int32(0)
    p = case.problem()
    rc = f(C.byref(p), _capi.as_ptr(match, C.c_int32), C.cast(C.byref(nm), _capi.c_int32_p))
    if rc != 0:
        raise DefslamError(prefix + "search_by_projection", rc)
    return match[:n_cur], nm.value


def make_case(seed: int, n_last: int = 1200, n_clutter: int = 400, flip_bits: int = 24, move_px: float = 3.0,
              stereo: bool = False) -> ProjSearchCase:
    """Two consecutive frames of a surface at depth ~1: the current frame re-observes most map points a few
    pixels away with a few descriptor bits flipped, plus clutter keypoints; some map points are unusable,
    some keypoints are already taken, a few near-duplicate keypoints force the order-dependent choices."""
    rng = np.random.default_rng(seed)
    fx = fy = 435.2047
    cx, cy = 367.4517, 252.2009
    W, H = 640, 480
    nlev = 8
    scale = (1.2 ** np.arange(nlev)).astype(np.float32)
    px = np.stack([rng.uniform(10, W - 10, n_last), rng.uniform(10, H - 10, n_last)], 1)
    z = rng.uniform(0.8, 1.3, n_last)
    Xc_last = np.stack([(px[:, 0] - cx) / fx * z, (px[:, 1] - cy) / fy * z, z], 1)
    T_lw = np.eye(4, dtype=np.float32)
    ang = np.deg2rad(1.0)
    T_cw = np.eye(4, dtype=np.float32)
    T_cw[:3, :3] = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
    T_cw[:3, 3] = np.float32([0.01, -0.005, 0.02 if stereo else 0.002])
    world = Xc_last.astype(np.float32)                 # last camera = world
    last_desc = rng.integers(0, 256, (n_last, 32), dtype=np.uint8)
    last_oct = rng.integers(0, nlev, n_last).astype(np.int32)
    last_ang = rng.uniform(0, 360, n_last).astype(np.float32)
    state = (rng.uniform(size=n_last) < 0.85).astype(np.uint8)
    has_obs = (rng.uniform(size=n_last) < 0.97).astype(np.uint8)
    Pc = world.astype(np.float64) @ T_cw[:3, :3].T.astype(np.float64) + T_cw[:3, 3]
    uv = np.stack([fx * Pc[:, 0] / Pc[:, 2] + cx, fy * Pc[:, 1] / Pc[:, 2] + cy], 1)
    seen = rng.uniform(size=n_last) < 0.8
    idx = np.flatnonzero(seen)
    obs_xy = uv[idx] + rng.normal(0, move_px, (len(idx), 2))
    obs_desc = last_desc[idx].copy()
    for r in range(len(idx)):   # flip a few bits
        bits = rng.choice(256, rng.integers(0, 2 * flip_bits), replace=False)
        for b in bits:
            obs_desc[r, b >> 3] ^= np.uint8(1 << (b & 7))
    obs_oct = np.clip(last_oct[idx] + rng.integers(-1, 2, len(idx)), 0, nlev - 1).astype(np.int32)
    rot = 7.0
    obs_ang = np.mod(last_ang[idx] - rot + rng.normal(0, 2.0, len(idx)), 360).astype(np.float32)
    bad_rot = rng.uniform(size=len(idx)) < 0.05
    obs_ang[bad_rot] = rng.uniform(0, 360, int(bad_rot.sum())).astype(np.float32)
    # near-duplicates: a second keypoint next to some observations with an equally good descriptor
    dup = rng.choice(len(idx), len(idx) // 12, replace=False)
    dup_xy = obs_xy[dup] + rng.normal(0, 1.0, (len(dup), 2))
    clutter_xy = np.stack([rng.uniform(0, W, n_clutter), rng.uniform(0, H, n_clutter)], 1)
    cur_xy = np.concatenate([obs_xy, dup_xy, clutter_xy]).astype(np.float32)
    cur_desc = np.concatenate([obs_desc, obs_desc[dup], rng.integers(0, 256, (n_clutter, 32), dtype=np.uint8)])
    cur_oct = np.concatenate([obs_oct, obs_oct[dup], rng.integers(0, nlev, n_clutter)]).astype(np.int32)
    cur_ang = np.concatenate([obs_ang, obs_ang[dup], rng.uniform(0, 360, n_clutter)]).astype(np.float32)
    truth = np.concatenate([idx, idx[dup], np.full(n_clutter, -1)]).astype(np.int32)
    n_cur = len(cur_xy)
    perm = rng.permutation(n_cur)                     # keypoints come in detector order, not map order
    cur_xy, cur_desc, cur_oct, cur_ang, truth = cur_xy[perm], cur_desc[perm], cur_oct[perm], cur_ang[perm], truth[perm]
    taken = (rng.uniform(size=n_cur) < 0.03).astype(np.uint8)
    uright = np.full(n_cur, -1.0, np.float32)
    mb = mbf = 0.0
    if stereo:
        mbf, mb = 40.0, 40.0 / fx
        has = rng.uniform(size=n_cur) < 0.7
        ok = truth >= 0
        zc = np.where(ok, Pc[np.maximum(truth, 0), 2], 1.0)
        uright = np.where(has, cur_xy[:, 0] - mbf / zc + rng.normal(0, 2.0, n_cur), -1.0).astype(np.float32)
    return ProjSearchCase(
        last_state=state, last_has_obs=has_obs, last_world_xyz=np.ascontiguousarray(world),
        last_desc=np.ascontiguousarray(last_desc), last_octave=last_oct, last_angle=last_ang,
        cur_xy=np.ascontiguousarray(cur_xy), cur_octave=np.ascontiguousarray(cur_oct),
        cur_angle=np.ascontiguousarray(cur_ang), cur_desc=np.ascontiguousarray(cur_desc), cur_uright=uright,
        cur_taken=taken, scale_factors=scale, T_cw=T_cw, T_lw=T_lw, fx=fx, fy=fy, cx=cx, cy=cy, mb=mb, mbf=mbf,
        width=W, height=H, mono=0 if stereo else 1, truth=truth)


# ----------------------------------------------------------------------------- warp-guided search
@dataclass
class WarpSearchCase:
    bbs: _capi.Bbs
    x: np.ndarray            # f64 [2*NC], the reference's layout (NC u-coordinates, then NC v-coordinates)
    kp1_norm: np.ndarray     # f32 [n1,2]
    kp1_state: np.ndarray    # u8
    kp1_desc: np.ndarray     # u8 [n1,32]
    kp2_xy: np.ndarray       # f32 [n2,2]
    kp2_has_mp: np.ndarray   # u8
    kp2_desc: np.ndarray
    fx: float = 435.2047
    fy: float = 435.2047
    cx: float = 367.4517
    cy: float = 252.2009
    width: int = 640
    height: int = 480
    radius: float = 2.0
    th_low: int = 50
    truth: np.ndarray = field(default=None)

    def problem(self) -> _capi.WarpSearchProblem:
        p = _capi.WarpSearchProblem()
        p.bbs = self.bbs
        p.x = _capi.as_ptr(self.x, C.c_double)
        p.n1, p.n2 = len(self.kp1_state), len(self.kp2_has_mp)
        for name, ct in (("kp1_norm", C.c_float), ("kp1_state", C.c_uint8), ("kp1_desc", C.c_uint8),
                         ("kp2_xy", C.c_float), ("kp2_has_mp", C.c_uint8), ("kp2_desc", C.c_uint8)):
            setattr(p, name, _capi.as_ptr(getattr(self, name), ct))
        p.fx, p.fy, p.cx, p.cy = self.fx, self.fy, self.cx, self.cy
        p.min_x, p.max_x, p.min_y, p.max_y = 0.0, float(self.width), 0.0, float(self.height)
        p.grid_width_inv = float(np.float32(64) / np.float32(self.width))
        p.grid_height_inv = float(np.float32(48) / np.float32(self.height))
        p.radius, p.th_low = self.radius, self.th_low
        return p


def search_by_schwarp(case: WarpSearchCase, lib=None, prefix: str = "defslam_"):
    """returns (match12[n1] -> keypoint of keyframe 2 or -1, nmatches)"""
    lib = lib if lib is not None else _capi.load()
    f = getattr(lib, prefix + "search_by_schwarp")
    if prefix != "defslam_":
        f.restype, f.argtypes = _capi.PROTOTYPES["defslam_search_by_schwarp"]
    n1 = len(case.kp1_state)
    match = np.full(max(n1, 1), -1, np.int32)
    nm = C.c_int32(0)
    p = case.problem()
    rc = f(C.byref(p), _capi.as_ptr(match, C.c_int32), C.cast(C.byref(nm), _capi.c_int32_p))
    if rc != 0:
        raise DefslamError(prefix + "search_by_schwarp", rc)
    return match[:n1], nm.value


def make_warp_case(seed: int, n1: int = 1200, n_clutter: int = 500, nptsu: int = 13, nptsv: int = 15) -> WarpSearchCase:
    """keyframe 1 keypoints, a smooth warp to keyframe 2 (control points = identity grid + a low-frequency
    displacement), keyframe 2 keypoints = warped keypoints + sub-pixel noise + clutter + near-duplicates"""
    from . import nrsfm
    rng = np.random.default_rng(seed)
    fx = fy = 435.2047
    cx, cy = 367.4517, 252.2009
    W, H = 640, 480
    px = np.stack([rng.uniform(12, W - 12, n1), rng.uniform(12, H - 12, n1)], 1)
    q1 = np.stack([(px[:, 0] - cx) / fx, (px[:, 1] - cy) / fy], 1).astype(np.float32)
    umin, umax, vmin, vmax = nrsfm.keyframe_domain(q1)
    bbs = nrsfm.make_bbs(umin, umax, vmin, vmax, nptsu, nptsv, 2)
    NC = nptsu * nptsv
    # a cubic B-spline reproduces linear functions from control values sampled at the Greville sites:
    # uniform knots -> control point i sits at umin + (i - 1) * (umax - umin) / (nptsu - 3)
    gu = umin + (np.arange(nptsu) - 1) * (umax - umin) / (nptsu - 3)
    gv = vmin + (np.arange(nptsv) - 1) * (vmax - vmin) / (nptsv - 3)
    GU, GV = np.meshgrid(gu, gv, indexing="ij")
    ctrl_u = GU + 0.02 * np.sin(2.0 * GU) * np.cos(1.5 * GV) + 0.01
    ctrl_v = GV + 0.015 * np.cos(1.7 * GU) - 0.005
    x = np.concatenate([ctrl_u.reshape(-1), ctrl_v.reshape(-1)])
    # warped positions through the library-independent formula are not needed: the test compares libraries;
    # keyframe 2 keypoints are produced by the ORACLE-free closed form of the displacement (close enough to
    # the spline for a 2 px window) plus noise
    u, v = q1[:, 0].astype(float), q1[:, 1].astype(float)
    wu = u + 0.02 * np.sin(2.0 * u) * np.cos(1.5 * v) + 0.01
    wv = v + 0.015 * np.cos(1.7 * u) - 0.005
    p2 = np.stack([wu * fx + cx, wv * fy + cy], 1) + rng.normal(0, 0.5, (n1, 2))
    desc1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = desc1.copy()
    for r in range(n1):
        for b in rng.choice(256, rng.integers(0, 40), replace=False):
            d2[r, b >> 3] ^= np.uint8(1 << (b & 7))
    dup = rng.choice(n1, n1 // 15, replace=False)
    kp2 = np.concatenate([p2, p2[dup] + rng.normal(0, 0.6, (len(dup), 2)),
                          np.stack([rng.uniform(0, W, n_clutter), rng.uniform(0, H, n_clutter)], 1)]).astype(np.float32)
    desc2 = np.concatenate([d2, d2[dup], rng.integers(0, 256, (n_clutter, 32), dtype=np.uint8)])
    truth = np.concatenate([np.arange(n1), dup, np.full(n_clutter, -1)]).astype(np.int32)
    perm = rng.permutation(len(kp2))
    kp2, desc2, truth = kp2[perm], desc2[perm], truth[perm]
    return WarpSearchCase(bbs=bbs, x=np.ascontiguousarray(x), kp1_norm=np.ascontiguousarray(q1),
                          kp1_state=(rng.uniform(size=n1) < 0.8).astype(np.uint8), kp1_desc=np.ascontiguousarray(desc1),
                          kp2_xy=np.ascontiguousarray(kp2), kp2_has_mp=(rng.uniform(size=len(kp2)) < 0.1).astype(np.uint8),
                          kp2_desc=np.ascontiguousarray(desc2), truth=truth)
