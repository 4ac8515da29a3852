"""Thin Python harness over the C ABI of the NRSfM mapping stages (Schwarp fit, isometric
normals, shape-from-normals) plus the synthetic keyframe-window generator that tests and
bench.py feed to it.

Marshalling only: numpy arrays -> defslam_*_problem structs -> a library that exports the
entry points under a prefix ("defslam_" = the CUDA library, the only product path;
"oracle_" / "emu_" are passed in by tests).  No computation and no fallback here.

Reference conventions followed by the generator:
  keypoint normalisation + spline domain   Modules/Common/DefKeyFrame.cc:94-133
  control grid 13 x 15, valdim 2           Modules/Common/DefKeyFrame.cc:49-56
  invSigma = sqrt(invLevelSigma2[octave])  Modules/Mapping/SchwarpDatabase.cc:181-182
  (fy, fx) handed to Warps::Warp           Modules/Mapping/SchwarpDatabase.cc:199-201
  regularisers 0.05 / 0.7                  scripts/hamlyn_exploration_template.yaml:132-133
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from .sft import DefslamError
from .synthetic import CX, CY, FX, FY, IMG_H, IMG_W, inv_level_sigma2

NCU, NCV = 13, 15
SCHWARP_REG = 0.05
BENDING_REG = 0.7


def make_bbs(umin, umax, vmin, vmax, nptsu=NCU, nptsv=NCV, valdim=2) -> _capi.Bbs:
    b = _capi.Bbs()
    b.umin, b.umax, b.nptsu, b.vmin, b.vmax, b.nptsv, b.valdim = umin, umax, nptsu, vmin, vmax, nptsv, valdim
    return b


def keyframe_domain(q32: np.ndarray):
    """running min/max with the +-0.10 margin, in visit order (DefKeyFrame.cc:116-131)"""
    umin, umax, vmin, vmax = 0.75, -0.75, 0.75, -0.75
    for x, y in q32:
        x, y = float(x), float(y)
        if x < umin:
            umin = float(np.float32(x)) - 0.10
        if x > umax:
            umax = float(np.float32(x)) + 0.10
        if y < vmin:
            vmin = float(np.float32(y)) - 0.10
        if y > vmax:
            vmax = float(np.float32(y)) + 0.10
    return umin, umax, vmin, vmax


# ----------------------------------------------------------------------------- Schwarp --
@dataclass
class SchwarpCase:
    bbs: _capi.Bbs
    kp1: np.ndarray  # [n,2] f32
    kp2: np.ndarray  # [n,2] f32
    inv_sigma: np.ndarray  # [n] f32
    lam: float = SCHWARP_REG
    fx: float = FY  # the reference passes (fy, fx)
    fy: float = FX
    px_fx: float = FX
    px_fy: float = FY
    max_iterations: int = 3
    initialize: int = 1
    x0: np.ndarray | None = None  # [2*NC]

    @property
    def NC(self):
        return self.bbs.nptsu * self.bbs.nptsv

    @property
    def n(self):
        return len(self.kp1)

    def problem(self, x: np.ndarray) -> _capi.SchwarpProblem:
        p = _capi.SchwarpProblem()
        p.bbs = self.bbs
        p.n_matches = self.n
        p.kp1 = _capi.as_ptr(self.kp1, C.c_float)
        p.kp2 = _capi.as_ptr(self.kp2, C.c_float)
        p.inv_sigma = _capi.as_ptr(self.inv_sigma, C.c_float)
        p.lambda_ = self.lam
        p.fx, p.fy, p.px_fx, p.px_fy = self.fx, self.fy, self.px_fx, self.px_fy
        p.max_iterations = self.max_iterations
        p.initialize = self.initialize
        p.x = _capi.as_ptr(x, C.c_double)
        return p


class DiffPropOut:
    def __init__(self, n: int):
        self.warp_uv = np.zeros((n, 2), np.float32)
        self.J12 = np.zeros((n, 4), np.float32)
        self.J21 = np.zeros((n, 4), np.float32)
        self.H12 = np.zeros((n, 6), np.float32)
        self.keep = np.zeros(max(n, 1), np.uint8)
        self.x = None
        d = _capi.DiffProp()
        d.warp_uv = _capi.as_ptr(self.warp_uv, C.c_float)
        d.J12 = _capi.as_ptr(self.J12, C.c_float)
        d.J21 = _capi.as_ptr(self.J21, C.c_float)
        d.H12 = _capi.as_ptr(self.H12, C.c_float)
        d.keep = _capi.as_ptr(self.keep, C.c_uint8)
        self.d = d


# ----------------------------------------------------------------------------- normals --
@dataclass
class NormalsCase:
    pair_ptr: np.ndarray  # [n+1] i32
    J12: np.ndarray
    J21: np.ndarray
    H12: np.ndarray
    I1: np.ndarray
    I2: np.ndarray
    pair_from_ref: np.ndarray  # u8
    k_first: np.ndarray  # f32 [npairs,2]
    k_init: np.ndarray  # f64 [n,2]
    ref_uv: np.ndarray  # f32 [n,2]
    max_iterations: int = 200
    corrected_t2: int = 0

    @property
    def n(self):
        return len(self.pair_ptr) - 1

    @property
    def npairs(self):
        return int(self.pair_ptr[-1])

    def problem(self) -> _capi.NormalsProblem:
        p = _capi.NormalsProblem()
        p.n_points = self.n
        p.pair_ptr = _capi.as_ptr(self.pair_ptr, C.c_int32)
        for name in ("J12", "J21", "H12", "I1", "I2", "k_first", "ref_uv"):
            setattr(p, name, _capi.as_ptr(getattr(self, name), C.c_float))
        p.pair_from_ref = _capi.as_ptr(self.pair_from_ref, C.c_uint8)
        p.k_init = _capi.as_ptr(self.k_init, C.c_double)
        p.max_iterations = self.max_iterations
        p.corrected_t2 = self.corrected_t2
        return p


class NormalsOut:
    def __init__(self, n: int, npairs: int):
        self.k = np.zeros((n, 2))
        self.cov = np.zeros((n, 4))
        self.normal = np.zeros((n, 3), np.float32)
        self.status = np.zeros(max(n, 1), np.uint8)
        self.iters = np.zeros(max(n, 1), np.int32)
        self.pair_normal = np.zeros((max(npairs, 1), 3), np.float32)
        self.pair_valid = np.zeros(max(npairs, 1), np.uint8)

    def args(self):
        return (_capi.as_ptr(self.k, C.c_double), _capi.as_ptr(self.cov, C.c_double),
                _capi.as_ptr(self.normal, C.c_float), _capi.as_ptr(self.status, C.c_uint8),
                _capi.as_ptr(self.iters, C.c_int32), _capi.as_ptr(self.pair_normal, C.c_float),
                _capi.as_ptr(self.pair_valid, C.c_uint8))


# ----------------------------------------------------------------------------- SfN ------
@dataclass
class SfnCase:
    bbs: _capi.Bbs  # valdim 1
    uv: np.ndarray  # f32 [n,2]
    normals: np.ndarray  # f32 [n,3]
    eval_uv: np.ndarray  # f32 [m,2]
    bending: float = BENDING_REG
    mean_depth: float = 1.0
    ctrl: np.ndarray = field(default=None)
    xyz: np.ndarray = field(default=None)

    @property
    def NC(self):
        return self.bbs.nptsu * self.bbs.nptsv

    def problem(self) -> _capi.SfnProblem:
        if self.ctrl is None:  # kept alive (and shared) across problem() calls: the struct points into them
            self.ctrl = np.zeros(self.NC)
            self.xyz = np.zeros((max(len(self.eval_uv), 1), 3), np.float32)
        p = _capi.SfnProblem()
        p.bbs = self.bbs
        p.n_normals = len(self.uv)
        p.uv = _capi.as_ptr(self.uv, C.c_float)
        p.normals = _capi.as_ptr(self.normals, C.c_float)
        p.bending, p.mean_depth = self.bending, self.mean_depth
        p.n_eval = len(self.eval_uv)
        p.eval_uv = _capi.as_ptr(self.eval_uv, C.c_float)
        p.ctrl_out = _capi.as_ptr(self.ctrl, C.c_double)
        p.xyz_out = _capi.as_ptr(self.xyz, C.c_float)
        return p


# ----------------------------------------------------------------------------- Sim(3) ---
@dataclass
class Sim3Case:
    pts1: np.ndarray  # f32 [n,3] surface points (world frame)
    pts2: np.ndarray  # f32 [n,3] stored map-point positions
    scale: float = 1.0
    rot: tuple = (0.0, 0.0, 0.0, 1.0)
    trans: tuple = (0.0, 0.0, 0.0)
    chi: float = 0.07 ** 2
    huber: float = 0.01
    max_iterations: int = 50

    def __post_init__(self):
        self.pts1 = np.ascontiguousarray(self.pts1, dtype=np.float32)
        self.pts2 = np.ascontiguousarray(self.pts2, dtype=np.float32)

    def problem(self) -> _capi.Sim3Problem:
        p = _capi.Sim3Problem()
        p.n_points = len(self.pts1)
        p.pts1 = _capi.as_ptr(self.pts1, C.c_float)
        p.pts2 = _capi.as_ptr(self.pts2, C.c_float)
        p.rot = (C.c_double * 4)(*self.rot)
        p.trans = (C.c_double * 3)(*self.trans)
        p.scale, p.chi, p.huber, p.max_iterations = self.scale, self.chi, self.huber, self.max_iterations
        return p


# ----------------------------------------------------------------------------- API ------
class Api:
    """The NRSfM entry points of one library under one symbol prefix."""

    def __init__(self, lib=None, prefix: str = "defslam_"):
        self.lib = lib if lib is not None else _capi.load()
        self.prefix = prefix
        if prefix != "defslam_":
            P = _capi.PROTOTYPES
            for name in ("schwarp_fit", "schwarp_evaluate", "schwarp_initial", "normals_batched", "polysolver_coefficients",
                         "sfn_solve", "sfn_system", "schwarp_fit_batched", "sfn_solve_batched",
                         "sim3_register_batched", "scale_min_median", "new_map_points"):
                if hasattr(self.lib, prefix + name):
                    f = getattr(self.lib, prefix + name)
                    f.restype, f.argtypes = P["defslam_" + name]
            if hasattr(self.lib, prefix + "schwarp_init"):
                f = getattr(self.lib, prefix + "schwarp_init")
                f.restype, f.argtypes = C.c_int, [C.POINTER(_capi.SchwarpProblem), _capi.c_double_p]

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, name, rc):
        if rc != 0:
            raise DefslamError(self.prefix + name, rc)

    # -- Schwarp
    def schwarp_evaluate(self, case: SchwarpCase, x: np.ndarray, jac: bool = True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        NR, NP = 2 * case.n + 4 * case.NC, 2 * case.NC
        r = np.zeros(NR)
        J = np.zeros((NR, NP)) if jac else None
        p = case.problem(x)
        self._check("schwarp_evaluate", self._f("schwarp_evaluate")(
            C.byref(p), _capi.as_ptr(r, C.c_double), _capi.as_ptr(J, C.c_double) if jac else None))
        return r, J

    def schwarp_init(self, case: SchwarpCase):
        x = np.zeros(2 * case.NC)
        p = case.problem(x)
        self._check("schwarp_init", self._f("schwarp_init")(C.byref(p), _capi.as_ptr(x, C.c_double)))
        return x

    def schwarp_initial(self, case: SchwarpCase):
        """DefORBmatcher::CalculateInitialSchwarp (DefORBmatcher.cc:111-187): (x0, keep[n], err[n])"""
        x = np.zeros(2 * case.NC)
        keep = np.zeros(case.n, np.uint8)
        err = np.zeros(case.n)
        p = case.problem(x)
        self._check("schwarp_initial", self._f("schwarp_initial")(
            C.byref(p), _capi.as_ptr(keep, C.c_uint8), _capi.as_ptr(err, C.c_double)))
        return x, keep, err

    def schwarp_fit(self, case: SchwarpCase) -> DiffPropOut:
        out = DiffPropOut(case.n)
        out.x = np.zeros(2 * case.NC) if case.x0 is None else np.array(case.x0, dtype=np.float64)
        p = case.problem(out.x)
        self._check("schwarp_fit", self._f("schwarp_fit")(C.byref(p), C.byref(out.d)))
        return out

    def schwarp_prepare(self, cases):
        """Descriptors over the host arrays, built once (HostBatch-style): calling the returned
        function is exactly one C-ABI call."""
        n = len(cases)
        outs = [DiffPropOut(c.n) for c in cases]
        probs = (_capi.SchwarpProblem * n)()
        dps = (_capi.DiffProp * n)()
        x0 = []
        for i, (c, o) in enumerate(zip(cases, outs)):
            o.x = np.zeros(2 * c.NC) if c.x0 is None else np.array(c.x0, dtype=np.float64)
            x0.append(o.x.copy())
            probs[i] = c.problem(o.x)
            dps[i] = o.d
        fn = self._f("schwarp_fit_batched")

        def call(device: int = -1):
            for o, x in zip(outs, x0):   # x is in/out
                o.x[:] = x
            self._check("schwarp_fit_batched", fn(n, probs, dps, device))
            for i, o in enumerate(outs):
                o.d = dps[i]
            return outs
        call.keepalive = (cases, outs, probs, dps)
        return call

    def schwarp_fit_batched(self, cases, device: int = -1):
        return self.schwarp_prepare(cases)(device)

    # -- normals
    def polysolver_coefficients(self, J12, H12, I1, I2):
        n = len(J12)
        J12, H12, I1, I2 = (np.ascontiguousarray(a, dtype=np.float32) for a in (J12, H12, I1, I2))
        e1, e2 = np.zeros((n, 10)), np.zeros((n, 10))
        self._check("polysolver_coefficients", self._f("polysolver_coefficients")(
            n, _capi.as_ptr(J12, C.c_float), _capi.as_ptr(H12, C.c_float), _capi.as_ptr(I1, C.c_float),
            _capi.as_ptr(I2, C.c_float), _capi.as_ptr(e1, C.c_double), _capi.as_ptr(e2, C.c_double)))
        return e1, e2

    def normals_prepare(self, case: NormalsCase):
        out = NormalsOut(case.n, case.npairs)
        p = case.problem()
        args = out.args()
        fn = self._f("normals_batched")

        def call():
            self._check("normals_batched", fn(C.byref(p), *args))
            return out
        call.keepalive = (case, out, p, args)
        return call

    def normals(self, case: NormalsCase) -> NormalsOut:
        return self.normals_prepare(case)()

    # -- shape from normals
    def sfn_system(self, case: SfnCase):
        rows = 2 * len(case.uv) + case.NC + 1
        A, b = np.zeros((rows, case.NC)), np.zeros(rows)
        p = case.problem()
        self._check("sfn_system", self._f("sfn_system")(C.byref(p), _capi.as_ptr(A, C.c_double),
                                                        _capi.as_ptr(b, C.c_double)))
        return A, b

    def sfn_solve(self, case: SfnCase):
        p = case.problem()
        self._check("sfn_solve", self._f("sfn_solve")(C.byref(p)))
        return case.ctrl, case.xyz

    def sfn_prepare(self, cases):
        n = len(cases)
        probs = (_capi.SfnProblem * n)()
        for i, c in enumerate(cases):
            probs[i] = c.problem()
        rcs = np.zeros(n, np.int32)
        fn = self._f("sfn_solve_batched")

        def call(device: int = -1):
            self._check("sfn_solve_batched", fn(n, probs, _capi.as_ptr(rcs, C.c_int32), device))
            return rcs
        call.keepalive = (cases, probs, rcs)
        return call

    def sfn_solve_batched(self, cases, device: int = -1):
        return self.sfn_prepare(cases)(device)


    # -- Sim(3) registration
    def sim3_register(self, cases, device: int = -1):
        n = len(cases)
        probs = (_capi.Sim3Problem * n)()
        for i, c in enumerate(cases):
            probs[i] = c.problem()
        res = (_capi.Sim3Result * n)()
        self._check("sim3_register_batched", self._f("sim3_register_batched")(n, probs, res, device))
        return [dict(rot=np.array(r.rot[:]), trans=np.array(r.trans[:]), scale=r.scale, chi2=r.chi2, inliers=r.inliers,
                     acceptable=r.acceptable, iterations=tuple(r.iterations[:])) for r in res]

    def scale_min_median(self, mono, stereo, seed: int = 1):
        mono = np.ascontiguousarray(mono, np.float32)
        stereo = np.ascontiguousarray(stereo, np.float32)
        out = C.c_float(0)
        self._check("scale_min_median", self._f("scale_min_median")(
            len(mono), _capi.as_ptr(mono, C.c_float), _capi.as_ptr(stereo, C.c_float), C.c_uint64(seed),
            C.cast(C.byref(out), _capi.c_float_p)))
        return out.value


def _new_map_points(self, kp_xy, kp_state, rows: int, cols: int, surf_xyz=None, T_wc=None):
    """DefLocalMapping::CreateNewMapPoints / needNewTemplate (DefLocalMapping.cc:240-347,355-403):
    returns (action[n], world_xyz[n,3] or None, n_new)."""
    kp_xy = np.ascontiguousarray(kp_xy, np.float32)
    kp_state = np.ascontiguousarray(kp_state, np.uint8)
    n = len(kp_state)
    place = surf_xyz is not None
    if place:
        surf_xyz = np.ascontiguousarray(surf_xyz, np.float32)
        T_wc = np.ascontiguousarray(T_wc, np.float32)
    p = _capi.NewPointsProblem(n, rows, cols, _capi.as_ptr(kp_xy, C.c_float), _capi.as_ptr(kp_state, C.c_uint8),
                               _capi.as_ptr(surf_xyz, C.c_float) if place else None,
                               _capi.as_ptr(T_wc, C.c_float) if place else None)
    action = np.zeros(max(n, 1), np.uint8)
    world = np.zeros((max(n, 1), 3), np.float32) if place else None
    n_new = C.c_int32(0)
    self._check("new_map_points", self._f("new_map_points")(
        C.byref(p), _capi.as_ptr(action, C.c_uint8), _capi.as_ptr(world, C.c_float) if place else None,
        C.cast(C.byref(n_new), _capi.c_int32_p)))
    return action[:n], (world[:n] if place else None), n_new.value


Api.new_map_points = _new_map_points


def sim3_case(seed: int, n: int = 600, noise: float = 0.003, outlier_frac: float = 0.05) -> Sim3Case:
    """surface cloud vs map cloud related by a similarity close to identity (what NRSfM registers)"""
    rng = np.random.default_rng(seed)
    P = np.stack([rng.uniform(-0.8, 0.8, n), rng.uniform(-0.6, 0.6, n), rng.uniform(0.8, 1.3, n)], 1)
    R = _rot(rng.normal(size=3), np.deg2rad(rng.uniform(1, 5)))
    s, t = rng.uniform(0.7, 1.4), rng.normal(size=3) * 0.05
    Q = s * (R @ P.T).T + t + rng.normal(size=P.shape) * noise
    bad = rng.uniform(size=n) < outlier_frac
    Q[bad] += rng.normal(size=(int(bad.sum()), 3)) * 0.3
    return Sim3Case(pts1=P.astype(np.float32), pts2=Q.astype(np.float32), scale=float(s * rng.uniform(0.9, 1.1)))


# ----------------------------------------------------------------------------- synthetic
def _rot(axis, angle):
    axis = np.asarray(axis, float)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


def surface_depth(u, v, amp=0.05):
    return 1.0 + amp * np.sin(2.0 * u) * np.cos(2.5 * v)


@dataclass
class KeyframeWindow:
    """One reference keyframe observed again from n_views later keyframes."""
    q1: np.ndarray  # [N,2] f32 normalised keypoints of the reference keyframe
    octave: np.ndarray
    bbs2: _capi.Bbs  # warp spline (valdim 2) on the reference keyframe's domain
    bbs1: _capi.Bbs  # depth spline (valdim 1), same domain
    views: list  # per view: dict(idx=[n] indices into q1, q2=[n,2] f32)
    X1: np.ndarray  # ground truth 3-D points in the reference camera
    normals_gt: np.ndarray  # ground-truth (k1,k2)


def make_window(seed: int, n_keypoints: int = 1200, n_views: int = 4, match_frac: float = 0.5, noise_px: float = 0.3,
                nptsu: int = NCU, nptsv: int = NCV) -> KeyframeWindow:
    rng = np.random.default_rng(seed)
    px = np.stack([rng.uniform(8, IMG_W - 8, n_keypoints), rng.uniform(8, IMG_H - 8, n_keypoints)], 1)
    q1 = np.stack([(px[:, 0] - CX) / FX, (px[:, 1] - CY) / FY], 1).astype(np.float32)
    octave = rng.integers(0, 6, n_keypoints)
    umin, umax, vmin, vmax = keyframe_domain(q1)
    u, v = q1[:, 0].astype(float), q1[:, 1].astype(float)
    d = surface_depth(u, v)
    X1 = np.stack([u * d, v * d, d], 1)
    # analytic normal parameters k = -grad(d)/d  (n ~ (k1, k2, 1 - k1 u - k2 v))
    du = 0.05 * 2.0 * np.cos(2.0 * u) * np.cos(2.5 * v)
    dv = -0.05 * 2.5 * np.sin(2.0 * u) * np.sin(2.5 * v)
    normals_gt = np.stack([-du / d, -dv / d], 1)
    views = []
    for k in range(n_views):
        R = _rot(rng.normal(size=3), np.deg2rad(rng.uniform(2.0, 6.0)))
        t = rng.normal(size=3) * 0.04
        # mild non-rigid bending on top of the rigid motion
        bend = 0.01 * np.sin(3.0 * u + rng.uniform(0, 6.28))
        X2 = (R @ (X1 + np.stack([0 * u, 0 * u, bend], 1)).T).T + t
        q2 = X2[:, :2] / X2[:, 2:3]
        q2 = q2 + rng.normal(size=q2.shape) * noise_px / FX
        sel = np.sort(rng.choice(n_keypoints, int(match_frac * n_keypoints), replace=False))
        views.append(dict(idx=sel.astype(np.int32), q2=q2[sel].astype(np.float32)))
    return KeyframeWindow(q1=q1, octave=octave, bbs2=make_bbs(umin, umax, vmin, vmax, nptsu, nptsv, 2),
                          bbs1=make_bbs(umin, umax, vmin, vmax, nptsu, nptsv, 1), views=views, X1=X1,
                          normals_gt=normals_gt)


def schwarp_cases(win: KeyframeWindow):
    isig = np.sqrt(inv_level_sigma2()).astype(np.float32)
    return [SchwarpCase(bbs=win.bbs2, kp1=np.ascontiguousarray(win.q1[vw["idx"]]), kp2=np.ascontiguousarray(vw["q2"]),
                        inv_sigma=np.ascontiguousarray(isig[win.octave[vw["idx"]]])) for vw in win.views]


def normals_case(win: KeyframeWindow, fits) -> NormalsCase:
    """CSR by map point of the kept DiffProp records of all views (WarpDatabase contents)."""
    N = len(win.q1)
    per_point = [[] for _ in range(N)]
    for vi, (vw, f) in enumerate(zip(win.views, fits)):
        for j, pi in enumerate(vw["idx"]):
            if f.keep[j]:
                per_point[pi].append((vi, j))
    ptr = np.zeros(N + 1, np.int32)
    rows = []
    for i in range(N):
        rows.extend(per_point[i])
        ptr[i + 1] = len(rows)
    npairs = len(rows)
    J12 = np.zeros((npairs, 4), np.float32); J21 = np.zeros((npairs, 4), np.float32)
    H12 = np.zeros((npairs, 6), np.float32); I1 = np.zeros((npairs, 2), np.float32); I2 = np.zeros((npairs, 2), np.float32)
    for r, (vi, j) in enumerate(rows):
        f, vw = fits[vi], win.views[vi]
        J12[r], J21[r], H12[r] = f.J12[j], f.J21[j], f.H12[j]
        I1[r], I2[r] = win.q1[vw["idx"][j]], vw["q2"][j]
    return NormalsCase(pair_ptr=ptr, J12=J12, J21=J21, H12=H12, I1=I1, I2=I2,
                       pair_from_ref=np.ones(max(npairs, 1), np.uint8),
                       k_first=np.full((max(npairs, 1), 2), np.nan, np.float32), k_init=np.zeros((N, 2)),
                       ref_uv=np.ascontiguousarray(win.q1))


def sfn_case(win: KeyframeWindow, nout: NormalsOut) -> SfnCase:
    ok = (nout.status[:len(win.q1)] == 1) & np.all(np.isfinite(nout.normal), 1)
    return SfnCase(bbs=win.bbs1, uv=np.ascontiguousarray(win.q1[ok]), normals=np.ascontiguousarray(nout.normal[ok]),
                   eval_uv=np.ascontiguousarray(win.q1))
