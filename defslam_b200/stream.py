"""Tracking + mapping loop over a synthetic stream (config C3 of SURVEY.md 8(d): "Hamlyn f5 phantom
stream, 17x17 grid, NRSfM every 5th keyframe").

The loop has the shape of DefTracking::Track (Modules/Tracking/DefTracking.cc:86-330) and
DefLocalMapping::Run / NRSfM / updateTemplate (Modules/Mapping/DefLocalMapping.cc:100-234):

  every frame        SfT solve from the previous frame's nodes and pose     (DefPoseOptimization)
  every 10th frame   keyframe: the normalised keypoints are kept            (DefTracking.cc:175-178)
  every 5th keyframe Schwarps current keyframe -> previous keyframes, isometric normals, shape from
                     normals, Sim(3) registration onto the stored map points, template rebuilt from
                     the registered surface (Surface::getVertex, mesh Laplacian, point embedding)

Every stage goes through a `Backend`: one library + symbol prefix ("defslam_" = the CUDA library;
tests and bench.py build the same Backend over the CPU oracle to check / time the loop).  This module
never imports the oracle.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import time

import numpy as np

from . import _capi, nrsfm, synthetic

# scripts/hamlyn_exploration_template.yaml:68-75
HAMLYN = dict(fx=755.312744, fy=420.477722, cx=327.875, cy=165.484406, w=720, h=288)


@dataclass
class StreamConfig:
    G: int = 17
    n_points: int = 600
    n_frames: int = 300
    kf_every: int = 10
    nrsfm_every_kf: int = 5
    n_views: int = 4
    seed: int = 4234            # 1234 + config*1000 (C3)
    noise_px: float = 1.0
    outlier_frac: float = 0.05
    amp: float = 0.03
    max_iterations: int = 50
    intr: dict = field(default_factory=lambda: dict(HAMLYN))
    nptsu: int = 17             # "NRSfM (BBS 17x17 => NC=289)"
    nptsv: int = 17
    chi_limit: float = 0.07


class Backend:
    """The entry points of one library under one prefix.  sft_solve: frame -> output with
    .nodes/.T_cw/.outlier/.r (defslam_b200.sft.solve for the CUDA library)."""

    def __init__(self, lib, prefix, sft_solve):
        self.lib, self.prefix, self.sft_solve = lib, prefix, sft_solve
        self.api = nrsfm.Api(lib, prefix)
        P = _capi.PROTOTYPES
        for nm in ("mesh_laplacian", "embed_points", "surface_vertices"):
            f = getattr(lib, prefix + nm)
            f.restype, f.argtypes = P["defslam_" + nm]

    def _f(self, nm):
        return getattr(self.lib, self.prefix + nm)

    def build_template(self, nodes, facets, G):
        """LaplacianMesh constants through the library (defslam_mesh_laplacian)."""
        n, nf, max_ring = nodes.shape[0], facets.shape[0], 8
        nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        facets = np.ascontiguousarray(facets, dtype=np.int32)
        cnt = np.zeros(n, np.int32); idx = np.zeros((n, max_ring), np.int32); w = np.zeros((n, max_ring))
        bd = np.zeros(n, np.uint8); k0 = np.zeros(n); ne = C.c_int32(0)
        ab = np.zeros((3 * nf, 2), np.int32); l0 = np.zeros(3 * nf); med = C.c_double(0)
        rc = self._f("mesh_laplacian")(
            n, _capi.as_ptr(nodes, C.c_double), nf, _capi.as_ptr(facets, C.c_int32), max_ring,
            _capi.as_ptr(cnt, C.c_int32), _capi.as_ptr(idx, C.c_int32), _capi.as_ptr(w, C.c_double),
            _capi.as_ptr(bd, C.c_uint8), _capi.as_ptr(k0, C.c_double), C.cast(C.byref(ne), _capi.c_int32_p),
            _capi.as_ptr(ab, C.c_int32), _capi.as_ptr(l0, C.c_double), C.cast(C.byref(med), _capi.c_double_p))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}mesh_laplacian rc={rc}")
        ptr = np.zeros(n + 1, np.int32)
        ptr[1:] = np.cumsum(cnt)
        nbr_idx = np.concatenate([idx[i, :cnt[i]] for i in range(n)]).astype(np.int32)
        nbr_w = np.concatenate([w[i, :cnt[i]] for i in range(n)]).astype(np.float64)
        E = ne.value
        return synthetic.MeshTemplate(nodes, facets, ptr, np.ascontiguousarray(nbr_idx), np.ascontiguousarray(nbr_w), bd,
                                      k0, np.ascontiguousarray(ab[:E]), np.ascontiguousarray(l0[:E]), med.value, None, G)

    def embed(self, nodes, facets, pts32):
        n, nf, npt = nodes.shape[0], facets.shape[0], pts32.shape[0]
        nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        facets = np.ascontiguousarray(facets, dtype=np.int32)
        pts32 = np.ascontiguousarray(pts32, dtype=np.float32)
        of = np.zeros(npt, np.int32); on = np.zeros((npt, 3), np.int32); ob = np.zeros((npt, 3), np.float32)
        rc = self._f("embed_points")(n, _capi.as_ptr(nodes, C.c_double), nf, _capi.as_ptr(facets, C.c_int32), npt,
                                     _capi.as_ptr(pts32, C.c_float), _capi.as_ptr(of, C.c_int32),
                                     _capi.as_ptr(on, C.c_int32), _capi.as_ptr(ob, C.c_float))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}embed_points rc={rc}")
        return of, on, ob

    def surface_vertices(self, bbs, ctrl, xs, ys):
        out = np.zeros((xs * ys, 3), np.float32)
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        rc = self._f("surface_vertices")(C.byref(bbs), _capi.as_ptr(ctrl, C.c_double), xs, ys, _capi.as_ptr(out, C.c_float))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}surface_vertices rc={rc}")
        return out


def cuda_backend():
    from . import sft
    return Backend(_capi.load(), "defslam_", sft.solve)


# --------------------------------------------------------------------------- ground truth
def _rodrigues(axis, angle):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


class Scene:
    """A surface z = d_t(u, v) over the image domain that bends slowly, and a camera that sways."""

    def __init__(self, cfg: StreamConfig):
        self.cfg = cfg
        it = cfg.intr
        self.dom = synthetic.image_domain(it["fx"], it["fy"], it["cx"], it["cy"], it["w"], it["h"])
        rng = np.random.default_rng(cfg.seed)
        self.axis = rng.normal(size=3)
        self.tdir = rng.normal(size=3)
        self.tdir /= np.linalg.norm(self.tdir)

    def depth(self, u, v, t):
        du = self.dom[1] - self.dom[0]
        return synthetic.template_surface_depth(u, v) + self.cfg.amp * np.sin(np.pi * u / du + 2 * np.pi * t / 120.0) \
            - self.cfg.amp * np.sin(np.pi * u / du)

    def points(self, u, v, t):
        d = self.depth(u, v, t)
        return np.stack([u * d, v * d, d], 1)

    def pose(self, t):
        T = np.eye(4)
        T[:3, :3] = _rodrigues(self.axis, np.deg2rad(2.0) * np.sin(2 * np.pi * t / 200.0))
        T[:3, 3] = 0.02 * np.sin(2 * np.pi * t / 150.0) * self.tdir
        return T


@dataclass
class StreamResult:
    rmse: list            # per frame: node RMSE vs ground truth (camera frame), relative to the RMS node norm
    inliers: list
    trials: list
    nodes_cam: list       # per frame: estimated nodes in the camera frame (float64)
    n_nrsfm: int = 0
    n_template_updates: int = 0
    template_rmse: list = field(default_factory=list)  # per update: new rest nodes vs ground truth, relative
    t_sft: float = 0.0    # seconds inside the per-frame SfT solve (the call DefTracking makes)
    t_nrsfm: float = 0.0  # seconds inside the NRSfM events (Schwarp fits, normals, SfN, registration, template rebuild)


def run_stream(be: Backend, cfg: StreamConfig, keep_nodes: bool = True, logs_dir: str | None = None) -> StreamResult:
    it = cfg.intr
    fx, fy, cx, cy = it["fx"], it["fy"], it["cx"], it["cy"]
    rng = np.random.default_rng(cfg.seed)
    scene = Scene(cfg)
    G = cfg.G
    umin, umax, vmin, vmax = scene.dom
    # material points of the map: uniform in the image, on the surface at t = 0
    px = np.stack([rng.uniform(8, it["w"] - 8, cfg.n_points), rng.uniform(8, it["h"] - 8, cfg.n_points)], 1)
    pu, pv = (px[:, 0] - cx) / fx, (px[:, 1] - cy) / fy
    octave = rng.integers(0, 6, cfg.n_points)
    inv_s2 = synthetic.inv_level_sigma2()[octave].astype(np.float32)
    # first template: the surface at t = 0 sampled like Surface::getVertex (fp32 nodes, camera = world)
    t_in = 0.03
    xs = np.arange(G, dtype=np.float64)
    U = (umax - umin - 2 * t_in) * xs / (G - 1) + (umin + t_in)
    V = (vmax - vmin - 2 * t_in) * xs / (G - 1) + (vmin + t_in)
    uu, vv = (a.reshape(-1) for a in np.meshgrid(U, V, indexing="ij"))
    facets = synthetic.regular_triangulation(G, G)
    node_uv = np.stack([uu, vv], 1)          # material coordinates of the template nodes
    tmpl = be.build_template(scene.points(uu, vv, 0).astype(np.float32).astype(np.float64), facets, G)
    pts_world = scene.points(pu, pv, 0).astype(np.float32)
    f_id, m_nodes, m_bary = be.embed(tmpl.nodes_rest, facets, pts_world)
    valid = f_id >= 0
    nodes = tmpl.nodes_rest.copy()
    T_cw = np.eye(4, dtype=np.float32)
    res = StreamResult([], [], [], [])
    keyframes = []
    log = None
    if logs_dir is not None:
        from . import logs
        log = logs.ResultLogs(logs_dir)
    for t in range(cfg.n_frames):
        Tgt = scene.pose(t)
        Pc = scene.points(pu, pv, t) @ Tgt[:3, :3].T + Tgt[:3, 3]
        uv = np.stack([fx * Pc[:, 0] / Pc[:, 2] + cx, fy * Pc[:, 1] / Pc[:, 2] + cy], 1)
        uv += rng.normal(0, cfg.noise_px, uv.shape)
        gross = rng.uniform(size=cfg.n_points) < cfg.outlier_frac
        uv[gross] += rng.uniform(-30, 30, (int(gross.sum()), 2))
        uv32 = uv.astype(np.float32)
        sel = np.flatnonzero(valid)
        frame = synthetic.SftFrame(
            template=tmpl, node_xyz=np.ascontiguousarray(nodes), match_nodes=np.ascontiguousarray(m_nodes[sel]),
            match_bary=np.ascontiguousarray(m_bary[sel].astype(np.float64)), match_uv=np.ascontiguousarray(uv32[sel]),
            match_inv_sigma2=np.ascontiguousarray(inv_s2[sel]), T_cw=T_cw.copy(), n_frame_keypoints=cfg.n_points,
            fx=fx, fy=fy, cx=cx, cy=cy, max_iterations=cfg.max_iterations)
        t_call = time.perf_counter()
        out = be.sft_solve(frame)
        res.t_sft += time.perf_counter() - t_call
        nodes = np.array(out.nodes, dtype=np.float64)
        T_cw = np.array(out.T_cw, dtype=np.float32).reshape(4, 4)
        Tc = T_cw.astype(np.float64)
        est_cam = nodes @ Tc[:3, :3].T + Tc[:3, 3]
        gt_cam = scene.points(node_uv[:, 0], node_uv[:, 1], t) @ Tgt[:3, :3].T + Tgt[:3, 3]
        res.rmse.append(float(np.sqrt(((est_cam - gt_cam) ** 2).sum(1).mean()) / np.sqrt((gt_cam ** 2).sum(1).mean())))
        res.inliers.append(int(out.r.n_inliers))
        res.trials.append(int(out.r.lm_trials))
        if keep_nodes:
            res.nodes_cam.append(est_cam)
        if log is not None:   # Matches.txt / ErrorGTs<frame>.txt like DefTracking.cc:321-328, GroundTruthFrame.cc:259-264
            n_in = int(out.r.n_inliers)
            log.frame(t, n_in, len(sel) - n_in, int(valid.sum()))
            mp_est = (m_bary[sel][:, :, None].astype(np.float64) * nodes[m_nodes[sel]]).sum(1) @ Tc[:3, :3].T + Tc[:3, 3]
            log.errors(t, np.sqrt(((mp_est - Pc[sel]) ** 2).sum(1)))
        if (t + 1) % cfg.kf_every:
            continue
        # ---- keyframe (normalised keypoints, DefKeyFrame.cc:94-133)
        q = np.stack([(uv32[:, 0] - np.float32(cx)) / np.float32(fx), (uv32[:, 1] - np.float32(cy)) / np.float32(fy)], 1)
        keyframes.append(dict(t=t, q=q.astype(np.float32), ok=~gross, T_cw=T_cw.copy(), nodes=nodes.copy()))
        if len(keyframes) % cfg.nrsfm_every_kf or len(keyframes) <= cfg.n_views:
            continue
        # ---- NRSfM on the current keyframe against the n_views before it
        res.n_nrsfm += 1
        t_call = time.perf_counter()
        new = _nrsfm_template(be, cfg, scene, keyframes, octave, tmpl, m_nodes, m_bary, valid, G)
        res.t_nrsfm += time.perf_counter() - t_call
        if new is None:
            continue
        tmpl, m_nodes, m_bary, valid, node_uv, trel = new
        nodes = tmpl.nodes_rest.copy()
        res.n_template_updates += 1
        res.template_rmse.append(trel)
    if log is not None:
        log.close()
    return res


def _nrsfm_template(be, cfg, scene, keyframes, octave, tmpl, m_nodes, m_bary, valid, G):
    api = be.api
    ref = keyframes[-1]
    q1 = ref["q"]
    umin, umax, vmin, vmax = nrsfm.keyframe_domain(q1)
    bbs2 = nrsfm.make_bbs(umin, umax, vmin, vmax, cfg.nptsu, cfg.nptsv, 2)
    bbs1 = nrsfm.make_bbs(umin, umax, vmin, vmax, cfg.nptsu, cfg.nptsv, 1)
    isig = np.sqrt(synthetic.inv_level_sigma2()).astype(np.float32)
    views = []
    for kf in keyframes[-1 - cfg.n_views:-1]:
        idx = np.flatnonzero(ref["ok"] & kf["ok"]).astype(np.int32)
        views.append(dict(idx=idx, q2=np.ascontiguousarray(kf["q"][idx])))
    win = nrsfm.KeyframeWindow(q1=q1, octave=octave, bbs2=bbs2, bbs1=bbs1, views=views, X1=None, normals_gt=None)
    cases = nrsfm.schwarp_cases(win)
    try:
        fits = api.schwarp_fit_batched(cases) if hasattr(api.lib, api.prefix + "schwarp_fit_batched") \
            else [api.schwarp_fit(c) for c in cases]
        nout = api.normals(nrsfm.normals_case(win, fits))
        scase = nrsfm.sfn_case(win, nout)
        if len(scase.uv) < 30:        # Surface::enoughNormals
            return None
        ctrl, xyz = api.sfn_solve(scase)
    except nrsfm.DefslamError:
        return None
    # ---- registration of the up-to-scale surface onto the stored map points (SurfaceRegistration.cc:48-153)
    Tcw = ref["T_cw"].astype(np.float64)
    Twc = np.linalg.inv(Tcw)
    surf_w = (xyz.astype(np.float64) @ Twc[:3, :3].T + Twc[:3, 3]).astype(np.float32)
    map_w = (m_bary[:, :, None].astype(np.float64) * ref["nodes"][m_nodes]).sum(1).astype(np.float32)
    use = valid & ref["ok"] & np.all(np.isfinite(surf_w), 1)
    if use.sum() < 15:
        return None
    try:
        s0 = api.scale_min_median(surf_w[use], map_w[use], seed=cfg.seed + ref["t"])
        if not (s0 > 0):
            return None
        r = api.sim3_register([nrsfm.Sim3Case(pts1=surf_w[use], pts2=map_w[use], scale=float(s0),
                                              chi=cfg.chi_limit ** 2)])[0]
    except nrsfm.DefslamError:
        return None
    if not r["acceptable"]:
        return None
    x, y, z, w = r["rot"]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    S = lambda P: r["scale"] * (P @ R.T) + r["trans"]
    # ---- new template from the registered surface (Surface::getVertex -> LaplacianMesh)
    verts = be.surface_vertices(bbs1, ctrl, G, G).astype(np.float64)
    nodes_w = S(verts @ Twc[:3, :3].T + Twc[:3, 3]).astype(np.float32).astype(np.float64)
    facets = synthetic.regular_triangulation(G, G)
    new_tmpl = be.build_template(nodes_w, facets, G)
    pts_w = S(surf_w.astype(np.float64)).astype(np.float32)
    f_id, n_nodes, n_bary = be.embed(new_tmpl.nodes_rest, facets, pts_w)
    if (f_id >= 0).sum() < 50:
        return None
    # material coordinates of the new nodes: the keyframe's normalised grid (ground truth bookkeeping only)
    t_in = 0.03
    xs = np.arange(G, dtype=np.float64)
    U = (umax - umin - 2 * t_in) * xs / (G - 1) + (umin + t_in)
    V = (vmax - vmin - 2 * t_in) * xs / (G - 1) + (vmin + t_in)
    uu, vv = (a.reshape(-1) for a in np.meshgrid(U, V, indexing="ij"))
    # the keyframe sees material point (u0, v0) at normalised (u, v): invert by the ground-truth camera at t
    Tgt = scene.pose(ref["t"])
    gt_uv = _material_coords(scene, Tgt, uu, vv, ref["t"])
    gt_cam = scene.points(gt_uv[:, 0], gt_uv[:, 1], ref["t"]) @ Tgt[:3, :3].T + Tgt[:3, 3]
    est_cam = nodes_w @ Tcw[:3, :3].T + Tcw[:3, 3]
    trel = float(np.sqrt(((est_cam - gt_cam) ** 2).sum(1).mean()) / np.sqrt((gt_cam ** 2).sum(1).mean()))
    return new_tmpl, n_nodes, n_bary, f_id >= 0, gt_uv, trel


def _material_coords(scene, Tgt, u_img, v_img, t, iters=20):
    """material (u0, v0) whose point at time t projects to normalised image coords (u_img, v_img)"""
    u0, v0 = u_img.copy(), v_img.copy()
    R, tr = Tgt[:3, :3], Tgt[:3, 3]
    for _ in range(iters):
        P = scene.points(u0, v0, t) @ R.T + tr
        pu, pv = P[:, 0] / P[:, 2], P[:, 1] / P[:, 2]
        u0 += u_img - pu
        v0 += v_img - pv
    return np.stack([u0, v0], 1)
