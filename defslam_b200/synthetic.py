"""Synthetic Mandala-shaped SfT workloads.

Generator side only: it produces the *inputs* (template constants, matches,
observations) that tests and bench.py feed to the C ABI and to the oracle.  The
NumPy template builder here is an independent restatement used to generate data and
to cross-check the oracle; it is not on the product path (the product's mesh
Laplacian is the CUDA kernel behind ``defslam_mesh_laplacian``).

Reference conventions followed:
  intrinsics                     scripts/stereo0_template.yaml:11-14
  node layout                    Modules/Mapping/Surface.cc:125-161 (u outer, v inner, fp32)
  triangulation                  Modules/Template/TriangularMesh.cc:92-107
  mean-value weights / kappa0    Modules/Template/LaplacianMesh.cc:53-148
  barycentric embedding (fp32)   Modules/Template/TriangularMesh.cc:133-236
  invSigma2 per octave (fp32)    Thirdparty/ORBSLAM_2/src/ORBextractor.cc:416-431
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi

FX = FY = 435.2047
CX, CY = 367.4517, 252.2009
IMG_W, IMG_H = 640, 480

# BASELINE.json configs -> concrete shapes
CONFIGS = {
    "C1": dict(G=9, M=300, B=1, max_iterations=50, cfg=1),
    "C2": dict(G=13, M=1000, B=1, max_iterations=10, cfg=2),
    "C3": dict(G=17, M=600, B=1, max_iterations=50, cfg=3),
    "C4": dict(G=10, M=400, B=256, max_iterations=50, cfg=4),
    "C5": dict(G=25, M=2000, B=64, max_iterations=50, cfg=5),
}


def regular_triangulation(nv: int, nh: int) -> np.ndarray:
    f = []
    for j in range(nh - 1):
        for i in range(nv - 1):
            f.append((i + nh * j, i + nh * j + 1, nh * (j + 1) + i))
            f.append((i + nh * j + 1, nh * (j + 1) + i, nh * (j + 1) + i + 1))
    return np.asarray(f, dtype=np.int32)


@dataclass
class MeshTemplate:
    nodes_rest: np.ndarray  # (n,3) f64
    facets: np.ndarray  # (nf,3) i32
    nbr_ptr: np.ndarray
    nbr_idx: np.ndarray
    nbr_w: np.ndarray
    boundary: np.ndarray
    kappa0: np.ndarray
    edge_ab: np.ndarray
    edge_len0: np.ndarray
    edge_median_len: float
    uv: np.ndarray | None = None  # (n,2) normalised coords of the nodes (generator only)
    G: int = 0
    _desc: object = field(default=None, repr=False)

    @property
    def n_nodes(self) -> int:
        return self.nodes_rest.shape[0]

    def desc(self) -> _capi.TemplateDesc:
        if self._desc is None:
            d = _capi.TemplateDesc()
            d.n_nodes = self.n_nodes
            d.n_edges = self.edge_ab.shape[0]
            d.n_facets = self.facets.shape[0]
            d.node_rest_xyz = _capi.as_ptr(self.nodes_rest, C.c_double)
            d.node_boundary = _capi.as_ptr(self.boundary, C.c_uint8)
            d.nbr_ptr = _capi.as_ptr(self.nbr_ptr, C.c_int32)
            d.nbr_idx = _capi.as_ptr(self.nbr_idx, C.c_int32)
            d.nbr_w = _capi.as_ptr(self.nbr_w, C.c_double)
            d.node_kappa0 = _capi.as_ptr(self.kappa0, C.c_double)
            d.edge_ab = _capi.as_ptr(self.edge_ab, C.c_int32)
            d.edge_len0 = _capi.as_ptr(self.edge_len0, C.c_double)
            d.facets = _capi.as_ptr(self.facets, C.c_int32)
            d.edge_median_len = float(self.edge_median_len)
            self._desc = d
        return self._desc


def build_template(nodes: np.ndarray, facets: np.ndarray, uv=None, G=0) -> MeshTemplate:
    """NumPy/Python restatement of the template constants (edges, 1-ring, mean-value
    weights, boundary flags, kappa0, median edge length)."""
    X = np.ascontiguousarray(nodes, dtype=np.float64)
    n = X.shape[0]
    nbrs = [set() for _ in range(n)]
    edges, seen = [], set()
    for f in facets:
        v1, v2, v3 = (int(t) for t in f)
        for a, b in ((v1, v2), (v2, v3), (v1, v3)):
            key = (min(a, b), max(a, b))
            if key in seen:
                continue
            seen.add(key)
            edges.append(key)
            nbrs[a].add(b)
            nbrs[b].add(a)
    edge_ab = np.asarray(edges, dtype=np.int32).reshape(-1, 2)
    d = X[edge_ab[:, 0]] - X[edge_ab[:, 1]]
    edge_len0 = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2)
    median = float(np.sort(edge_len0)[len(edge_len0) // 2]) if len(edge_len0) else 0.10
    ring = [sorted(s) for s in nbrs]
    boundary = np.zeros(n, dtype=np.uint8)
    w = [dict() for _ in range(n)]
    for i in range(n):
        Ni = X[i]
        for j in ring[i]:
            common = [c for c in ring[j] if c in nbrs[i]]
            if len(common) == 0:
                continue
            if len(common) == 1:
                boundary[j] = 1
                continue
            Nj, Nj1, Nj_1 = X[j], X[common[0]], X[common[1]]
            t1 = np.linalg.norm(np.cross(Nj_1 - Ni, Nj - Ni)) / np.dot(Nj_1 - Ni, Nj - Ni)
            t2 = np.linalg.norm(np.cross(Nj1 - Ni, Nj - Ni)) / np.dot(Nj1 - Ni, Nj - Ni)
            w[i][j] = (np.tan(abs(np.arctan(t1)) / 2) + np.tan(abs(np.arctan(t2)) / 2)) / np.linalg.norm(Ni - Nj)
    kappa0 = np.zeros(n)
    for i in range(n):
        if boundary[i] or len(ring[i]) <= 1:
            continue
        L = np.zeros(3)
        sw = 0.0
        for j in ring[i]:
            wij = w[i].get(j, 0.0)
            L = L + wij * X[j]
            sw = sw + wij
        kappa0[i] = np.linalg.norm(X[i] - L / sw)
    nbr_ptr = np.zeros(n + 1, dtype=np.int32)
    for i in range(n):
        nbr_ptr[i + 1] = nbr_ptr[i] + len(ring[i])
    nbr_idx = np.asarray([j for r in ring for j in r], dtype=np.int32)
    nbr_w = np.asarray([w[i].get(j, 0.0) for i in range(n) for j in ring[i]], dtype=np.float64)
    return MeshTemplate(X, np.ascontiguousarray(facets, dtype=np.int32), nbr_ptr, nbr_idx, nbr_w, boundary, kappa0,
                        edge_ab, np.ascontiguousarray(edge_len0), median, uv, G)


def image_domain(fx=FX, fy=FY, cx=CX, cy=CY, w=IMG_W, h=IMG_H):
    """BBS/mesh domain from the image corners -/+ 0.10 (DefKeyFrame.cc:116-131)."""
    umin, umax = (0 - cx) / fx - 0.10, (w - cx) / fx + 0.10
    vmin, vmax = (0 - cy) / fy - 0.10, (h - cy) / fy + 0.10
    return umin, umax, vmin, vmax


def template_surface_depth(u, v):
    return 1.0 + 0.05 * np.sin(2 * np.pi * u) * np.cos(2 * np.pi * v)


def make_template(G: int, fx=FX, fy=FY, cx=CX, cy=CY) -> MeshTemplate:
    umin, umax, vmin, vmax = image_domain(fx, fy, cx, cy)
    t = 0.03
    xs = np.arange(G, dtype=np.float64)
    U = (umax - umin - 2 * t) * xs / (G - 1) + (umin + t)
    V = (vmax - vmin - 2 * t) * xs / (G - 1) + (vmin + t)
    uu, vv = np.meshgrid(U, V, indexing="ij")  # node index = x*G + j, u outer
    uu, vv = uu.reshape(-1), vv.reshape(-1)
    d = template_surface_depth(uu, vv)
    nodes32 = np.stack([uu * d, vv * d, d], axis=1).astype(np.float32)  # cv::Mat CV_32F, T_wc = I
    nodes = nodes32.astype(np.float64)
    return build_template(nodes, regular_triangulation(G, G), uv=np.stack([uu, vv], 1), G=G)


def inv_level_sigma2(nlevels=6, scale=1.2) -> np.ndarray:
    sf = np.ones(nlevels, dtype=np.float32)
    for i in range(1, nlevels):
        sf[i] = np.float32(sf[i - 1] * np.float32(scale))
    return (np.float32(1.0) / (sf * sf)).astype(np.float32)


def _point_in_triangle32(q, p0, p1, p2):
    """fp32 restatement of TriangularMesh::pointInTriangle, vectorised over points."""
    f = np.float32
    u, v, w = (p1 - p0).astype(f), (p2 - p0).astype(f), (q - p0).astype(f)
    n = np.cross(u, v).astype(f)
    nn = np.einsum("ij,ij->i", n, n).astype(f)
    gamma = (np.einsum("ij,ij->i", np.cross(u, w).astype(f), n).astype(f) / nn).astype(f)
    beta = (np.einsum("ij,ij->i", np.cross(w, v).astype(f), n).astype(f) / nn).astype(f)
    alpha = (f(1) - gamma - beta).astype(f)
    newp = (p0 * alpha[:, None] + p1 * beta[:, None] + p2 * gamma[:, None]).astype(f)
    d2 = np.einsum("ij,ij->i", newp - q, newp - q)
    ok = (d2 <= 1e-1) & (alpha >= 0) & (alpha <= 1) & (beta >= 0) & (beta <= 1) & (gamma >= 0) & (gamma <= 1)
    return ok, np.stack([alpha, beta, gamma], 1)


def embed_points(tmpl_nodes: np.ndarray, facets: np.ndarray, pts32: np.ndarray):
    """calculateFeaturesCoordinates: closest node, then its incident facets in index order
    (vectorised over the points: slot s of every point's incident-facet list at a time)."""
    n = tmpl_nodes.shape[0]
    npts = len(pts32)
    inc = [[] for _ in range(n)]
    for fi, f in enumerate(facets):
        for v in f:
            inc[int(v)].append(fi)
    max_inc = max((len(x) for x in inc), default=0)
    inc_tab = np.full((n, max(max_inc, 1)), -1, dtype=np.int64)
    for v, lst in enumerate(inc):
        inc_tab[v, :len(lst)] = lst
    dist = np.sqrt(((tmpl_nodes[None, :, :] - pts32[:, None, :].astype(np.float64)) ** 2).sum(-1))
    closest = dist.argmin(1)
    near = dist[np.arange(npts), closest] < 100
    out_f = np.full(npts, -1, dtype=np.int32)
    out_nodes = np.full((npts, 3), -1, dtype=np.int32)
    out_b = np.zeros((npts, 3), dtype=np.float32)
    nodes32 = tmpl_nodes.astype(np.float32)
    fsorted = np.sort(facets, axis=1)
    todo = near.copy()
    for s in range(max_inc):
        cand = inc_tab[closest, s]
        sel = np.flatnonzero(todo & (cand >= 0))
        if len(sel) == 0:
            continue
        v = fsorted[cand[sel]]
        ok, b = _point_in_triangle32(pts32[sel], nodes32[v[:, 0]], nodes32[v[:, 1]], nodes32[v[:, 2]])
        hit = sel[ok]
        out_f[hit] = cand[hit]
        out_nodes[hit] = v[ok]
        out_b[hit] = b[ok]
        todo[hit] = False
    return out_f, out_nodes, out_b


def _rodrigues(axis, angle):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


@dataclass
class SftFrame:
    template: MeshTemplate
    node_xyz: np.ndarray  # (n,3) f64 current (initial) nodes
    match_nodes: np.ndarray  # (M,3) i32
    match_bary: np.ndarray  # (M,3) f64
    match_uv: np.ndarray  # (M,2) f32
    match_inv_sigma2: np.ndarray  # (M,) f32
    T_cw: np.ndarray  # (4,4) f32
    n_frame_keypoints: int = 1200
    fx: float = FX
    fy: float = FY
    cx: float = CX
    cy: float = CY
    reg_lap: float = 700.0
    reg_inex: float = 12000.0
    reg_temp: float = 0.05
    neighbour_layers: int = 2
    max_iterations: int = 50
    matches_given: int = 0        # 1: the matches-given overload (DefOptimizer.h:58-61)
    curv_edge_len: float = 0.0    # its lenghtEdge_ (quirk C8)
    gt_nodes: np.ndarray | None = None
    gt_T_cw: np.ndarray | None = None
    is_gross_outlier: np.ndarray | None = None

    @property
    def n_matches(self) -> int:
        return self.match_nodes.shape[0]

    def problem(self, tmpl_handle=None) -> _capi.SftProblem:
        p = _capi.SftProblem()
        p.tmpl = tmpl_handle
        p.tmpl_desc = C.pointer(self.template.desc()) if tmpl_handle is None else None
        p.node_xyz = _capi.as_ptr(self.node_xyz, C.c_double)
        p.n_matches = self.n_matches
        p.n_frame_keypoints = self.n_frame_keypoints
        p.match_nodes = _capi.as_ptr(self.match_nodes, C.c_int32)
        p.match_bary = _capi.as_ptr(self.match_bary, C.c_double)
        p.match_uv = _capi.as_ptr(self.match_uv, C.c_float)
        p.match_inv_sigma2 = _capi.as_ptr(self.match_inv_sigma2, C.c_float)
        p.fx, p.fy, p.cx, p.cy = self.fx, self.fy, self.cx, self.cy
        for i, v in enumerate(np.asarray(self.T_cw, dtype=np.float32).reshape(-1)):
            p.T_cw[i] = float(v)
        p.reg_lap, p.reg_inex, p.reg_temp = self.reg_lap, self.reg_inex, self.reg_temp
        p.neighbour_layers = self.neighbour_layers
        p.max_iterations = self.max_iterations
        p.matches_given = self.matches_given
        p.curv_edge_len = self.curv_edge_len
        return p


def make_frame(template: MeshTemplate, M: int, seed: int, max_iterations=50, noise_px=1.0, outlier_frac=0.05,
               amp=0.03, shear=0.01, rot_deg=2.0, trans=0.02, n_frame_keypoints=1200) -> SftFrame:
    """One synthetic frame: matches uniform in the image on the rest template, observations
    from a deformed template seen by a slightly moved camera."""
    rng = np.random.default_rng(seed)
    G = template.G
    nodes = template.nodes_rest
    uvn = template.uv
    U = uvn[::G, 0]  # u of grid row x
    V = uvn[:G, 1]
    fx, fy, cx, cy = FX, FY, CX, CY
    mn, mb, px_keep = [], [], []
    need = M
    while need > 0:
        k = int(need * 1.5) + 16
        px = np.stack([rng.uniform(0, IMG_W, k), rng.uniform(0, IMG_H, k)], 1)
        un, vn = (px[:, 0] - cx) / fx, (px[:, 1] - cy) / fy
        ix = np.searchsorted(U, un) - 1
        iv = np.searchsorted(V, vn) - 1
        inside = (ix >= 0) & (ix < G - 1) & (iv >= 0) & (iv < G - 1)
        ix, iv, un, vn = ix[inside], iv[inside], un[inside], vn[inside]
        a = (un - U[ix]) / (U[ix + 1] - U[ix])
        b = (vn - V[iv]) / (V[iv + 1] - V[iv])
        # cell nodes: n00=(ix,iv) n01=(ix,iv+1) n10=(ix+1,iv) n11=(ix+1,iv+1); diagonal n01-n10
        n00, n01, n10, n11 = ix * G + iv, ix * G + iv + 1, (ix + 1) * G + iv, (ix + 1) * G + iv + 1
        lower = (a + b) <= 1.0
        # the ray through (un,vn) hits the planar facet; barycentrics w.r.t. the projected
        # triangle are perspective-distorted, so intersect in 3D instead.
        tri = np.where(lower[:, None], np.stack([n00, n01, n10], 1), np.stack([n01, n10, n11], 1))
        P0, P1, P2 = nodes[tri[:, 0]], nodes[tri[:, 1]], nodes[tri[:, 2]]
        nrm = np.cross(P1 - P0, P2 - P0)
        ray = np.stack([un, vn, np.ones_like(un)], 1)
        s = np.einsum("ij,ij->i", nrm, P0) / np.einsum("ij,ij->i", nrm, ray)
        pts32 = (ray * s[:, None]).astype(np.float32)
        f_id, e_nodes, e_bary = embed_points(nodes, template.facets, pts32)
        ok = f_id >= 0
        take = np.flatnonzero(ok)[:need]
        mn.append(e_nodes[take])
        mb.append(e_bary[take].astype(np.float64))
        need -= len(take)
    match_nodes = np.ascontiguousarray(np.concatenate(mn), dtype=np.int32)
    match_bary = np.ascontiguousarray(np.concatenate(mb), dtype=np.float64)
    M = match_nodes.shape[0]
    octave = rng.integers(0, 6, M)
    inv_s2 = inv_level_sigma2()[octave].astype(np.float32)

    # ground-truth deformation + camera
    du = U[-1] - U[0]
    phi = rng.uniform(0, 2 * np.pi)
    gt = nodes.copy()
    gt[:, 2] += amp * np.sin(np.pi * uvn[:, 0] / du + phi)
    gt[:, 0] += shear * nodes[:, 1]
    axis = rng.normal(size=3)
    R = _rodrigues(axis, np.deg2rad(rot_deg))
    tdir = rng.normal(size=3)
    t = trans * tdir / np.linalg.norm(tdir)
    T_gt = np.eye(4)
    T_gt[:3, :3], T_gt[:3, 3] = R, t
    Pw = (match_bary[:, :, None] * gt[match_nodes]).sum(1)
    Pc = Pw @ R.T + t
    uv = np.stack([fx * Pc[:, 0] / Pc[:, 2] + cx, fy * Pc[:, 1] / Pc[:, 2] + cy], 1)
    uv += rng.normal(0, noise_px, uv.shape)
    gross = rng.uniform(size=M) < outlier_frac
    uv[gross] += rng.uniform(-30, 30, (int(gross.sum()), 2))
    return SftFrame(
        template=template,
        node_xyz=np.ascontiguousarray(nodes.copy()),
        match_nodes=match_nodes,
        match_bary=match_bary,
        match_uv=np.ascontiguousarray(uv, dtype=np.float32),
        match_inv_sigma2=np.ascontiguousarray(inv_s2),
        T_cw=np.eye(4, dtype=np.float32),
        n_frame_keypoints=n_frame_keypoints,
        max_iterations=max_iterations,
        gt_nodes=gt,
        gt_T_cw=T_gt,
        is_gross_outlier=gross,
    )


def make_config_frames(name: str, nframes: int | None = None, template: MeshTemplate | None = None):
    """Frames of one BASELINE.json config; seed = 1234 + config*1000 + frame."""
    c = CONFIGS[name]
    tmpl = template or make_template(c["G"])
    nf = c["B"] if nframes is None else nframes
    return tmpl, [make_frame(tmpl, c["M"], 1234 + c["cfg"] * 1000 + i, max_iterations=c["max_iterations"])
                  for i in range(nf)]
