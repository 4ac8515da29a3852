"""ctypes loader for the CPU oracle.  TEST INFRASTRUCTURE ONLY (see sft_oracle.c):
imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never by
the defslam_b200 package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from defslam_b200 import _capi

_DIR = os.path.dirname(os.path.abspath(__file__))
_lib = None
_ref = None


def build(native: bool = False) -> str:
    target = "_build/liboracle_native.so" if native else "_build/liboracle.so"
    subprocess.run(["make", "-C", _DIR, target], check=True, capture_output=True)
    return os.path.join(_DIR, target)


def build_ref() -> str | None:
    subprocess.run(["make", "-C", _DIR, "ref"], check=True, capture_output=True)
    p = os.path.join(_DIR, "_ref", "libbbs_ref.so")
    return p if os.path.exists(p) else None


def load(native: bool = False):
    global _lib
    if _lib is not None and not native:
        return _lib
    path = os.path.join(_DIR, "_build", "liboracle_native.so" if native else "liboracle.so")
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".c", ".h"))]
    srcs.append(os.path.join(_DIR, "..", "include", "defslam_b200.h"))
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        build(native)
    lib = C.CDLL(path)
    P = _capi
    lib.oracle_sft_solve.restype = C.c_int
    lib.oracle_sft_solve.argtypes = [C.POINTER(P.SftProblem), C.POINTER(P.SftResult)]
    lib.oracle_sft_normal_equations.restype = C.c_int
    lib.oracle_sft_normal_equations.argtypes = [C.POINTER(P.SftProblem), P.c_double_p, P.c_double_p, P.c_double_p]
    lib.oracle_sft_residuals.restype = C.c_int
    lib.oracle_sft_residuals.argtypes = [C.POINTER(P.SftProblem), P.c_double_p, P.c_double_p, C.c_int]
    lib.oracle_sft_residuals_pose.restype = C.c_int
    lib.oracle_sft_residuals_pose.argtypes = [C.POINTER(P.SftProblem), P.c_double_p, P.c_double_p, P.c_double_p,
                                              P.c_double_p, C.c_int]
    lib.oracle_sft_apply_update.restype = C.c_int
    lib.oracle_sft_apply_update.argtypes = [C.POINTER(P.SftProblem), P.c_double_p, P.c_double_p, P.c_float_p,
                                            P.c_double_p, P.c_double_p]
    lib.oracle_mappoints_recalculate.restype = C.c_int
    lib.oracle_mappoints_recalculate.argtypes = [C.c_int32, P.c_double_p, C.c_int32, P.c_int32_p, P.c_double_p,
                                                 P.c_float_p]
    lib.oracle_search_by_projection.restype = C.c_int
    lib.oracle_search_by_projection.argtypes = _capi.PROTOTYPES["defslam_search_by_projection"][1]
    lib.oracle_search_by_schwarp.restype = C.c_int
    lib.oracle_search_by_schwarp.argtypes = _capi.PROTOTYPES["defslam_search_by_schwarp"][1]
    lib.oracle_new_map_points.restype = C.c_int
    lib.oracle_new_map_points.argtypes = _capi.PROTOTYPES["defslam_new_map_points"][1]
    lib.oracle_regular_triangulation.restype = C.c_int
    lib.oracle_regular_triangulation.argtypes = [C.c_int, C.c_int, P.c_int32_p]
    lib.oracle_mesh_laplacian.restype = C.c_int
    lib.oracle_mesh_laplacian.argtypes = _capi.PROTOTYPES["defslam_mesh_laplacian"][1]
    lib.oracle_embed_points.restype = C.c_int
    lib.oracle_embed_points.argtypes = _capi.PROTOTYPES["defslam_embed_points"][1]
    if not native:
        _lib = lib
    return lib


def load_g2o_ref():
    """oracle/_ref/libg2o_sft_ref.so: the reference's OWN SfT code (sft_types.h, se3quat.h, base_*_edge.hpp,
    the Levenberg driver...) behind g2o_ref_harness.cc.  Built only where /root/reference exists; returns None
    when the prebuilt library is absent."""
    global _ref
    if _ref is not None:
        return _ref
    path = os.path.join(_DIR, "_ref", "libg2o_sft_ref.so")
    if not os.path.exists(path):
        return None
    P = _capi
    lib = C.CDLL(path)
    lib.ref_sft_solve.restype = C.c_int
    lib.ref_sft_solve.argtypes = [C.POINTER(P.SftProblem), C.POINTER(P.SftResult)]
    lib.ref_sft_normal_equations.restype = C.c_int
    lib.ref_sft_normal_equations.argtypes = [C.POINTER(P.SftProblem), P.c_double_p, P.c_double_p, P.c_double_p]
    lib.ref_sft_residuals.restype = C.c_int
    lib.ref_sft_residuals.argtypes = [C.POINTER(P.SftProblem), P.c_double_p, P.c_double_p, C.c_int]
    lib.ref_se3_oplus.restype = None
    lib.ref_se3_oplus.argtypes = [P.c_double_p] * 5
    lib.ref_sim3_optimize_horn.restype = C.c_int
    lib.ref_sim3_optimize_horn.argtypes = [C.POINTER(P.Sim3Problem), C.POINTER(P.Sim3Result)]
    lib.ref_mesh_laplacian.restype = C.c_int
    lib.ref_mesh_laplacian.argtypes = [C.c_int32, P.c_double_p, C.c_int32, P.c_int32_p, C.c_int32, P.c_int32_p,
                                       P.c_int32_p, P.c_double_p, P.c_uint8_p, P.c_double_p, P.c_int32_p]
    lib.ref_huber.restype = None
    lib.ref_huber.argtypes = [C.c_float, C.c_double, P.c_double_p]
    _ref = lib
    return lib


def ref_mesh_laplacian(lib, nodes, facets, max_ring=8):
    """LaplacianMesh::ExtractMeanCurvatures (the reference's own lines) -> dict like tests.helpers.mesh_laplacian_call"""
    n, nf = len(nodes), len(facets)
    nodes = np.ascontiguousarray(nodes, np.float64)
    facets = np.ascontiguousarray(facets, np.int32)
    cnt = np.zeros(n, np.int32)
    idx = np.zeros((n, max_ring), np.int32)
    w = np.zeros((n, max_ring))
    bd = np.zeros(n, np.uint8)
    k0 = np.zeros(n)
    nb = C.c_int32(0)
    rc = lib.ref_mesh_laplacian(n, _capi.as_ptr(nodes, C.c_double), nf, _capi.as_ptr(facets, C.c_int32), max_ring,
                                _capi.as_ptr(cnt, C.c_int32), _capi.as_ptr(idx, C.c_int32), _capi.as_ptr(w, C.c_double),
                                _capi.as_ptr(bd, C.c_uint8), _capi.as_ptr(k0, C.c_double),
                                C.cast(C.byref(nb), _capi.c_int32_p))
    return rc, dict(cnt=cnt, idx=idx, w=w, boundary=bd, kappa0=k0, n_bad=nb.value)


def sft_residuals(frame, lib=None, fn="oracle_sft_residuals", jac=True):
    """(res, J) of oracle_sft_residuals / ref_sft_residuals: rows 2*n_rep, 3*n_ref, n_curv, n_str"""
    lib = lib or load()
    f = getattr(lib, fn)
    D = 3 * frame.template.n_nodes + 6
    p = frame.problem()
    rows = f(C.byref(p), None, None, 0)
    res = np.zeros(rows)
    J = np.zeros((rows, D)) if jac else None
    rc = f(C.byref(p), _capi.as_ptr(res, C.c_double), _capi.as_ptr(J, C.c_double) if jac else None, rows)
    if rc != rows:
        raise RuntimeError(f"{fn} rc={rc}")
    return res, J


class SftOutput:
    """numpy-side holder for a defslam_sft_result (works for oracle and product alike)."""

    def __init__(self, n_nodes: int, n_matches: int, trace_capacity: int = 64):
        self.nodes = np.zeros((n_nodes, 3))
        self.outlier = np.zeros(max(n_matches, 1), dtype=np.uint8)
        self.role = np.zeros(n_nodes, dtype=np.uint8)
        self.trace = np.zeros((trace_capacity, 4))
        self.n_matches = n_matches
        r = _capi.SftResult()
        r.node_xyz_out = _capi.as_ptr(self.nodes, C.c_double)
        r.outlier_out = _capi.as_ptr(self.outlier, C.c_uint8)
        r.node_role_out = _capi.as_ptr(self.role, C.c_uint8)
        r.trace = _capi.as_ptr(self.trace, C.c_double)
        r.trace_capacity = trace_capacity
        self.r = r

    @property
    def T_cw(self):
        return np.array(list(self.r.T_cw_out), dtype=np.float32).reshape(4, 4)


def sft_solve(frame, lib=None, fn="oracle_sft_solve"):
    lib = lib or load()
    out = SftOutput(frame.template.n_nodes, frame.n_matches)
    p = frame.problem()
    rc = getattr(lib, fn)(C.byref(p), C.byref(out.r))
    if rc != 0:
        raise RuntimeError(f"{fn} rc={rc}")
    return out


def sft_normal_equations(frame, lib=None, fn="oracle_sft_normal_equations"):
    lib = lib or load()
    D = 3 * frame.template.n_nodes + 6
    H = np.zeros((D, D))
    b = np.zeros(D)
    chi = C.c_double(0)
    p = frame.problem()
    rc = getattr(lib, fn)(C.byref(p), _capi.as_ptr(H, C.c_double), _capi.as_ptr(b, C.c_double),
                          C.cast(C.byref(chi), _capi.c_double_p))
    if rc != 0:
        raise RuntimeError(f"{fn} rc={rc}")
    return H, b, chi.value


# ------------------------------------------------------------------ BBS ------------------
def _bbs_struct(umin, umax, nptsu, vmin, vmax, nptsv, valdim):
    b = _capi.Bbs()
    b.umin, b.umax, b.nptsu, b.vmin, b.vmax, b.nptsv, b.valdim = umin, umax, nptsu, vmin, vmax, nptsv, valdim
    return b


def _bind_bbs(lib, prefix="oracle_"):
    P = _capi
    f = getattr(lib, prefix + "bbs_eval")
    f.restype = C.c_int
    f.argtypes = [C.POINTER(P.Bbs), P.c_double_p, C.c_int32, P.c_double_p, P.c_double_p, C.c_int32, C.c_int32,
                  P.c_double_p]
    f = getattr(lib, prefix + "bbs_coloc")
    f.restype = C.c_int
    f.argtypes = [C.POINTER(P.Bbs), C.c_int32, P.c_double_p, P.c_double_p, C.c_int32, C.c_int32, P.c_double_p]
    f = getattr(lib, prefix + "bbs_bending")
    f.restype = C.c_int
    f.argtypes = [C.POINTER(P.Bbs), P.c_double_p]
    f = getattr(lib, prefix + "surface_vertices")
    f.restype = C.c_int
    f.argtypes = [C.POINTER(P.Bbs), P.c_double_p, C.c_int32, C.c_int32, P.c_float_p]


class BbsApi:
    """eval / coloc / bending / surface_vertices behind one of: the oracle ("oracle_"), the
    emulated kernel sources ("emu_") or the CUDA library ("defslam_")."""

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix
        if prefix != "defslam_":
            _bind_bbs(lib, prefix)

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def eval(self, bbs, ctrl, u, v, du=0, dv=0):
        u = np.ascontiguousarray(u, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        out = np.zeros((len(u), bbs.valdim))
        rc = self._f("bbs_eval")(C.byref(bbs), _capi.as_ptr(ctrl, C.c_double), len(u), _capi.as_ptr(u, C.c_double),
                                 _capi.as_ptr(v, C.c_double), du, dv, _capi.as_ptr(out, C.c_double))
        return rc, out

    def coloc(self, bbs, u, v, du=0, dv=0):
        u = np.ascontiguousarray(u, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.zeros((len(u), bbs.nptsu * bbs.nptsv))
        rc = self._f("bbs_coloc")(C.byref(bbs), len(u), _capi.as_ptr(u, C.c_double), _capi.as_ptr(v, C.c_double),
                                  du, dv, _capi.as_ptr(out, C.c_double))
        return rc, out

    def bending(self, bbs):
        NC = bbs.nptsu * bbs.nptsv
        out = np.zeros((NC, NC))
        rc = self._f("bbs_bending")(C.byref(bbs), _capi.as_ptr(out, C.c_double))
        return rc, out

    def surface_vertices(self, bbs, ctrl, xs, ys):
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        out = np.zeros((xs * ys, 3), dtype=np.float32)
        rc = self._f("surface_vertices")(C.byref(bbs), _capi.as_ptr(ctrl, C.c_double), xs, ys,
                                         _capi.as_ptr(out, C.c_float))
        return rc, out


def bbs_oracle():
    return BbsApi(load(), "oracle_")


class BbsReference:
    """The reference's own Thirdparty/BBS/bbs.cc, compiled where it lies (oracle/_ref/).
    C++ symbols are called through their mangled names; all arguments are plain pointers."""

    class _bbs_t(C.Structure):
        _fields_ = [("umin", C.c_double), ("umax", C.c_double), ("nptsu", C.c_int), ("vmin", C.c_double),
                    ("vmax", C.c_double), ("nptsv", C.c_int), ("valdim", C.c_int)]

    def __init__(self):
        path = os.path.join(_DIR, "_ref", "libbbs_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self._eval = self.lib._ZN3BBS4evalEPNS_6_bbs_tEPdS2_S2_iS2_ii
        self._eval.restype = None
        self._eval.argtypes = [C.POINTER(self._bbs_t), _capi.c_double_p, _capi.c_double_p, _capi.c_double_p,
                               C.c_int, _capi.c_double_p, C.c_int, C.c_int]
        self._coloc_deriv = self.lib._ZN3BBS11coloc_derivEPNS_6_bbs_tEPdS2_iiiS2_PmS3_
        self._coloc_deriv.restype = C.c_int
        self._coloc_deriv.argtypes = [C.POINTER(self._bbs_t), _capi.c_double_p, _capi.c_double_p, C.c_int, C.c_int,
                                      C.c_int, _capi.c_double_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        self._coloc = self.lib._ZN3BBS5colocEPNS_6_bbs_tEPdS2_iS2_PmS3_
        self._coloc.restype = C.c_int
        self._coloc.argtypes = [C.POINTER(self._bbs_t), _capi.c_double_p, _capi.c_double_p, C.c_int,
                                _capi.c_double_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        self._bending = self.lib._ZN3BBS10bending_urEPNS_6_bbs_tEdPdPmS3_
        self._bending.restype = None
        self._bending.argtypes = [C.POINTER(self._bbs_t), C.c_double, _capi.c_double_p, C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_size_t)]

    def _t(self, b):
        t = self._bbs_t()
        t.umin, t.umax, t.nptsu, t.vmin, t.vmax, t.nptsv, t.valdim = b.umin, b.umax, b.nptsu, b.vmin, b.vmax, \
            b.nptsv, b.valdim
        return t

    def eval(self, bbs, ctrl, u, v, du=0, dv=0):
        u = np.ascontiguousarray(u, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
        out = np.zeros((len(u), bbs.valdim))
        t = self._t(bbs)
        self._eval(C.byref(t), _capi.as_ptr(ctrl, C.c_double), _capi.as_ptr(u, C.c_double),
                   _capi.as_ptr(v, C.c_double), len(u), _capi.as_ptr(out, C.c_double), du, dv)
        return 0, out

    def coloc(self, bbs, u, v, du=0, dv=0):
        """sparse (pr, ir, jc) of the reference expanded to dense rows"""
        u = np.ascontiguousarray(u, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        n, NC = len(u), bbs.nptsu * bbs.nptsv
        pr = np.zeros(16 * n)
        ir = np.zeros(16 * n, dtype=np.uint64)
        jc = np.zeros(NC + 1, dtype=np.uint64)
        t = self._t(bbs)
        if du == 0 and dv == 0:
            rc = self._coloc(C.byref(t), _capi.as_ptr(u, C.c_double), _capi.as_ptr(v, C.c_double), n,
                             _capi.as_ptr(pr, C.c_double), ir.ctypes.data_as(C.POINTER(C.c_size_t)),
                             jc.ctypes.data_as(C.POINTER(C.c_size_t)))
        else:
            rc = self._coloc_deriv(C.byref(t), _capi.as_ptr(u, C.c_double), _capi.as_ptr(v, C.c_double), n, du, dv,
                                   _capi.as_ptr(pr, C.c_double), ir.ctypes.data_as(C.POINTER(C.c_size_t)),
                                   jc.ctypes.data_as(C.POINTER(C.c_size_t)))
        out = np.zeros((n, NC))
        if rc == 0:
            for col in range(NC):
                for k in range(int(jc[col]), int(jc[col + 1])):
                    out[int(ir[k]), col] = pr[k]
        return rc, out

    def bending(self, bbs, lam=1.0):
        """upper-right sparse matrix of bending_ur, symmetrised to dense"""
        NC = bbs.nptsu * bbs.nptsv
        cap = NC * 49
        pr = np.zeros(cap)
        ir = np.zeros(cap, dtype=np.uint64)
        jc = np.zeros(NC + 1, dtype=np.uint64)
        t = self._t(bbs)
        self._bending(C.byref(t), lam, _capi.as_ptr(pr, C.c_double), ir.ctypes.data_as(C.POINTER(C.c_size_t)),
                      jc.ctypes.data_as(C.POINTER(C.c_size_t)))
        out = np.zeros((NC, NC))
        for col in range(NC):
            for k in range(int(jc[col]), int(jc[col + 1])):
                out[int(ir[k]), col] = pr[k]
                out[col, int(ir[k])] = pr[k]
        return 0, out
