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
    lib.oracle_regular_triangulation.restype = C.c_int
    lib.oracle_regular_triangulation.argtypes = [C.c_int, C.c_int, P.c_int32_p]
    lib.oracle_mesh_laplacian.restype = C.c_int
    lib.oracle_mesh_laplacian.argtypes = _capi.PROTOTYPES["defslam_mesh_laplacian"][1]
    lib.oracle_embed_points.restype = C.c_int
    lib.oracle_embed_points.argtypes = _capi.PROTOTYPES["defslam_embed_points"][1]
    if not native:
        _lib = lib
    return lib


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


def sft_solve(frame, lib=None):
    lib = lib or load()
    out = SftOutput(frame.template.n_nodes, frame.n_matches)
    p = frame.problem()
    rc = lib.oracle_sft_solve(C.byref(p), C.byref(out.r))
    if rc != 0:
        raise RuntimeError(f"oracle_sft_solve rc={rc}")
    return out


def sft_normal_equations(frame, lib=None):
    lib = lib or load()
    D = 3 * frame.template.n_nodes + 6
    H = np.zeros((D, D))
    b = np.zeros(D)
    chi = C.c_double(0)
    p = frame.problem()
    rc = lib.oracle_sft_normal_equations(C.byref(p), _capi.as_ptr(H, C.c_double), _capi.as_ptr(b, C.c_double),
                                         C.cast(C.byref(chi), _capi.c_double_p))
    if rc != 0:
        raise RuntimeError(f"oracle_sft_normal_equations rc={rc}")
    return H, b, chi.value
