"""Shared test helpers: loaders for the oracle, the emulated kernel sources and the golden files."""
import ctypes as C
import os

import numpy as np

from defslam_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


_emu = None


def emu_lib():
    """tests/_emu/libdefslam_emu.so: the kernel sources compiled with g++ (test infrastructure)."""
    global _emu
    if _emu is None:
        from tests.emu import build
        lib = C.CDLL(build.build())
        P = _capi
        lib.emu_sft_solve_batched.restype = C.c_int
        lib.emu_sft_solve_batched.argtypes = [C.c_int32, C.POINTER(P.SftProblem), C.POINTER(P.SftResult), C.c_int]
        lib.emu_sft_solve.restype = C.c_int
        lib.emu_sft_solve.argtypes = [C.POINTER(P.SftProblem), C.POINTER(P.SftResult)]
        lib.emu_sft_normal_equations.restype = C.c_int
        lib.emu_sft_normal_equations.argtypes = [C.POINTER(P.SftProblem), P.c_double_p, P.c_double_p, P.c_double_p]
        lib.emu_template_create.restype = C.c_int
        lib.emu_template_create.argtypes = [C.POINTER(P.TemplateDesc), C.c_int, C.POINTER(C.c_void_p)]
        lib.emu_template_destroy.restype = None
        lib.emu_template_destroy.argtypes = [C.c_void_p]
        lib.emu_plan_info.restype = C.c_int
        lib.emu_plan_info.argtypes = [C.c_void_p, P.c_int32_p]
        for nm in ("emu_mesh_laplacian", "emu_embed_points", "emu_mappoints_recalculate", "emu_new_map_points",
                   "emu_search_by_projection", "emu_search_by_schwarp"):
            f = getattr(lib, nm)
            f.restype = C.c_int
            f.argtypes = P.PROTOTYPES["defslam_" + nm[4:]][1]
        _emu = lib
    return _emu


def emu_solve_batched(frames):
    from defslam_b200 import sft
    lib = emu_lib()
    probs = (_capi.SftProblem * len(frames))()
    for i, f in enumerate(frames):
        probs[i] = f.problem()
    outs = [sft.SftOutput(f.template.n_nodes, f.n_matches) for f in frames]
    res = (_capi.SftResult * len(frames))()
    for i, o in enumerate(outs):
        o.fill(res[i])
    rc = lib.emu_sft_solve_batched(len(frames), probs, res, -1)
    return rc, outs


def emu_normal_equations(frame):
    lib = emu_lib()
    D = 3 * frame.template.n_nodes + 6
    H = np.zeros((D, D))
    b = np.zeros(D)
    chi = C.c_double(0)
    p = frame.problem()
    rc = lib.emu_sft_normal_equations(C.byref(p), _capi.as_ptr(H, C.c_double), _capi.as_ptr(b, C.c_double),
                                      C.cast(C.byref(chi), _capi.c_double_p))
    return rc, H, b, chi.value


def rel_nodes(a, b):
    return np.abs(a - b).max() / np.sqrt((b ** 2).sum(1).mean())


def mesh_laplacian_call(fn, nodes, facets, max_ring=8):
    n, nf = nodes.shape[0], facets.shape[0]
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    facets = np.ascontiguousarray(facets, dtype=np.int32)
    cnt = np.zeros(n, np.int32)
    idx = np.zeros((n, max_ring), np.int32)
    w = np.zeros((n, max_ring))
    bd = np.zeros(n, np.uint8)
    k0 = np.zeros(n)
    ne = C.c_int32(0)
    ab = np.zeros((3 * nf, 2), np.int32)
    l0 = np.zeros(3 * nf)
    med = C.c_double(0)
    rc = fn(n, _capi.as_ptr(nodes, C.c_double), nf, _capi.as_ptr(facets, C.c_int32), max_ring,
            _capi.as_ptr(cnt, C.c_int32), _capi.as_ptr(idx, C.c_int32), _capi.as_ptr(w, C.c_double),
            _capi.as_ptr(bd, C.c_uint8), _capi.as_ptr(k0, C.c_double), C.cast(C.byref(ne), _capi.c_int32_p),
            _capi.as_ptr(ab, C.c_int32), _capi.as_ptr(l0, C.c_double), C.cast(C.byref(med), _capi.c_double_p))
    return rc, dict(cnt=cnt, idx=idx, w=w, boundary=bd, kappa0=k0, n_edges=ne.value, edge_ab=ab[:ne.value],
                    edge_len0=l0[:ne.value], median=med.value)


def embed_call(fn, nodes, facets, pts):
    n, nf, npt = nodes.shape[0], facets.shape[0], pts.shape[0]
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    facets = np.ascontiguousarray(facets, dtype=np.int32)
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    of = np.zeros(npt, np.int32)
    on = np.zeros((npt, 3), np.int32)
    ob = np.zeros((npt, 3), np.float32)
    rc = fn(n, _capi.as_ptr(nodes, C.c_double), nf, _capi.as_ptr(facets, C.c_int32), npt, _capi.as_ptr(pts, C.c_float),
            _capi.as_ptr(of, C.c_int32), _capi.as_ptr(on, C.c_int32), _capi.as_ptr(ob, C.c_float))
    return rc, of, on, ob
