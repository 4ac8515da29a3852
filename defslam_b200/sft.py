"""Thin Python harness over the C ABI of the SfT solve (tests and bench.py use it).

Everything here is marshalling: numpy arrays -> defslam_sft_problem structs -> the CUDA
library.  There is no computation and no fallback on this side."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi


class DefslamError(RuntimeError):
    def __init__(self, fn: str, rc: int):
        names = {-1: "EBADARG", -2: "ECUDA", -3: "ENUMERIC", -4: "ETOOLARGE", -5: "ENOTIMPL"}
        super().__init__(f"{fn} failed: {rc} ({names.get(rc, '?')})")
        self.rc = rc


def _check(fn: str, rc: int):
    if rc != 0:
        raise DefslamError(fn, rc)


class SftOutput:
    """numpy-side holder of one defslam_sft_result."""

    def __init__(self, n_nodes: int, n_matches: int, trace_capacity: int = 64):
        self.nodes = np.zeros((n_nodes, 3))
        self.outlier = np.zeros(max(n_matches, 1), dtype=np.uint8)
        self.role = np.zeros(n_nodes, dtype=np.uint8)
        self.trace = np.zeros((trace_capacity, 4))
        self.n_matches = n_matches
        self.r = None

    def fill(self, r: _capi.SftResult):
        r.node_xyz_out = _capi.as_ptr(self.nodes, C.c_double)
        r.outlier_out = _capi.as_ptr(self.outlier, C.c_uint8)
        r.node_role_out = _capi.as_ptr(self.role, C.c_uint8)
        r.trace = _capi.as_ptr(self.trace, C.c_double)
        r.trace_capacity = self.trace.shape[0]
        self.r = r

    @property
    def T_cw(self):
        return np.array(list(self.r.T_cw_out), dtype=np.float32).reshape(4, 4)


class Template:
    """Device-resident plan of a mesh template (defslam_template_create/destroy)."""

    def __init__(self, mesh, device: int = -1):
        self.mesh = mesh
        self.lib = _capi.load()
        h = C.c_void_p()
        _check("defslam_template_create", self.lib.defslam_template_create(C.byref(mesh.desc()), device, C.byref(h)))
        self.handle = h

    def info(self):
        v = [C.c_int32() for _ in range(5)]
        _check("defslam_template_info", self.lib.defslam_template_info(self.handle, *[C.byref(x) for x in v]))
        return dict(zip(["bandwidth", "band_ld", "dn_pad", "n_blocks", "smem_bytes"], [x.value for x in v]))

    def close(self):
        if self.handle:
            self.lib.defslam_template_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _problems(frames, template: Template | None):
    arr = (_capi.SftProblem * len(frames))()
    for i, f in enumerate(frames):
        arr[i] = f.problem(template.handle if template is not None else None)
    return arr


def _results(frames):
    outs = [SftOutput(f.template.n_nodes, f.n_matches) for f in frames]
    arr = (_capi.SftResult * len(frames))()
    for i, o in enumerate(outs):
        o.fill(arr[i])
    return outs, arr


class HostBatch:
    """Descriptor arrays (defslam_sft_problem / defslam_sft_result) over HOST buffers, prepared
    once; solve() is exactly one defslam_sft_solve_batched call: marshalling into pinned memory,
    H2D, the kernel, D2H and the scatter into the result arrays all happen inside it."""

    def __init__(self, frames, template: Template | None = None, device: int = -1):
        self.lib = _capi.load()
        self.frames = frames
        self.device = device
        self.probs = _problems(frames, template)
        self.outs, self.res = _results(frames)

    def solve(self):
        _check("defslam_sft_solve_batched",
               self.lib.defslam_sft_solve_batched(len(self.frames), self.probs, self.res, self.device))
        return self.outs


def solve_batched(frames, template: Template | None = None, device: int = -1):
    """defslam_sft_solve_batched on host buffers (H2D + kernel + D2H inside the call)."""
    lib = _capi.load()
    probs = _problems(frames, template)
    outs, res = _results(frames)
    _check("defslam_sft_solve_batched", lib.defslam_sft_solve_batched(len(frames), probs, res, device))
    return outs


def solve(frame, template: Template | None = None):
    lib = _capi.load()
    probs = _problems([frame], template)
    outs, res = _results([frame])
    _check("defslam_sft_solve", lib.defslam_sft_solve(C.byref(probs[0]), C.byref(res[0])))
    return outs[0]


def normal_equations(frame, template: Template | None = None):
    lib = _capi.load()
    D = 3 * frame.template.n_nodes + 6
    H = np.zeros((D, D))
    b = np.zeros(D)
    chi = C.c_double(0)
    p = frame.problem(template.handle if template is not None else None)
    _check("defslam_sft_normal_equations",
           lib.defslam_sft_normal_equations(C.byref(p), _capi.as_ptr(H, C.c_double), _capi.as_ptr(b, C.c_double),
                                            C.cast(C.byref(chi), _capi.c_double_p)))
    return H, b, chi.value


class ResidentBatch:
    """Frames marshalled and uploaded once; run() is kernel-only."""

    def __init__(self, frames, template: Template | None = None, device: int = -1):
        self.lib = _capi.load()
        self.frames = frames
        self._probs = _problems(frames, template)
        h = C.c_void_p()
        _check("defslam_sft_batch_create",
               self.lib.defslam_sft_batch_create(len(frames), self._probs, device, C.byref(h)))
        self.handle = h

    def run(self) -> float:
        _check("defslam_sft_batch_run", self.lib.defslam_sft_batch_run(self.handle))
        return self.lib.defslam_last_kernel_ms()

    def fetch(self):
        outs, res = _results(self.frames)
        _check("defslam_sft_batch_fetch", self.lib.defslam_sft_batch_fetch(self.handle, res))
        return outs

    def info(self):
        g, t, s = C.c_int32(), C.c_int32(), C.c_int32()
        h2d, d2h = C.c_int64(), C.c_int64()
        ms = C.c_double()
        _check("defslam_sft_batch_info",
               self.lib.defslam_sft_batch_info(self.handle, C.byref(g), C.byref(t), C.byref(s), C.byref(h2d),
                                               C.byref(d2h), C.cast(C.byref(ms), _capi.c_double_p)))
        return dict(grid=g.value, threads=t.value, smem_bytes=s.value, h2d_bytes=h2d.value, d2h_bytes=d2h.value,
                    last_kernel_ms=ms.value)

    def close(self):
        if self.handle:
            self.lib.defslam_sft_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
