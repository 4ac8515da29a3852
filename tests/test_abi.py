"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from defslam_b200 import _capi, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "defslam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(defslam_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cuda_lib):
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(cuda_lib, s), f"{s} declared in the header but not exported"


def test_ctypes_mirror_covers_the_header():
    assert set(header_symbols()) == set(_capi.PROTOTYPES.keys())


def test_library_is_cuda_code_for_sm_100a():
    # the product library must contain device code, not be a host-only shim
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_version_and_counters(cuda_lib):
    assert b"sm_100a" in cuda_lib.defslam_version()
    assert cuda_lib.defslam_kernel_launch_count() >= 0
    assert cuda_lib.defslam_device_count() >= 0


def test_no_cpu_fallback_without_device(cuda_lib):
    """On a box without a GPU every compute entry point fails with DEFSLAM_ECUDA."""
    if cuda_lib.defslam_device_count() > 0:
        pytest.skip("a CUDA device is present")
    tmpl, frames = synthetic.make_config_frames("C1", nframes=1)
    from defslam_b200 import sft
    with pytest.raises(sft.DefslamError) as e:
        sft.solve_batched(frames)
    assert e.value.rc == _capi.ECUDA
    h = C.c_void_p()
    assert cuda_lib.defslam_template_create(C.byref(tmpl.desc()), -1, C.byref(h)) == _capi.ECUDA
    b = _capi.Bbs(-1.0, 1.0, 13, -1.0, 1.0, 15, 1)
    ctrl = np.zeros(13 * 15)
    u = np.zeros(4)
    val = np.zeros(4)
    rc = cuda_lib.defslam_bbs_eval(C.byref(b), _capi.as_ptr(ctrl, C.c_double), 4, _capi.as_ptr(u, C.c_double),
                                   _capi.as_ptr(u, C.c_double), 0, 0, _capi.as_ptr(val, C.c_double))
    assert rc == _capi.ECUDA


def test_bad_arguments_are_rejected_before_touching_the_device(cuda_lib):
    assert cuda_lib.defslam_sft_solve(None, None) == _capi.EBADARG
    assert cuda_lib.defslam_template_create(None, -1, None) == _capi.EBADARG
    assert cuda_lib.defslam_bbs_bending(None, None) == _capi.EBADARG


def test_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under defslam_b200/ may reference it."""
    pkg = os.path.join(ROOT, "defslam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "oracle/" not in txt, f


def test_host_thread_pool():
    """the persistent marshalling pool of csrc/ds_host.h (pure C++, compiled for the host)"""
    import subprocess
    exe = os.path.join(ROOT, "tests", "_emu", "test_host_pool")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_host_pool.cc")],
                   check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "host pool: ok" in r.stdout, r.stdout + r.stderr
