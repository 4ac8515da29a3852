"""Builds tests/_emu/libdefslam_emu.so: the kernel sources compiled with g++ as a
one-thread team.  TEST INFRASTRUCTURE ONLY (see emu_sft.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(ROOT, "tests", "_emu", "libdefslam_emu.so")


def build() -> str:
    srcs = [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith(".cpp")]
    csrc = os.path.join(ROOT, "defslam_b200", "csrc")
    deps = srcs + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
    deps.append(os.path.join(ROOT, "include", "defslam_b200.h"))
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-g", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-Wall", "-Wno-unused-function",
           "-o", OUT] + srcs
    subprocess.run(cmd, check=True, capture_output=True)
    return OUT
