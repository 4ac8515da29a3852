import sys, time, os
sys.path.insert(0, "/root/repo")
os.chdir("/root/repo")
import bench
from defslam_b200 import _capi
lib = _capi.load()
wl = bench.nrsfm_workload()
api = wl["api"]
for name, call, n in (("schwarp", api.schwarp_prepare(wl["pairs"]), len(wl["pairs"])), ("sfn", api.sfn_prepare(wl["keyframes"]), len(wl["keyframes"]))):
    for _ in range(3): call()
    t = time.perf_counter(); k = 0.0
    for _ in range(10):
        call(); k += lib.defslam_last_kernel_ms()
    dt = (time.perf_counter() - t) / 10
    print("%s: %d units, wall %.3f ms (%.0f /s end to end), kernels %.3f ms (%.0f /s)" % (name, n, dt * 1e3, n / dt, k / 10, n / (k / 10) * 1e3))
