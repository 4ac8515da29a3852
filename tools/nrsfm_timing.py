"""Quick device timing of the three NRSfM stages (diagnostics; bench.py is the judged number)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from defslam_b200 import nrsfm, _capi
from oracle import oracle_py

lib = _capi.load()
api = nrsfm.Api(lib, "defslam_")
orc = nrsfm.Api(oracle_py.load(), "oracle_")
nwin = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wins = [nrsfm.make_window(100 + i, n_keypoints=1200, n_views=4) for i in range(nwin)]
cases = [c for w in wins for c in nrsfm.schwarp_cases(w)]
rep = max(1, 592 // len(cases))
batch = cases * rep
for _ in range(2):
    t = time.time(); outs = api.schwarp_fit_batched(batch); wall = time.time() - t
print("schwarp: %d pairs  kernel %.3f ms  wall %.3f ms  -> %.0f fits/s resident, %.0f e2e" % (
    len(batch), lib.defslam_last_kernel_ms(), wall * 1e3, len(batch) / lib.defslam_last_kernel_ms() * 1e3, len(batch) / wall), flush=True)
t = time.time(); fo = [orc.schwarp_fit(c) for c in cases[:8]]; print("  oracle %.1f ms/fit" % ((time.time() - t) / 8 * 1e3))
fits = outs[:len(cases)]
ncs = [nrsfm.normals_case(w, fits[4 * i:4 * i + 4]) for i, w in enumerate(wins)]
# one big normals problem: concatenate windows
def cat(ncs, reps):
    import copy
    ptr = [0]; 
    for _ in range(reps):
        for nc in ncs:
            ptr.extend((nc.pair_ptr[1:] + ptr[-1]).tolist())
    f = lambda name: np.ascontiguousarray(np.concatenate([getattr(nc, name)[:nc.npairs] if name not in ("k_init", "ref_uv") else getattr(nc, name) for nc in ncs] * reps))
    return nrsfm.NormalsCase(pair_ptr=np.array(ptr, np.int32), J12=f("J12"), J21=f("J21"), H12=f("H12"), I1=f("I1"), I2=f("I2"),
                             pair_from_ref=f("pair_from_ref"), k_first=f("k_first"), k_init=f("k_init"), ref_uv=f("ref_uv"))
big = cat(ncs, max(1, 400 // nwin))
ncall = api.normals_prepare(big)  # output arrays allocated once, like a caller that reuses its buffers
for _ in range(3):
    t = time.time(); no = ncall(); wall = time.time() - t
print("normals: %d points %d pairs  kernel %.3f ms  wall %.3f ms -> %.2f Mpoints/s resident, %.2f e2e; iters mean %.1f max %d" % (
    big.n, big.npairs, lib.defslam_last_kernel_ms(), wall * 1e3, big.n / lib.defslam_last_kernel_ms() / 1e3, big.n / wall / 1e6, no.iters.mean(), no.iters.max()), flush=True)
t = time.time(); oo = orc.normals(ncs[0]); print("  oracle %.2f ms per %d points" % ((time.time() - t) * 1e3, ncs[0].n))
nouts = [api.normals(nc) for nc in ncs]
scs = [nrsfm.sfn_case(w, no) for w, no in zip(wins, nouts)]
sb = scs * max(1, 296 // len(scs))
for _ in range(2):
    t = time.time(); rcs = api.sfn_solve_batched(sb); wall = time.time() - t
print("sfn: %d keyframes  kernel %.3f ms  wall %.3f ms -> %.0f solves/s resident, %.0f e2e  rc ok %s" % (
    len(sb), lib.defslam_last_kernel_ms(), wall * 1e3, len(sb) / lib.defslam_last_kernel_ms() * 1e3, len(sb) / wall, (rcs == 0).all()), flush=True)
t = time.time(); orc.sfn_solve(scs[0]); print("  oracle %.1f ms/solve" % ((time.time() - t) * 1e3))
