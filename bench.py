#!/usr/bin/env python
"""bench.py -- SfT solves/sec on the BASELINE.json headline shape (640x480, 1000 matches, 13x13 grid).

A "step" is one pass of the hot path over one batch of synthetic frames (config C2 of
SURVEY.md 8(d): G=13, M=1000, 10 LM iterations per frame, independent frames).

  value      solves/s of the whole job with the batch already resident in HBM: one launch of
             the persistent LM kernel per step, timed with CUDA events on the launching stream.
  e2e        the same metric through defslam_sft_solve_batched() on HOST buffers (marshalling,
             H2D, kernel, D2H inside the timed region).
  roofline   algorithmic bytes of the LM kernel (banded-H variant) / its measured
             duration, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the CPU oracle (a restatement of the reference's g2o path: dense LDLT per LM
             trial) on a bounded sample of the same frames, one frame per host thread.

`--impl reference` times that CPU path instead (the reference itself cannot be built in this
image: Eigen/OpenCV/Ceres are absent -- see DESIGN.md).

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torch.distributed.run
(one rank per GPU, NCCL only for the barrier and the max-over-ranks of the timings).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CFG = "C2"
N_DISTINCT = 64          # distinct synthetic frames, tiled to the batch size
FRAMES_PER_GPU = 2368    # 148 SMs x 2 resident CTAs x 8 waves
CPU_SAMPLE_FRAMES = 64   # bounded CPU sample: ~64 x 0.25 s of CPU work


METRIC = "SfT solves/sec (640x480, 1000 matches, 13x13 grid, 10 LM iterations)"


def workload_string(cfg: str) -> str:
    """the SAME string in both arms (the batch size lives in config.frames_per_step_per_gpu)"""
    from defslam_b200 import synthetic
    c = synthetic.CONFIGS[cfg]
    return (f"{cfg}: G={c['G']} mesh (D={6 + 3 * c['G'] ** 2}), M={c['M']} matches, {c['max_iterations']} LM "
            f"iterations, independent frames")


def algorithmic_bytes_per_iteration(G: int, M: int) -> float:
    """B_iter = B_in + B_H(banded) + B_b + B_x per LM iteration per frame."""
    Nn = G * G
    Nint = (G - 2) ** 2
    E = 3 * G * G - 4 * G + 1
    D = 6 + 3 * Nn
    b_in = 64 * M + 48 * Nn + 128 * Nint + 16 * E
    b_h = 8 * (6 * D + 9 * 19 * Nn / 2)
    return b_in + b_h + 8 * D + 8 * D


def fp64_roofline(G: int, trials: float, iters: float, launch_s: float, clocks: dict):
    """Compute view of the LM kernel: algorithmic FP64 flops (banded Cholesky D*bw^2 + two band sweeps 4*D*bw per
    trial, assembly ~1 MFLOP per iteration -- SURVEY.md 8(d)) / launch time, against the FP64 peak measured with
    tools/microbench.cu on this pool's B200 (DFMA and DMMA both 64 FMA/clk/SM) at the SM clock of the run."""
    Dn = 3 * G * G
    bw = 3 * 2 * G + 2
    flops = trials * (Dn * bw * bw + 4.0 * Dn * bw) + iters * 0.96e6 * (G * G) / 169.0
    mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
    peak = 64 * 2 * 148 * mhz * 1e6 / 1e12
    ach = flops / launch_s / 1e12
    return {"achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": "measured: 64 FP64 FMA/clk/SM (tools/microbench.cu) x 148 SMs x SM clock"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_frames(n_total: int):
    from defslam_b200 import synthetic
    tmpl, base = synthetic.make_config_frames(CFG, nframes=N_DISTINCT)
    frames = [base[i % N_DISTINCT] for i in range(n_total)]
    return tmpl, base, frames


def cpu_solve_rate(frames, n_threads: int):
    """Oracle (dense-LDLT restatement of the reference path), one frame per thread."""
    from oracle import oracle_py
    try:
        lib = oracle_py.load(native=True)
        build = "-O3 -march=native"
    except Exception:
        lib = oracle_py.load()
        build = "-O3 -march=x86-64-v3"
    work = list(frames)
    lock = threading.Lock()
    pos = [0]

    def worker():
        while True:
            with lock:
                i = pos[0]
                pos[0] += 1
            if i >= len(work):
                return
            oracle_py.sft_solve(work[i], lib=lib)  # ctypes releases the GIL

    oracle_py.sft_solve(work[0], lib=lib)  # warm-up
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker) for _ in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return len(work) / dt, dt, build


# ------------------------------------------------------------------ NRSfM stages -----------
NRSFM_WINDOWS = 8        # distinct synthetic keyframe windows (1200 keypoints, 4 later views each)
NRSFM_PAIRS = 592        # Schwarp fits per step  (148 SMs x 4 waves)
NRSFM_POINTS = 480000    # map points per normals step
NRSFM_KEYFRAMES = 296    # shape-from-normals solves per step


def nrsfm_workload():
    """One mapping workload per stage, built from NRSFM_WINDOWS keyframe windows (inputs of the later
    stages come from the oracle-independent product path itself: fits -> normals -> SfN)."""
    from defslam_b200 import nrsfm
    api = nrsfm.Api()
    wins = [nrsfm.make_window(100 + i, n_keypoints=1200, n_views=4) for i in range(NRSFM_WINDOWS)]
    cases = [c for w in wins for c in nrsfm.schwarp_cases(w)]
    fits = api.schwarp_fit_batched(cases)
    ncs = [nrsfm.normals_case(w, fits[4 * i:4 * i + 4]) for i, w in enumerate(wins)]
    reps = max(1, NRSFM_POINTS // sum(nc.n for nc in ncs))
    ptr = [0]
    for _ in range(reps):
        for nc in ncs:
            ptr.extend((nc.pair_ptr[1:] + ptr[-1]).tolist())

    def cat(name, per_point=False):
        return np.ascontiguousarray(np.concatenate(
            [getattr(nc, name) if per_point else getattr(nc, name)[:nc.npairs] for nc in ncs] * reps))
    big = nrsfm.NormalsCase(pair_ptr=np.array(ptr, np.int32), J12=cat("J12"), J21=cat("J21"), H12=cat("H12"),
                            I1=cat("I1"), I2=cat("I2"), pair_from_ref=cat("pair_from_ref"), k_first=cat("k_first"),
                            k_init=cat("k_init", True), ref_uv=cat("ref_uv", True))
    scs = [nrsfm.sfn_case(w, api.normals(nc)) for w, nc in zip(wins, ncs)]
    return dict(api=api, wins=wins, pairs=cases * max(1, NRSFM_PAIRS // len(cases)), pair_base=cases, normals=big,
                normals_base=ncs, keyframes=scs * max(1, NRSFM_KEYFRAMES // len(scs)), keyframes_base=scs)


def nrsfm_gpu(lib, wl, steps, warmup, flush_fn):
    """[units per step, device ms, wall s] per stage; device time = CUDA events around the stage's
    kernel on the library's stream, wall = the whole C-ABI call on host buffers."""
    api = wl["api"]
    # descriptors over the host arrays are built once; each call below is ONE C-ABI call
    stages = {"schwarp_fit": (len(wl["pairs"]), api.schwarp_prepare(wl["pairs"])),
              "normals": (wl["normals"].n, api.normals_prepare(wl["normals"])),
              "sfn": (len(wl["keyframes"]), api.sfn_prepare(wl["keyframes"]))}
    res = {}
    for name, (units, call) in stages.items():
        for _ in range(min(warmup, 2)):
            call()
        dev_ms, wall_s = 0.0, 0.0
        for _ in range(steps):
            flush_fn()
            t0 = time.perf_counter()
            call()
            wall_s += time.perf_counter() - t0
            dev_ms += lib.defslam_last_kernel_ms()
        if name in ("normals", "schwarp_fit"):
            # the end-to-end call pipelines a large batch in chunks (pack / upload / kernel / download / unpack of
            # neighbouring chunks overlap); the kernel-only number is the same batch in one launch
            os.environ["DEFSLAM_NO_PIPELINE"] = "1"
            dev_ms = 0.0
            call()
            for _ in range(steps):
                flush_fn()
                call()
                dev_ms += lib.defslam_last_kernel_ms()
            del os.environ["DEFSLAM_NO_PIPELINE"]
        res[name] = [units, dev_ms, wall_s]
    return res


def _threaded(fn, items, n_threads):
    lock, pos = threading.Lock(), [0]

    def worker():
        while True:
            with lock:
                i = pos[0]
                pos[0] += 1
            if i >= len(items):
                return
            fn(items[i])
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker) for _ in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return time.perf_counter() - t0


def nrsfm_cpu(wl, cores):
    """The oracle (dense Jacobians / dense Cholesky / Householder QR like the reference's Ceres+Eigen
    path) on a bounded sample of the same units, one unit per host thread."""
    from defslam_b200 import nrsfm
    from oracle import oracle_py
    try:
        lib = oracle_py.load(native=True)
    except Exception:
        lib = oracle_py.load()
    orc = nrsfm.Api(lib, "oracle_")
    out = {}
    pairs = (wl["pair_base"] * 8)[:max(cores, 32)]
    out["schwarp_fit"] = len(pairs) / _threaded(orc.schwarp_fit, pairs, cores)
    ncs = (wl["normals_base"] * 64)[:max(cores, 64) * 4]
    out["normals"] = sum(nc.n for nc in ncs) / _threaded(orc.normals, ncs, cores)
    import copy
    kfs = [copy.copy(k) for k in (wl["keyframes_base"] * 8)[:max(cores, 16)]]
    for k in kfs:
        k.ctrl = None
    out["sfn"] = len(kfs) / _threaded(orc.sfn_solve, kfs, cores)
    return out


def stream_bench(lib, with_cpu: bool):
    """Config C3 of SURVEY.md 8(d): 300-frame stream, 17x17 mesh, ~600 matches, keyframe every 10 frames,
    NRSfM (17x17 control grid) every 5th keyframe -- frames/s of the whole loop through the C ABI on
    host buffers, one frame at a time (this is the latency the tracking thread sees)."""
    from defslam_b200 import stream
    cfg = stream.StreamConfig()
    be = stream.cuda_backend()
    stream.run_stream(be, stream.StreamConfig(n_frames=12), keep_nodes=False)      # warm-up (context, allocations)
    l0 = lib.defslam_kernel_launch_count()
    t0 = time.perf_counter()
    r = stream.run_stream(be, cfg, keep_nodes=False)
    dt = time.perf_counter() - t0
    out = {"workload": f"C3: {cfg.n_frames}-frame stream, G={cfg.G} mesh, {cfg.n_points} map points, keyframe every "
                       f"{cfg.kf_every} frames, NRSfM ({cfg.nptsu}x{cfg.nptsv} control grid, {cfg.n_views} views) every "
                       f"{cfg.nrsfm_every_kf}th keyframe; wall clock of the loop incl. synthetic observations",
           "value": cfg.n_frames / dt, "unit": "frames/s", "ms_per_frame": 1e3 * dt / cfg.n_frames,
           "nrsfm_events": r.n_nrsfm, "template_updates": r.n_template_updates,
           "gpu_launches": int(lib.defslam_kernel_launch_count() - l0),
           "node_rmse_vs_ground_truth": {"median": float(np.median(r.rmse)), "max": float(np.max(r.rmse))},
           "lm_trials_per_frame": float(np.mean(r.trials)),
           # where the wall clock goes: the SfT call of every frame, the NRSfM events, and the rest of the host loop
           # (synthetic observations, NumPy marshalling) -- the like-for-like ratio against the CPU side is tracking_ms
           "stages": {"tracking_ms_per_frame": 1e3 * r.t_sft / cfg.n_frames,
                      "nrsfm_ms_per_event": 1e3 * r.t_nrsfm / max(r.n_nrsfm, 1),
                      "host_loop_ms_per_frame": 1e3 * (dt - r.t_sft - r.t_nrsfm) / cfg.n_frames}}
    if with_cpu:
        from oracle import oracle_py
        try:
            olib = oracle_py.load(native=True)
        except Exception:
            olib = oracle_py.load()
        ob = stream.Backend(olib, "oracle_", lambda f: oracle_py.sft_solve(f, olib))
        n = 12
        t0 = time.perf_counter()
        ro = stream.run_stream(ob, stream.StreamConfig(n_frames=n), keep_nodes=True)
        dto = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n / dto, "unit": "frames/s", "cores": 1, "kind": "port",
                               "tracking_ms_per_frame": 1e3 * ro.t_sft / n,
                               "sample": f"first {n} frames of the same stream (tracking only: the first NRSfM event "
                                         f"is at frame 50), oracle on one core, {dto:.1f} s"}
        rg = stream.run_stream(be, stream.StreamConfig(n_frames=n), keep_nodes=True)
        out["node_rel_err_vs_oracle"] = float(max(np.abs(a - b).max() / np.sqrt((b ** 2).sum(1).mean())
                                                  for a, b in zip(rg.nodes_cam, ro.nodes_cam)))
    return out


def matching_bench(lib, with_cpu: bool):
    """Match production that feeds the SfT solve: one call per frame through the C ABI on
    host buffers (this is a latency path: 1200 map points x ~1400 keypoints), the oracle's literal loop beside it."""
    from defslam_b200 import matching
    out = {}
    reps = 40
    for name, case, fn in (("search_by_projection", matching.make_case(1), matching.search_by_projection),
                           ("search_by_schwarp", matching.make_warp_case(1), matching.search_by_schwarp)):
        for _ in range(3):
            fn(case)
        l0 = lib.defslam_kernel_launch_count()
        t0 = time.perf_counter()
        for _ in range(reps):
            m, n = fn(case)
        dt = (time.perf_counter() - t0) / reps
        out[name] = {"value": 1.0 / dt, "unit": "frames/s", "ms_per_call": 1e3 * dt, "matches": int(n),
                     "gpu_launches": int(lib.defslam_kernel_launch_count() - l0)}
        if with_cpu:
            from oracle import oracle_py
            olib = oracle_py.load()
            mo, no = fn(case, olib, "oracle_")
            t0 = time.perf_counter()
            for _ in range(reps):
                fn(case, olib, "oracle_")
            dto = (time.perf_counter() - t0) / reps
            out[name]["cpu_baseline"] = {"value": 1.0 / dto, "unit": "frames/s", "ms_per_call": 1e3 * dto, "cores": 1,
                                         "kind": "port"}
            out[name]["identical_to_oracle"] = bool(no == n and np.array_equal(mo, m))
    out["workload"] = "1200 map points / keypoints of the previous (key)frame against ~1400-1800 keypoints, 256-bit ORB descriptors"
    return out


def _median_ms(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(t))


def _oracle_native():
    from oracle import oracle_py
    try:
        return oracle_py.load(native=True)
    except Exception:
        return oracle_py.load()


def configs_bench(lib, args, rank, world, dist, flush_fn, with_cpu: bool):
    """Every BASELINE.json config as the survey shapes it; C2-batched is the headline above.
      C1  G=9,  M=300,  B=1   the reference's own CPU-runnable case: oracle ms on ONE core (GPU latency beside it)
      C2  G=13, M=1000, B=1   the call DefTracking makes every frame: defslam_sft_solve on host buffers vs one core
      C4  G=10, M=400,  B=256 sharded 256/N per rank (shard.shard_range), STRONG scaling: batch latency + solves/s
      C5  G=25, M=2000, B=64  stress mesh with its own roofline (fp64 and HBM views)
    Called by every rank (C4 reduces over ranks); C1/C2/C5 run on rank 0."""
    from defslam_b200 import sft, shard, synthetic
    from oracle import oracle_py
    out = {}
    olib = _oracle_native() if (with_cpu and rank == 0) else None
    import torch

    # ---- C1 and C2 at B = 1 (latency)
    if rank == 0:
        for name, reps_cpu in (("C1", 20), ("C2", 5)):
            c = synthetic.CONFIGS[name]
            tmpl, frames = synthetic.make_config_frames(name, nframes=4)
            T = sft.Template(tmpl)
            f = frames[0]
            gpu_ms = _median_ms(lambda: sft.solve(f, template=T), 30, warm=3)
            kern_ms = float(lib.defslam_last_kernel_ms())
            o = sft.solve(f, template=T)
            ent = {"workload": workload_string(name), "frames": 1,
                   "gpu_ms": gpu_ms, "gpu_kernel_ms": kern_ms,
                   "how": "median of 30 defslam_sft_solve calls on host buffers (marshal + H2D + kernel + D2H inside)",
                   "lm_iterations": int(o.r.lm_iterations), "lm_trials": int(o.r.lm_trials)}
            if olib is not None:
                cpu_ms = _median_ms(lambda: oracle_py.sft_solve(f, lib=olib), reps_cpu, warm=1)
                ref = oracle_py.sft_solve(f, lib=olib)
                ent["cpu_ms"] = cpu_ms
                ent["cpu_baseline"] = {"value": 1e3 / cpu_ms, "unit": "solves/s", "cores": 1, "kind": "port",
                                       "sample": f"median of {reps_cpu} oracle solves of the same frame on one core"}
                ent["speedup_vs_1_core"] = cpu_ms / gpu_ms
                ent["node_rel_err_vs_oracle"] = float(np.abs(o.nodes - ref.nodes).max() / np.sqrt((ref.nodes ** 2).sum(1).mean()))
            out[name + "_b1"] = ent
            T.close()

    # ---- C4: 256 frames sharded over the ranks, strong scaling
    c4 = synthetic.CONFIGS["C4"]
    tmpl4, frames4 = synthetic.make_config_frames("C4", nframes=c4["B"])
    lo, hi = shard.shard_range(c4["B"], rank, world)
    mine = frames4[lo:hi]
    T4 = sft.Template(tmpl4)
    k_ms, e_ms = 0.0, 0.0
    reps = 5
    if mine:
        rb4 = sft.ResidentBatch(mine, template=T4)
        hb4 = sft.HostBatch(mine, template=T4)
        for _ in range(2):
            rb4.run(); hb4.solve()
        ks, es = [], []
        for _ in range(reps):
            flush_fn()
            ks.append(rb4.run())
            flush_fn()
            t0 = time.perf_counter()
            hb4.solve()
            es.append(1e3 * (time.perf_counter() - t0))
        k_ms, e_ms = float(np.median(ks)), float(np.median(es))
        rb4.close()
    dev = torch.device("cuda", torch.cuda.current_device())
    d = dist if (world > 1) else None
    k_job, solved = shard.reduce_job_stats(k_ms, len(mine), d, dev)
    e_job, _ = shard.reduce_job_stats(e_ms, len(mine), d, dev)
    if rank == 0:
        ent = {"workload": workload_string("C4"), "frames": c4["B"], "frames_per_gpu": -(-c4["B"] // world),
               "n_gpus": world, "scaling": "strong",
               "batch_latency_ms": k_job, "value": solved / (k_job * 1e-3), "unit": "solves/s",
               "e2e": {"batch_latency_ms": e_job, "value": solved / (e_job * 1e-3), "unit": "solves/s"},
               "how": f"median of {reps} runs, max over ranks; resident launch (CUDA events) and one "
                      "defslam_sft_solve_batched call on host buffers per rank"}
        if olib is not None:
            cores = os.cpu_count() or 1
            rate, dt, build = cpu_solve_rate(frames4, cores)
            ent["cpu_baseline"] = {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port",
                                   "sample": f"all {c4['B']} frames, one frame per thread, {dt:.1f} s wall, oracle {build}"}
        out["C4"] = ent
    T4.close()

    # ---- C5: 64 frames of the 25x25 stress mesh
    if rank == 0:
        c5 = synthetic.CONFIGS["C5"]
        tmpl5, base5 = synthetic.make_config_frames("C5", nframes=8)
        frames5 = [base5[i % 8] for i in range(c5["B"])]
        T5 = sft.Template(tmpl5)
        rb5 = sft.ResidentBatch(frames5, template=T5)
        hb5 = sft.HostBatch(frames5, template=T5)
        for _ in range(2):
            rb5.run(); hb5.solve()
        ks, es = [], []
        for _ in range(reps):
            flush_fn()
            ks.append(rb5.run())
            flush_fn()
            t0 = time.perf_counter()
            hb5.solve()
            es.append(1e3 * (time.perf_counter() - t0))
        o5 = rb5.fetch()
        info5 = rb5.info()
        k5, e5 = float(np.median(ks)), float(np.median(es))
        it5 = float(np.sum([o.r.lm_iterations for o in o5]))
        tr5 = float(np.sum([o.r.lm_trials for o in o5]))
        peaks, peak_kind = measured_peaks()
        b_iter5 = algorithmic_bytes_per_iteration(c5["G"], c5["M"])
        ach5 = it5 * b_iter5 / (k5 * 1e-3) / 1e9
        ent = {"workload": workload_string("C5"), "frames": c5["B"], "distinct_frames": 8,
               "batch_latency_ms": k5, "value": c5["B"] / (k5 * 1e-3), "unit": "solves/s",
               "e2e": {"batch_latency_ms": e5, "value": c5["B"] / (e5 * 1e-3), "unit": "solves/s"},
               "grid": info5["grid"], "threads": info5["threads"], "smem_bytes": info5["smem_bytes"],
               "lm_iterations_per_frame": it5 / c5["B"], "lm_trials_per_frame": tr5 / c5["B"],
               "roofline": {"bound": "hbm", "achieved": ach5, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": ach5 / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_kind,
                            "algorithmic_bytes_per_lm_iteration_per_frame": b_iter5,
                            "fp64": fp64_roofline(c5["G"], tr5, it5, k5 * 1e-3, None),
                            "note": "64 frames occupy 64 of 148 SMs (one CTA per frame, one CTA per SM at this size)"}}
        if olib is not None:
            cores = os.cpu_count() or 1
            sample = [base5[i % 8] for i in range(min(cores, 16))]
            rate, dt, build = cpu_solve_rate(sample, min(cores, 16))
            ent["cpu_baseline"] = {"value": rate, "unit": "solves/s", "cores": min(cores, 16), "kind": "port",
                                   "sample": f"{len(sample)} frames, one frame per thread, {dt:.1f} s wall, oracle {build}"}
        out["C5"] = ent
        rb5.close()
        T5.close()
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    c = __import__("defslam_b200.synthetic", fromlist=["CONFIGS"]).CONFIGS[CFG]
    cores = os.cpu_count() or 1
    n_sample = max(cores, min(CPU_SAMPLE_FRAMES, 4 * cores))
    _, base, _ = make_frames(0)
    sample = [base[i % N_DISTINCT] for i in range(n_sample)]
    for _ in range(args.warmup):
        cpu_solve_rate(sample[:cores], cores)
    rates, secs = [], []
    for _ in range(args.steps):
        r, dt, build = cpu_solve_rate(sample, cores)
        rates.append(r); secs.append(dt)
    value = len(sample) * args.steps / sum(secs)
    line = {
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(CFG), "frames_per_step": n_sample,
                   "sample": "bounded sample of the batch the GPU arm solves per step"},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port",
                         "sample": f"{n_sample} frames/step x {args.steps} steps, one frame per thread, oracle {build}"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU oracle = restatement of the reference's g2o LM + dense LDLT path; the reference itself "
                "needs Eigen/OpenCV/Ceres, absent from this image",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nrsfm", action="store_true", help="skip the NRSfM stage measurements")
    ap.add_argument("--no-stream", action="store_true", help="skip the C3 tracking+mapping stream")
    ap.add_argument("--no-matching", action="store_true", help="skip the match-production measurements")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config measurements (C1, C2 B=1, C4, C5)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from defslam_b200 import _capi, sft, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _capi.load()
    c = synthetic.CONFIGS[CFG]

    tmpl, base, frames = make_frames(args.frames)
    T = sft.Template(tmpl)
    rb = sft.ResidentBatch(frames, template=T)
    info = rb.info()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input throughput (value) ------------------------------------------
    for _ in range(args.warmup):
        flush.fill_(1)
        rb.run()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.defslam_kernel_launch_count()
    t_wall0 = time.perf_counter()
    kernel_ms = []
    for _ in range(args.steps):
        flush.fill_(1)           # L2 flush between timed iterations (not part of the step time)
        torch.cuda.synchronize()
        kernel_ms.append(rb.run())   # CUDA events around the launch, on the library's stream
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = lib.defslam_kernel_launch_count() - launches0
    outs = rb.fetch()
    iters = float(np.sum([o.r.lm_iterations for o in outs]))
    trials = float(np.sum([o.r.lm_trials for o in outs]))
    total_ms = float(np.sum(kernel_ms))

    # ---- end to end through the C ABI on host buffers (e2e) --------------------------
    hb = sft.HostBatch(frames, template=T)   # descriptors over the host arrays, built once
    for _ in range(min(args.warmup, 2)):
        hb.solve()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_outs = hb.solve()   # ONE C-ABI call: marshal to pinned + H2D + kernel + D2H + scatter
        e2e_s += time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()   # sampled over both timed regions (resident launches and end-to-end calls)
    assert all(o.r.status == 0 for o in e2e_outs)

    # ---- NRSfM stages (same timing rules; reported next to the headline metric) ---------
    nr = None
    if not args.no_nrsfm:
        def _flush():
            flush.fill_(1)
            torch.cuda.synchronize()
        wl = nrsfm_workload()
        barrier()
        l0 = lib.defslam_kernel_launch_count()
        nr = nrsfm_gpu(lib, wl, max(2, args.steps // 2), args.warmup, _flush)
        nrsfm_launches = lib.defslam_kernel_launch_count() - l0
        barrier()

    # ---- every BASELINE.json config (C1, C2 at B=1, C4 sharded/strong, C5): all ranks take part in C4
    cfg_line = None
    if not args.no_configs:
        def _flush2():
            flush.fill_(1)
            torch.cuda.synchronize()
        cfg_line = configs_bench(lib, args, rank, world, dist, _flush2, not args.no_cpu_baseline)
        barrier()

    # ---- C3: tracking + mapping loop over a 300-frame stream (rank 0; latency-bound: one frame at a time)
    st_line = None
    if not args.no_stream and rank == 0:
        st_line = stream_bench(lib, not args.no_cpu_baseline)

    # ---- match production (projection search, warp-guided search): per-frame calls on host buffers
    mt_line = None
    if not args.no_matching and rank == 0:
        mt_line = matching_bench(lib, not args.no_cpu_baseline)

    # ---- parity spot check against the oracle (not timed) ----------------------------
    rel = None
    if rank == 0:
        from oracle import oracle_py
        ref = oracle_py.sft_solve(frames[0])
        rel = float(np.abs(outs[0].nodes - ref.nodes).max() / np.sqrt((ref.nodes ** 2).sum(1).mean()))

    # ---- max over ranks ---------------------------------------------------------------
    t = torch.tensor([total_ms, e2e_s * 1e3, iters, trials], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms = float(tmax[0]), float(tmax[1])
        iters_all, trials_all = float(tsum[2]), float(tsum[3])
    else:
        e2e_ms = e2e_s * 1e3
        iters_all, trials_all = iters, trials

    nr_line = None
    if nr is not None:
        names = sorted(nr)
        tt = torch.tensor([[nr[k][1], nr[k][2] * 1e3] for k in names], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        nsteps = max(2, args.steps // 2)
        nr_line = {k: {"units_per_step_per_gpu": nr[k][0],
                       "value": nr[k][0] * world * nsteps / (float(tt[i, 0]) * 1e-3),
                       "e2e": nr[k][0] * world * nsteps / (float(tt[i, 1]) * 1e-3)} for i, k in enumerate(names)}

    if rank == 0:
        n_frames_job = args.frames * world
        value = n_frames_job * args.steps / (total_ms * 1e-3)
        e2e_value = n_frames_job * args.steps / (e2e_ms * 1e-3)
        peaks, peak_kind = measured_peaks()
        # roofline of the LM kernel on this rank (per launch)
        b_iter = algorithmic_bytes_per_iteration(c["G"], c["M"])
        alg_bytes = iters * b_iter
        avg_launch_s = (float(np.sum(kernel_ms)) / args.steps) * 1e-3
        achieved = alg_bytes / avg_launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("frames") == args.frames and tj.get("workload") == CFG:
                traffic = tj.get("dram_bytes_per_launch")
        line = {
            "metric": METRIC,
            "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(CFG), "frames_per_step_per_gpu": args.frames,
                       "distinct_frames": N_DISTINCT,
                       "l2": "flushed between timed iterations (256 MB write)",
                       "grid": info["grid"], "threads": info["threads"], "smem_bytes": info["smem_bytes"],
                       "lm_iterations_per_frame": iters_all / n_frames_job,
                       "lm_trials_per_frame": trials_all / n_frames_job,
                       "node_rel_err_vs_oracle": rel, "wall_s_timed_region": t_wall},
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": info["h2d_bytes"],
                    "d2h_bytes_per_step": info["d2h_bytes"]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
                         "kernel": "sft_lm_kernel (persistent LM solve, one CTA per frame)",
                         "algorithmic_bytes_per_lm_iteration_per_frame": b_iter,
                         "note": "the fused LM kernel keeps the whole LM loop on chip: it is bound by shared-memory "
                                 "bandwidth and the FP64 dependency chain of the banded factorisation, not by HBM; "
                                 "the fp64 object gives the compute view; see DESIGN.md",
                         "fp64": fp64_roofline(c["G"], trials, iters, avg_launch_s, clocks)},
        }
        if nr_line is not None:
            units = {"schwarp_fit": "keyframe-pair fits/s", "normals": "map-point normals/s", "sfn": "keyframe solves/s"}
            for k in nr_line:
                nr_line[k]["unit"] = units[k]
            # per-stage rooflines (resident rate x algorithmic work per unit)
            mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
            fp64_peak = 64 * 2 * 148 * mhz * 1e6 / 1e12
            NPs, bws = 2 * 13 * 15, 3 * 30 + 29      # Schwarp unknowns; half bandwidth of the block-banded normal matrix
            fl_fit = 4 * (NPs * bws * bws + 4 * NPs * bws)   # Warp::initialize solve + 3 LM steps, factor + two sweeps each
            a = nr_line["schwarp_fit"]["value"] / world * fl_fit / 1e12
            nr_line["schwarp_fit"]["roofline"] = {"bound": "fp64", "achieved": a, "peak": fp64_peak, "unit": "TFLOP/s",
                                                  "frac": a / fp64_peak, "flops_per_fit": fl_fit,
                                                  "kernel": "schwarp_fit_kernel (block-banded Cholesky, 390 unknowns)"}
            nb = wl["normals"]
            by_pt = (86.0 * nb.npairs + 89.0 * nb.n) / nb.n   # 73 B in + 13 B out per pair, 24 B in + 65 B out per point
            a = nr_line["normals"]["value"] / world * by_pt / 1e9
            nr_line["normals"]["roofline"] = {"bound": "hbm", "achieved": a, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                              "frac": a / peaks["hbm_gbs"], "bytes_per_point": by_pt,
                                              "kernel": "normals_kernel (thread per map point)"}
            NCs = 13 * 15
            n_nrm = float(np.mean([len(k.uv) for k in wl["keyframes_base"]]))
            fl_kf = NCs ** 3 / 3.0 + 2 * 2 * n_nrm * 256 + 5 * 2 * NCs * NCs   # packed Cholesky + M'M gather + solves/residual sweeps
            a = nr_line["sfn"]["value"] / world * fl_kf / 1e12
            nr_line["sfn"]["roofline"] = {"bound": "fp64", "achieved": a, "peak": fp64_peak, "unit": "TFLOP/s",
                                          "frac": a / fp64_peak, "flops_per_keyframe": fl_kf,
                                          "kernel": "sfn_solve_kernel (packed Cholesky, 195 unknowns)"}
            line["nrsfm"] = {"stages": nr_line, "gpu_launches": int(nrsfm_launches),
                             "workload": f"{NRSFM_WINDOWS} distinct synthetic keyframe windows (1200 keypoints, 4 views, "
                                         "13x15 control grid) tiled; value = units / device time of the stage kernel "
                                         "(CUDA events), e2e = units / wall time of the C-ABI call on host buffers"}
            if not args.no_cpu_baseline:
                cpu = nrsfm_cpu(wl, os.cpu_count() or 1)
                for k, v in cpu.items():
                    line["nrsfm"]["stages"][k]["cpu_baseline"] = v
                line["nrsfm"]["cpu_baseline"] = {"cores": os.cpu_count() or 1, "kind": "port",
                                                 "sample": "32 fits / 256 point sets / 16 keyframes, one unit per thread"}
        if cfg_line:
            line["configs"] = cfg_line
        if st_line is not None:
            line["stream"] = st_line
        if mt_line is not None:
            line["matching"] = mt_line
        if not args.no_cpu_baseline and world >= 1:
            cores = os.cpu_count() or 1
            n_sample = max(cores, min(CPU_SAMPLE_FRAMES, 8 * cores))
            sample = [base[i % N_DISTINCT] for i in range(n_sample)]
            rate, dt, build = cpu_solve_rate(sample, cores)
            line["cpu_baseline"] = {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_sample} frames of the same workload, one frame per thread, "
                                              f"{dt:.1f} s wall, oracle {build}"}
        print(json.dumps(line), flush=True)
    rb.close()
    T.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
