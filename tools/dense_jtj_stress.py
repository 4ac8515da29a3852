"""Config C5's side question (BASELINE.json: "fp32 vs fp16 J^T J tensor-core path"): what would the normal
matrix cost if it were formed as a dense GEMM H = J^T W J on the tensor cores instead of the structured
assembly the LM kernel does?  Reported for DESIGN.md, not shipped: the dense product does ~2 R D^2 flops to
produce what the sparse build gets in ~1 MFLOP per iteration.

  python tools/dense_jtj_stress.py        (on a B200; library GEMMs through torch, CUDA events)
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from defslam_b200 import sft, synthetic  # noqa: E402


def main():
    G, M, B = 25, 2000, 64
    n = G * G
    D = 3 * n + 6
    n_int, E = (G - 2) ** 2, 3 * G * G - 4 * G + 1
    R = 2 * M + 6 * n_int + E + 3 * n          # reprojection (2 rows), curvature copies, stretch, temporal
    dev = torch.device("cuda")
    out = {"shape": {"G": G, "matches": M, "rows": R, "D": D, "frames": B},
           "dense_flops_per_frame": 2.0 * R * D * D}
    g = torch.Generator(device=dev).manual_seed(0)
    for name, dt in (("fp64", torch.float64), ("fp32", torch.float32), ("tf32", torch.float32), ("fp16", torch.float16),
                     ("bf16", torch.bfloat16)):
        torch.backends.cuda.matmul.allow_tf32 = name == "tf32"
        J = (torch.rand((B, R, D), generator=g, device=dev, dtype=torch.float32) < 0.01).to(dt)   # sparse pattern, dense storage
        for _ in range(2):
            H = torch.bmm(J.transpose(1, 2), J)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        reps = 3
        for _ in range(reps):
            H = torch.bmm(J.transpose(1, 2), J)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = {"ms_per_frame": ms / B, "tflops": 2.0 * R * D * D * B / (ms * 1e-3) / 1e12}
        del J, H
    # the structured path: the whole LM solve of the same frames (assembly + banded factorisations + LM loop)
    tmpl, frames = synthetic.make_config_frames("C5", nframes=4)
    frames = [frames[i % 4] for i in range(148)]
    rb = sft.ResidentBatch(frames, template=sft.Template(tmpl))
    rb.run()
    ms = min(rb.run() for _ in range(3))
    outs = rb.fetch()
    iters = sum(o.r.lm_iterations for o in outs) / len(outs)
    out["structured_lm_solve"] = {"ms_per_frame_whole_solve": ms / 148 , "lm_iterations_per_frame": iters,
                                  "note": "148 frames resident, one launch; includes every build, factorisation and trial"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
