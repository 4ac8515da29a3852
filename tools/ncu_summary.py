"""Text summaries of ncu artefacts for profiles/:
  python tools/ncu_summary.py launches <launches.csv>          per-kernel launch count / time / share
  python tools/ncu_summary.py kernel <report.ncu-rep> [name]   the metrics the DESIGN rooflines quote"""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg",
]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    acc = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)
        name = r[ik].split("(")[0]
        acc[name][0] += 1
        acc[name][1] += v
    tot = sum(v[1] for v in acc.values())
    print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for k, (n, ms) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:8d} {ms:10.3f} {100 * ms / tot:6.1f}%")


def kernel(path, name=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for v in rows[2:]:
        kn = v[hdr.index("Kernel Name")]
        if name and name not in kn:
            continue
        print("kernel:", kn)
        for m in WANT:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:100s} {v[i]:>18s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
