"""Instruction / stall-sample totals of an ncu report grouped by source-line ranges of sft_core.h.
usage: python tools/ncu_regions.py report.ncu-rep"""
import csv
import subprocess
import sys
import re

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
src = open("defslam_b200/csrc/sft_core.h").read().splitlines()
# function start lines
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(DS_FN|DS_FN_NOINLINE|template <bool CHECK>)", l)
    if m:
        name = re.search(r"(\w+)\(", l if "(" in l else src[i])
        starts.append((i, name.group(1) if name else l[:30]))
cur_file = None
agg = {}
tot_s = tot_i = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0].isdigit() and len(r) > 8 and r[2] == "-":
        try:
            samples = int(r[4]); inst = int(r[7])
        except ValueError:
            continue
        ln = int(r[0])
        key = cur_file
        if cur_file == "sft_core.h":
            fn = "?"
            for s, n in starts:
                if s <= ln:
                    fn = n
            key = f"sft_core.h:{fn}"
            if fn == "factor_solve":
                key += ":" + str(ln // 25 * 25)
        a = agg.setdefault(key, [0, 0])
        a[0] += samples; a[1] += inst
        tot_s += samples; tot_i += inst
print("total samples", tot_s, "inst", tot_i)
for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100.0*s/tot_s:6.2f}% samples  {100.0*i/tot_i:6.2f}% inst  {k}")
