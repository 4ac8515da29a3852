"""Summarise an ncu report by CUDA source line: stall samples + instructions.
usage: python tools/ncu_lines.py report.ncu-rep [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
lines = []
tot = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if r[0].isdigit() and len(r) > 8 and r[2] == "-":
        try:
            samples = int(r[4]); inst = int(r[7])
        except ValueError:
            continue
        lines.append((samples, inst, cur_file, int(r[0]), r[1].strip()))
        tot += samples
lines.sort(reverse=True)
print(f"total samples {tot}")
for s, i, f, ln, src in lines[:top]:
    print(f"{100.0 * s / max(tot, 1):6.2f}% {s:8d} inst {i:10d}  {f}:{ln}  {src[:110]}")
