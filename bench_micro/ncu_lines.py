"""Per-source-line share of stall samples and executed instructions of the first kernel in a .ncu-rep (development aid).
Usage: ncu_lines.py report.ncu-rep [min_percent]"""
import csv, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data, kernels = [], 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1:
            break
    if len(r) > 8 and r[0].isdigit() and r[2] == "-":
        try:
            data.append((int(r[0]), r[1], int(r[4] or 0), int(r[7] or 0)))
        except ValueError:
            pass
ts, ti = sum(d[2] for d in data) or 1, sum(d[3] for d in data) or 1
print(f"samples {ts}  warp-instructions {ti}")
for d in data:
    if d[2] / ts * 100 >= thr or d[3] / ti * 100 >= thr:
        print(f"{d[0]:5d} {d[2] / ts * 100:5.1f}% smp {d[3] / ti * 100:5.1f}% ins  {d[1].strip()[:120]}")
