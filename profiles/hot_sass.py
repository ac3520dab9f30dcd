"""List the SASS instructions with the most warp-stall samples for one kernel of an ncu report.
usage: python profiles/hot_sass.py report.ncu-rep <kernel regex> [top N]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(out)) if len(r) >= 5 and r[0].startswith("0x")]
tot = sum(float(r[2]) for r in rows)
print("total samples", tot, "instructions", len(rows))
idx = sorted(range(len(rows)), key=lambda i: -float(rows[i][2]))[:top]
for i in sorted(idx):
    r = rows[i]
    print(f"{i:5d} {float(r[2]):7.0f} {100 * float(r[2]) / tot:5.1f}%  {r[1][:100]}")
