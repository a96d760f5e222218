"""Per-source-line instruction counts and stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python scripts/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, data = None, None, []
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":          # a source line summary row
        try:
            data.append((int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), cur, r[0], r[1]))
        except ValueError:
            pass
tot, tots = sum(d[0] for d in data), sum(d[1] for d in data)
print("total warp instructions %d, stall samples %d" % (tot, tots))
for n, s, f, ln, src in sorted(data, reverse=True)[:top]:
    print("%9d %5.1f%%  smp %5.1f%% | %s:%s | %s" % (n, 100.0 * n / tot, 100.0 * s / max(tots, 1), f, ln, src.strip()[:100]))
