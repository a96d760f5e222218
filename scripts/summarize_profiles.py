"""Turn gpurun_out/ artefacts (ncu launch list CSV, ncu-rep files, bench JSON lines) into the small text summaries
committed under profiles/.  usage: python scripts/summarize_profiles.py r01"""
import collections
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
go = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)


def launches(path, dst, title):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    ci = {h: i for i, h in enumerate(rows[hi])}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) < len(ci) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ci["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        v = float(r[ci["Metric Value"]].replace(",", ""))
        v = v / 1000 if r[ci["Metric Unit"]] == "ns" else (v * 1000 if r[ci["Metric Unit"]] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# %s\n# source: ncu --metrics gpu__time_duration.sum --clock-control none (per-launch, cold-cache, serialised: "
                "compare SHARES)\n# total %.1f us over %d launches\n" % (title, tot, sum(v[0] for v in agg.values())))
        f.write("%12s %8s %10s %7s  kernel\n" % ("total_us", "launches", "us/launch", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%12.1f %8d %10.2f %6.1f%%  %s\n" % (v[1], v[0], v[1] / v[0], 100 * v[1] / tot, k[:100]))


def ncu_rep(path, dst, title):
    raw = subprocess.run("ncu -i %s --page raw --csv" % path, shell=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct",
            "gpu__dram_throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__grid_size",
            "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__average_warps_issue_stalled_long_scoreboard_per",
            "sm__throughput.avg.pct", "launch__waves_per_multiprocessor", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
    with open(dst, "w") as f:
        f.write("# %s\n# source: ncu --set full --clock-control none --import-source on (%s)\n" % (title, os.path.basename(path)))
        for k in range(2, len(rows)):
            f.write("## kernel launch %d: %s\n" % (k - 2, rows[k][4] if len(rows[k]) > 4 else ""))
            for h, u, v in zip(rows[0], rows[1], rows[k]):
                if any(h.startswith(w) for w in want):
                    f.write("%-90s %-12s %s\n" % (h, u, v))


if os.path.exists(os.path.join(go, "launches_b8.csv")):
    launches(os.path.join(go, "launches_b8.csv"), os.path.join(out_dir, tag + "_launches_bench_b8.txt"),
             "bench.py --batch 8 --steps 1 --warmup 3 --no-graph (whole process: build, warm-up, timed step, family profile)")
for rep in glob.glob(os.path.join(go, "*.ncu-rep")):
    ncu_rep(rep, os.path.join(out_dir, tag + "_" + os.path.basename(rep).replace(".ncu-rep", ".txt")), os.path.basename(rep))
for name in ("bench_b64.log", "bench_2gpu.log", "detail_b64.txt"):
    p = os.path.join(go, name)
    if os.path.exists(p):
        with open(p) as f, open(os.path.join(out_dir, tag + "_" + name.replace(".log", ".json")), "w") as g:
            g.write(f.read())
print(sorted(os.listdir(out_dir)))
