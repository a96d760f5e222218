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


def rep_rows(path):
    raw = subprocess.run("ncu -i %s --page raw --csv" % path, shell=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    ci = {h: i for i, h in enumerate(rows[0])}
    return rows, ci


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


# (report, kernel-name substring) -> key of profiles/ncu_traffic.json (read by bench.py for roofline.traffic)
TRAFFIC_KEYS = [
    ("tcconv18_b64", "tc_conv_kernel", "tc_conv 64x64 18->18 k3 s1 B=64"),
    ("tcwgrad18_b64", "tc_wgrad2_kernel", "tc_wgrad 64x64 18->18 k3 s1 B=64"),
    ("tcconv144", "tc_conv_kernel", "tc_conv 8x8 144->144 k3 s1 B=64"),
    ("nce_b64", "nce_logits_kernel", "nce_logits B=64"),
    ("nce_b64", "nce_bwd_kernel", "nce_bwd B=64"),
    ("dense_affinity", "dense_affinity_kernel<0>", "dense_affinity_fwd B=32"),
    ("dense_affinity", "dense_affinity_kernel<1>", "dense_affinity_bwd B=32"),
    ("dense_v2", "dense_affinity_kernel<0>", "dense_affinity_fwd B=32"),
    ("dense_v2", "dense_affinity_kernel<1>", "dense_affinity_bwd B=32"),
]

for name in ("launches_b8.csv", "launches_b64.csv"):
    if os.path.exists(os.path.join(go, name)):
        b = name[len("launches_"):-4]
        launches(os.path.join(go, name), os.path.join(out_dir, tag + "_launches_bench_%s.txt" % b),
                 "ONE step of bench.py --batch %s --no-graph between cudaProfilerStart/Stop (bench.py --ncu-step)" % b[1:])
traffic = {}
tj = os.path.join(out_dir, "ncu_traffic.json")
if os.path.exists(tj):
    traffic = json.load(open(tj))
for rep in sorted(glob.glob(os.path.join(go, "*.ncu-rep"))):
    base = os.path.basename(rep).replace(".ncu-rep", "")
    dst = base if base[:1] == "r" and base[1:3].isdigit() else tag + "_" + base
    ncu_rep(rep, os.path.join(out_dir, dst + ".txt"), base)
    rows, ci = rep_rows(rep)
    for frag, kern, key in TRAFFIC_KEYS:
        if frag not in base:
            continue
        for r in rows[2:]:
            kn = r[ci["Kernel Name"]].replace("(bool)", "")
            kn = kn.replace("<true>", "<1>").replace("<false>", "<0>")
            if kern in kn:
                rd = to_bytes(r[ci["dram__bytes_read.sum"]], rows[1][ci["dram__bytes_read.sum"]])
                wr = to_bytes(r[ci["dram__bytes_write.sum"]], rows[1][ci["dram__bytes_write.sum"]])
                traffic[key] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr,
                                "us": float(r[ci["gpu__time_duration.sum"]].replace(",", "")),
                                "source": "profiles/%s.txt (ncu --set full --clock-control none, one launch)" % dst}
                break
json.dump(traffic, open(tj, "w"), indent=1, sort_keys=True)
