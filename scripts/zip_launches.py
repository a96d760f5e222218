"""Join the ncu launch list of one step (ncu --metrics gpu__time_duration.sum --csv) with the host-order conv launch log written by
`HCM_LAUNCH_LOG=... bench.py --ncu-step --no-graph`: true per-shape kernel durations.  usage: zip_launches.py launches.csv log.txt"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
ks = {"tc_conv": "tc_conv_kernel", "tc_wgrad": "tc_wgrad2_kernel", "tc_dgrad_s2": "tc_conv_kernel", "conv2d_fwd": "igemm_kernel",
      "conv2d_wgrad": "igemm_kernel", "conv2d_dgrad": "igemm_kernel"}
launches = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[1:] if any(k in r[kn] for k in ("tc_conv_kernel", "tc_wgrad2_kernel"))]
names = [l.strip() for l in open(sys.argv[2]) if l.strip() and not l.startswith("conv2d")]      # (SIMT igemm launches are not told apart from hcm_gemm's)
agg, i = defaultdict(lambda: [0.0, 0]), 0
for line in names:
    nm, n = line.rsplit("|", 1)
    want = ks[nm.split()[0]]
    for _ in range(int(n)):
        assert want in launches[i][0], (i, nm, launches[i])
        agg[nm][0] += launches[i][1]
        i += 1
    agg[nm][1] += 1
tot = sum(v[0] for v in agg.values())
print("# conv-family launches matched %d of %d, total %.1f us" % (i, len(launches), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%10.1f us %4d calls %8.1f us/call %5.1f%%  %s" % (v[0], v[1], v[0] / v[1], 100 * v[0] / tot, k))
