"""tc_conv with accumulate = 0 / 1 (the data gradient of a layer whose input has a second gradient contribution: every
BasicBlock / Bottleneck conv1), L2-warm and L2-cold.  usage: time_acc.py "B H W Cin Cout ks" ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from hcmoco_b200.kernels import CudaKernels  # noqa: E402

K = CudaKernels()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for spec in sys.argv[1:]:
    B, H, W, Cin, Cout, ks = [int(v) for v in spec.split()]
    x = torch.randn(B, H, W, Cin, device="cuda")
    w = torch.randn(Cout, Cin, ks, ks, device="cuda") * 0.1
    y = torch.zeros(B, H, W, Cout, device="cuda")
    wp = torch.zeros((K.tc_conv_wpack_bytes(B, H, W, Cin, Cout, ks) + 3) // 4, device="cuda")
    K.tc_conv_pack(w, 0, wp, B, H, W, Cin, Cout, ks, 0)
    out = []
    for acc in (0, 1):
        for cold in (0, 1):
            ts = []
            for it in range(8):
                if cold:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                K.tc_conv(x, wp, None, y, B, H, W, Cin, Cout, ks, 1, None, None, 0, acc)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            out.append("acc=%d %s %6.1f us" % (acc, "cold" if cold else "warm", ts[len(ts) // 2]))
    print("%-22s %s" % (spec, "   ".join(out)), flush=True)
