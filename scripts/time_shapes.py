"""Kernel-only time of tc_conv / tc_wgrad for a list of shapes (50 back-to-back launches between CUDA events, L2-warm).
usage: time_shapes.py "B H W Cin Cout ks" ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from hcmoco_b200.kernels import CudaKernels  # noqa: E402

K = CudaKernels()
for spec in sys.argv[1:]:
    B, H, W, Cin, Cout, ks = [int(v) for v in spec.split()]
    x = torch.randn(B, H, W, Cin, device="cuda")
    w = torch.randn(Cout, Cin, ks, ks, device="cuda") * 0.1
    dy = torch.randn(B, H, W, Cout, device="cuda")
    y = torch.empty(B, H, W, Cout, device="cuda")
    dw = torch.zeros_like(w)
    sc, sh = torch.rand(Cin, device="cuda") + 0.5, torch.randn(Cin, device="cuda")
    wp = torch.zeros((K.tc_conv_wpack_bytes(B, H, W, Cin, Cout, ks) + 3) // 4, device="cuda")
    K.tc_conv_pack(w, 0, wp, B, H, W, Cin, Cout, ks, 4 * K.tc_conv_rowcat_supported(Cout, ks, 1))
    res = []
    for fn in (lambda: K.tc_conv(x, wp, None, y, B, H, W, Cin, Cout, ks, 1, sc, sh, 1, 0),
               lambda: K.tc_wgrad(x, dy, dw, 0, B, H, W, Cin, Cout, ks, 1, sc, sh, 1)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 50 * 1e3)
    fl = 2.0 * B * H * W * Cin * Cout * ks * ks
    print("%-22s tc_conv %7.1f us (%6.1f TF/s useful)   tc_wgrad %7.1f us (%6.1f TF/s)" % (spec, res[0], fl / res[0] / 1e6, res[1], fl / res[1] / 1e6), flush=True)
