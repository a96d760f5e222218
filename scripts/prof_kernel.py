"""Run one conv shape a few times (for `ncu -k regex:tc_conv_kernel` etc.).  usage: prof_kernel.py B H W Cin Cout ks [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from hcmoco_b200.kernels import CudaKernels  # noqa: E402

B, H, W, Cin, Cout, ks = [int(v) for v in sys.argv[1:7]]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
K = CudaKernels()
x = torch.randn(B, H, W, Cin, device="cuda")
w = torch.randn(Cout, Cin, ks, ks, device="cuda") * 0.1
dy = torch.randn(B, H, W, Cout, device="cuda")
y = torch.empty(B, H, W, Cout, device="cuda")
dw = torch.zeros_like(w)
sc, sh = torch.rand(Cin, device="cuda") + 0.5, torch.randn(Cin, device="cuda")
wp = torch.zeros((K.tc_conv_wpack_bytes(B, H, W, Cin, Cout, ks) + 3) // 4, device="cuda")
K.tc_conv_pack(w, 0, wp, B, H, W, Cin, Cout, ks, 4 * K.tc_conv_rowcat_supported(Cout, ks, 1))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for i in range(reps):
    K.tc_conv(x, wp, None, y, B, H, W, Cin, Cout, ks, 1, sc, sh, 1, 0)
    K.tc_wgrad(x, dy, dw, 0, B, H, W, Cin, Cout, ks, 1, sc, sh, 1)
torch.cuda.synchronize()
ev[0].record()
for i in range(20):
    K.tc_conv(x, wp, None, y, B, H, W, Cin, Cout, ks, 1, sc, sh, 1, 0)
ev[1].record()
for i in range(20):
    K.tc_wgrad(x, dy, dw, 0, B, H, W, Cin, Cout, ks, 1, sc, sh, 1)
ev[2].record()
torch.cuda.synchronize()
fl = 2.0 * B * H * W * Cin * Cout * ks * ks
t1, t2 = ev[0].elapsed_time(ev[1]) / 20, ev[1].elapsed_time(ev[2]) / 20
print("tc_conv %.1f us (%.1f TFLOP/s useful)   tc_wgrad %.1f us (%.1f TFLOP/s useful)" % (
    t1 * 1e3, fl / t1 / 1e9, t2 * 1e3, fl / t2 / 1e9))
