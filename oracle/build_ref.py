"""Build oracle/_ref/libpn2_ref.so: the REFERENCE's own PointNet++ CUDA kernels (the only native code of hongfz16/HCMoCo:
pycontrast/networks/pointnet2/src/{ball_query,group_points,interpolate,sampling}_gpu.cu), compiled with nvcc for sm_100a from the
sources WHERE THEY LIE under /root/reference, plus the C shim oracle/pn2_ref_shim.cu.  TEST INFRASTRUCTURE: the GPU tests compare
hcm_pn2_* with these kernels bit for bit.  Outputs go to oracle/_ref/ only (git-ignored, travels to the GPU box with the snapshot).

The reference's build (networks/pointnet2/setup.py: torch CUDAExtension + pybind wrappers over THC, which torch 2.x no longer ships)
is not run: only the four kernel files are compiled; their headers need the torch include tree for the at::Tensor prototypes, nothing
from libtorch is linked (checked: no undefined torch symbols).  Without /root/reference (the GPU box) this is a no-op and the
prebuilt library is used; without the library the reference-comparison tests skip and the numpy oracle (oracle/pn2_oracle.py) pins
the kernels alone."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("HCMOCO_REFERENCE", "/root/reference") + "/pycontrast/networks/pointnet2/src"
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libpn2_ref.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FILES = ["ball_query_gpu.cu", "group_points_gpu.cu", "interpolate_gpu.cu", "sampling_gpu.cu"]


def build(force=False, verbose=False):
    if not os.path.isdir(SRC):
        return LIB if os.path.exists(LIB) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(SRC, f) for f in FILES] + [os.path.join(HERE, "pn2_ref_shim.cu")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    from torch.utils import cpp_extension as ce
    inc = sum([["-I", i] for i in ce.include_paths()], []) + ["-I", SRC]
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-w"]
    procs, objs = [], []
    for s in srcs:
        o = os.path.join(OUT, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        cmd = [NVCC] + flags + inc + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed on %s" % s)
    subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"])
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
