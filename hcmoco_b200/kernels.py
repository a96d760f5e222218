"""Tensor-level binding of the C-ABI (include/hcmoco.h) for the host engine.

`CudaKernels` exposes one method per `hcm_*` entry point (name without the prefix).  Arguments are
the C arguments in order, with torch CUDA tensors (or None) where the C signature has a pointer;
the trailing `cudaStream_t` is supplied from torch's current stream, so everything composes with
torch streams and CUDA-graph capture.  There is no CPU / PyTorch fallback: constructing
`CudaKernels` without the built library or without a CUDA device raises.
"""
import ctypes

import torch

from . import _lib


class CudaKernels:
    name = "cuda"

    def __init__(self):
        if not torch.cuda.is_available():
            raise _lib.HcmError("hcmoco_b200 needs a CUDA device: the compute path is sm_100a CUDA only "
                                "(no CPU / PyTorch fallback)")
        self.lib = _lib.load()
        self.launches = 0
        for name, _ret, args in _lib.parse_header():
            if name in ("hcm_abi_version", "hcm_last_error"):
                continue
            setattr(self, name[4:], self._make(name, args))

    def _make(self, name, args):
        fn = getattr(self.lib, name)
        kinds = []
        for ct, an in args:
            if an == "stream":
                kinds.append("stream")
            elif ct is ctypes.POINTER(ctypes.c_void_p):
                kinds.append("ptrlist")
            elif ct is ctypes.c_void_p:
                kinds.append("ptr")
            elif ct is ctypes.c_float or ct is ctypes.c_double:
                kinds.append("float")
            else:
                kinds.append("int")
        nuser = sum(1 for k in kinds if k != "stream")
        returns_rows = name.endswith(("_rows", "_supported", "_bytes", "_nqs"))     # queries: return the value, launch nothing

        def call(*a):
            if len(a) != nuser:
                raise TypeError("%s expects %d arguments, got %d" % (name, nuser, len(a)))
            conv, keep, it = [], [], iter(a)
            for k in kinds:
                if k == "stream":
                    conv.append(torch.cuda.current_stream().cuda_stream)
                    continue
                v = next(it)
                if k == "ptr":
                    conv.append(None if v is None else v.data_ptr())
                elif k == "ptrlist":
                    if v is None:
                        conv.append(None)
                    else:
                        arr = (ctypes.c_void_p * len(v))(*[None if t is None else t.data_ptr() for t in v])
                        keep.append(arr)
                        conv.append(ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p)))
                elif k == "float":
                    conv.append(float(v))
                else:
                    conv.append(int(v))
            rc = fn(*conv)
            if returns_rows:
                return rc
            if rc != 0:
                raise _lib.HcmError("%s failed (%d): %s" % (name, rc, self.lib.hcm_last_error().decode()))
            self.launches += 1
            return 0

        call.__name__ = name
        return call

    # allocation helpers the engine uses (device memory is torch's)
    device = "cuda"

    dtype = torch.float32

    def empty(self, *shape, dtype=None):
        return torch.empty(*shape, dtype=dtype or torch.float32, device="cuda")

    def zeros(self, *shape, dtype=None):
        return torch.zeros(*shape, dtype=dtype or torch.float32, device="cuda")
