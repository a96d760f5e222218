"""The C-ABI boundary without a GPU: the in-tree library builds (nvcc cross-compiles sm_100a), loads, exports every symbol
include/hcmoco.h declares, answers its host-side geometry queries, reports argument errors through the return code +
hcm_last_error() (never exit()), and the product path refuses to run without a CUDA device (no CPU / PyTorch fallback)."""
import ctypes
import os
import re

import pytest
import torch

from hcmoco_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    path = build.build(verbose=False)           # no-op when the objects are up to date
    assert os.path.exists(path)
    return _lib.load()


def test_header_and_library_agree(lib):
    protos = _lib.parse_header()
    names = [n for n, _, _ in protos]
    assert len(names) == len(set(names)) and len(names) >= 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # every entry point takes its stream last (queries and hcm_last_error / hcm_abi_version take none)
    src = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER).read(), flags=re.S)
    for name, _ret, args in protos:
        launches = "cudaStream_t stream" in src[src.index(name + "("):].split(";")[0]
        assert launches == (bool(args) and args[-1][1] == "stream"), name
    assert lib.hcm_abi_version() == 1


def test_host_side_queries(lib):
    # supported geometries of the tensor-core conv / wgrad kernels (HRNet-w18 / w32 shapes at B=64, 256x256)
    for (H, Cin, Cout, ks, stride) in [(64, 18, 18, 3, 1), (32, 36, 36, 3, 1), (16, 72, 72, 3, 1), (8, 144, 144, 3, 1),
                                       (64, 64, 256, 1, 1), (64, 256, 64, 1, 1), (64, 18, 36, 3, 2), (8, 256, 256, 3, 1),
                                       (64, 32, 32, 3, 1)]:
        assert lib.hcm_tc_conv_supported(64, H, H, Cin, Cout, ks, stride) == 1, (H, Cin, Cout, ks, stride)
        assert lib.hcm_tc_wgrad_supported(64, H, H, Cin, Cout, ks, stride) == 1, (H, Cin, Cout, ks, stride)
        assert lib.hcm_tc_conv_wpack_bytes(64, H, H, Cin, Cout, ks) >= Cin * Cout * ks * ks * 4   # hi + lo bf16, padded
    # configs[4] (HRNet-w32, 384x384, B=16 per GPU): every layer but the 3-channel stem runs on the tensor-core kernels
    for (H, Cin, Cout, ks, stride) in [(96, 32, 32, 3, 1), (48, 64, 64, 3, 1), (24, 128, 128, 3, 1), (12, 256, 256, 3, 1),
                                       (96, 64, 256, 1, 1), (96, 256, 32, 3, 1), (96, 256, 64, 3, 2), (192, 64, 64, 3, 2),
                                       (12, 256, 32, 1, 1), (96, 32, 128, 1, 1), (12, 256, 128, 1, 1)]:
        assert lib.hcm_tc_conv_supported(16, H, H, Cin, Cout, ks, stride) == 1, (H, Cin, Cout, ks, stride)
        assert lib.hcm_tc_wgrad_supported(16, H, H, Cin, Cout, ks, stride) == 1, (H, Cin, Cout, ks, stride)
        if stride == 2:
            assert lib.hcm_tc_dgrad_s2_supported(16, H, H, Cin, Cout) == 1
    assert lib.hcm_tc_conv_supported(64, 256, 256, 3, 64, 3, 2) == 0          # odd channel count: SIMT stem kernel instead
    assert lib.hcm_tc_conv_supported(2, 15, 15, 18, 18, 3, 2) == 0            # stride 2 needs even H, W
    assert lib.hcm_tc_conv_supported(2, 16, 16, 18, 18, 5, 1) == 0
    assert lib.hcm_tc_conv_rowcat_supported(18, 3, 1) == 0                    # experimental formulation is off by default
    assert lib.hcm_tc_dgrad_s2_supported(64, 64, 64, 18, 36) == 1 and lib.hcm_tc_dgrad_s2_nqs(64, 64, 64, 18, 36) in (1, 2, 4)
    rows = lib.hcm_colstat_rows(64 * 64 * 64, 18)
    assert 1 <= rows <= 4096 and lib.hcm_colstat_rows(10, 18) >= 1
    assert lib.hcm_conv2d_stat_rows(2, 224, 224, 3, 64, 3, 2) >= 1


def test_argument_errors_are_reported_not_fatal(lib):
    # argument checks run before anything touches the device: negative code + message, process stays alive
    assert lib.hcm_gather_l2norm(None, 0, None, 0, 1, 4, 128, None, 128, None, None) < 0
    assert b"gather_l2norm" in lib.hcm_last_error()
    assert lib.hcm_dense_affinity_fwd(None, None, None, None, None, 2, 400, 64, 128, ctypes.c_float(14.0), None, None, None, None) < 0
    assert lib.hcm_dense_affinity_work_bytes(32, 400) == 32 * 2 * 4 * 2 * 16 * (112 * 16 + 16)
    assert b"dense_affinity_fwd" in lib.hcm_last_error()
    assert lib.hcm_tc_conv(None, None, None, None, 2, 16, 16, 18, 18, 3, 1, None, None, 0, 0, None) < 0
    assert b"tc_conv" in lib.hcm_last_error()
    assert lib.hcm_bn_stats(None, 100, 18, None, None) < 0
    assert lib.hcm_nce_logits(None, None, None, None, None, None, 384, None, 2, 17, 128, ctypes.c_float(0.07), None, None) < 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a CUDA device")
def test_no_cpu_fallback():
    from hcmoco_b200.kernels import CudaKernels
    with pytest.raises(_lib.HcmError):
        CudaKernels()
