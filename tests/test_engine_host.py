"""Host-logic tests (CPU): the engine's launch programs — buffer wiring, lazy-BN hand-over, accumulate flags,
backward order, bank update, SGD — executed with the plain-PyTorch statement of each kernel
(tests/kernel_ref.py) and compared with the oracle.  The CUDA kernels themselves are checked in
test_kernels_gpu.py / test_parity_gpu.py."""
import pytest
import torch

from kernel_ref import TorchKernels
from engine_check import run_case

CASES = {
    "stage2_w18_coco": dict(stage=2, width=18, skeleton="coco_reduce", B=3, R=64, K=256, n=1000, S=100),
    "stage1_w18_mpii": dict(stage=1, width=18, skeleton="mpii", B=2, R=64, K=128, n=500, S=100),
    "stage2_w32_mpii": dict(stage=2, width=32, skeleton="mpii", B=2, R=64, K=256, n=1000, S=100),
}


@pytest.mark.parametrize("name", list(CASES))
def test_engine_program_matches_oracle(name):
    # float64 end to end: the engine program and the oracle must agree to rounding
    run_case(TorchKernels("cpu", torch.float64), CASES[name], nsteps=2, tol=1e-9, gtol=1e-8, dtype=torch.float64)
