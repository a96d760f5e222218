"""Deterministic per-key synthetic model state (test infrastructure).

The golden fixtures cannot carry 78 MB of weights, so both sides (the reference run by
make_golden.py and the oracle / CUDA engine in the tests) fill their state from this function.
Distributions follow the reference initialisers (official_hrnet.py:456-463 conv N(0,0.001^2);
nn.Linear default; sem_graph_conv.py:19-30) with BN affine / running statistics perturbed so
that a wrong gamma/beta/statistics path cannot hide.
"""
import math
import zlib

import torch


def _gen(seed, key):
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))


def synthetic_state(layout, seed=0):
    """layout: ordered mapping key -> shape.  Returns an OrderedDict of fp32 / int64 tensors."""
    out = type(layout)()
    for k, shp in layout.items():
        g = _gen(seed, k)
        if k.endswith("num_batches_tracked"):
            t = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            t = 0.05 * torch.randn(shp, generator=g)
        elif k.endswith("running_var"):
            t = 1.0 + 0.2 * torch.rand(shp, generator=g)
        elif k.endswith(".e"):
            t = 1.0 + 0.3 * torch.randn(shp, generator=g)
        elif k.endswith(".W"):
            bound = 1.414 * math.sqrt(6.0 / (shp[1] + shp[2]))
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif len(shp) == 4:
            std = 0.001 if "_linear" not in k else 1.0 / math.sqrt(shp[1])
            t = std * torch.randn(shp, generator=g)
        elif len(shp) == 2:
            t = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(shp[1])
        elif k.endswith("bn.weight") or ".bn" in k and k.endswith(".weight") or _is_bn_weight(k, layout):
            t = 0.75 + 0.5 * torch.rand(shp, generator=g)
        else:  # biases (BN beta, linear / gconv / 1x1-conv bias)
            t = 0.1 * torch.randn(shp, generator=g)
        out[k] = t
    return out


def _is_bn_weight(k, layout):
    return k.endswith(".weight") and (k[:-len("weight")] + "running_mean") in layout


def synthetic_banks(n_data, dim=128, seed=0):
    g = torch.Generator().manual_seed(seed + 4242)
    return [torch.nn.functional.normalize(torch.randn(n_data, dim, generator=g)) for _ in range(3)]
