"""Shared harness: run the engine (any kernel backend) and the CPU oracle on the same synthetic steps and
compare everything the step produces.  Used by the CPU host-logic tests (TorchKernels, float64: exact
wiring check) and the GPU parity tests (CudaKernels, fp32).

Gradient bar.  Through ~300 train-mode BatchNorms at small batch the fp32 gradient of this network is
ill-conditioned: the oracle (and the reference) in fp32 differ from the same computation in fp64 by
~1e-2 relative (measured: B=3,R=64 1.0e-2; B=4,R=128 9.6e-3), for ANY summation order.  Comparing two
fp32 implementations with each other therefore says nothing below that level; the GPU tests compare the
engine's gradient with the fp64 oracle and require its error to stay within max(`gfactor` x the fp32
oracle's own error against fp64, `gfloor`).  Measured on B200 (global relative gradient error vs fp64,
B=3 R=64 / B=2 R=224): fp32 oracle 1.05e-2 / 1.2e-2; engine with exact-fp32 SIMT convolutions 1.5e-2;
the same engine program executed with PyTorch's own fp32 CUDA ops 1.6e-2; engine with the bf16-split
tensor-core convolutions (~1e-5 per conv instead of 1e-7) 5.3e-2.  The north-star bar (1e-3) is on
embeddings and losses, where the engine is at 1e-5.  Forward embeddings, feature maps and all losses are compared with the
fp32 oracle directly at `tol` (north star: 1e-3 relative).
"""
import torch

from oracle import hcmoco_oracle as O
from synth import synthetic_banks, synthetic_state
from hcmoco_b200.engine import Engine
from hcmoco_b200.synthetic import make_batch, make_dense_idx, make_nce_idx


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def _cast(t, dtype):
    return t.to(dtype) if t.is_floating_point() else t


def global_grad_err(g, ref):
    num = den = 0.0
    worst = ("", 0.0)
    for k, v in ref.items():
        d = float((g[k].cpu().double() - v.double()).pow(2).sum())
        n = float(v.double().pow(2).sum())
        num += d
        den += n
        if n > 1e-16 and (d / n) ** 0.5 > worst[1]:
            worst = (k, (d / n) ** 0.5)
    return (num / den) ** 0.5, worst


def make_inputs(cfg, s, dtype=torch.float32):
    J = 16 if cfg["skeleton"] == "mpii" else 13
    d = [_cast(t, dtype) for t in make_batch(cfg["B"], cfg["R"], J, cfg["n"], seed=1234 + s)]
    nce = make_nce_idx(cfg["B"], cfg["K"], cfg["n"], d[1], seed=99 + s)
    dense = make_dense_idx(d[7], cfg["R"] // 4, cfg["S"], seed=7 + s)
    batch = dict(x=d[0], index=d[1], skeleton=d[2], joints_yx=d[4], joints_vis=d[5], use_depth=d[6], depth_mask=d[7])
    return batch, nce, dense


def oracle_state(cfg, dtype):
    layout = O.model_layout(cfg["width"], cfg["stage"], cfg["skeleton"])
    P = synthetic_state(layout, 0)
    P = type(P)((k, _cast(v, dtype)) for k, v in P.items())
    return layout, P, O.make_momentum(P), [b.to(dtype) for b in synthetic_banks(cfg["n"], 128, 0)]


def compare_forward(eng, res, ref, cfg, tol, verbose, tag=""):
    errs = dict(f=rel(eng.f, ref["f"]), nce=rel(res["nce_losses"], torch.stack(ref["nce_losses"])))
    if cfg["stage"] == 2:
        errs["lm1"] = rel(eng.nchw(eng.lm1), ref["linear_merge1"])
        errs["lm2"] = rel(eng.nchw(eng.lm2), ref["linear_merge2"])
        errs["feat3"] = rel(eng.feat3, ref["feat3"])
        errs["dense"] = rel(res["dense_losses"], torch.stack(ref["dense_losses"]))
        errs["joint"] = rel(res["joint_losses"], torch.stack(ref["joint_losses"]))
        errs["scl"] = rel(res["scl_loss"], ref["scl_loss"])
    errs["loss"] = rel(res["loss"], ref["loss"])
    if verbose:
        print(tag, {k: "%.2e" % v for k, v in errs.items()}, flush=True)
    for k, v in errs.items():
        assert v < tol, (tag, k, v)
    # accuracies are hit counts: identical unless a near-tie flips
    pairs = [(res["nce_accs"], ref["nce_accs"], 100.0 / cfg["B"])]
    if cfg["stage"] == 2:
        pairs += [(res["dense_accs"], ref["dense_accs"], 2.0 / cfg["S"]), (res["joint_accs"], ref["joint_accs"], 0.1)]
    for a, b, slack in pairs:
        b = torch.stack([torch.as_tensor(x, dtype=torch.float64) for x in b])
        assert float((a.double().cpu() - b).abs().max()) <= slack + 1e-9, (a, b)
    return errs


def run_case(K, cfg, nsteps=2, tol=1e-3, gtol=None, gfactor=10.0, gfloor=1e-2, verbose=False, dtype=torch.float32,
             resync=False, use_tc=True):
    """dtype: precision the reference ORACLE (and the synthetic inputs) run in.
    gtol given  -> gradients / final state are compared with that oracle at gtol (exact-wiring mode);
    gtol None   -> gradients are compared with a second, fp64 oracle under the bar described above.
    resync      -> before every step the engine state is reloaded from the oracle (single-step parity)."""
    layout, P, mom, banks = oracle_state(cfg, dtype)
    truth = oracle_state(cfg, torch.float64) if gtol is None else None
    eng = Engine(K, cfg["width"], cfg["stage"], cfg["skeleton"], cfg["B"], cfg["R"], cfg["n"], cfg["K"],
                 num_samples=cfg["S"], use_tc=use_tc)
    assert list(eng.store.keys.keys()) == list(layout.keys())
    eng.store.load_state_dict(P)
    eng.init_banks(banks)
    eng.build()
    report = {}
    for s in range(nsteps):
        batch, nce, dense = make_inputs(cfg, s, dtype)
        if resync and s > 0:
            eng.store.load_state_dict(P)
            for k in mom:
                eng.store.load(eng.store.m, k, mom[k])
            eng.init_banks(banks)
            eng.first_step = False
            if truth is not None:
                for k in P:
                    truth[1][k].copy_(P[k])
                for k in mom:
                    truth[2][k].copy_(mom[k])
                for m in range(3):
                    truth[3][m].copy_(banks[m])
        ref = O.train_step(P, mom, banks, batch, nce, dense, width=cfg["width"], skeleton=cfg["skeleton"],
                           stage=cfg["stage"], first=(s == 0))
        eng.set_batch(batch, nce, dense)
        eng.forward()
        eng.backward()
        res = eng.results()
        errs = compare_forward(eng, res, ref, cfg, tol, verbose, "step %d" % s)
        g = eng.store.grads_dict()
        if gtol is not None:
            ge, worst = global_grad_err(g, ref["grads"])
            bar = gtol
        else:
            lay64, P64, mom64, banks64 = truth
            b64, _, _ = make_inputs(cfg, s, torch.float64)
            t = O.train_step(P64, mom64, banks64, b64, nce, dense, width=cfg["width"], skeleton=cfg["skeleton"],
                             stage=cfg["stage"], first=(s == 0))
            ge, worst = global_grad_err(g, t["grads"])
            own, _ = global_grad_err(ref["grads"], t["grads"])
            bar = max(gfactor * own, gfloor)
            errs["grad_fp32_oracle_vs_fp64"] = own
        errs["grad_global"], errs["grad_worst"] = ge, worst
        if verbose:
            print("   grads: global %.2e (bar %.2e) worst %s %.2e" % (ge, bar, worst[0], worst[1]), flush=True)
        assert ge < bar, (ge, bar)
        eng.update_banks()
        eng.sgd()
        report[s] = errs
    if gtol is not None:
        # after the last step: parameters, BN statistics, banks against the same-precision oracle
        sd = eng.store.state_dict()
        pw = ("", 0.0)
        for k in layout:
            if k.endswith("num_batches_tracked"):
                assert int(sd[k]) == int(P[k]), k
                continue
            e = rel(sd[k], P[k])
            if e > pw[1]:
                pw = (k, e)
        if verbose:
            print("params worst", pw)
        assert pw[1] < gtol, pw
        for m in range(3):
            assert rel(eng.banks[m], banks[m]) < max(gtol, 1e-6)
    return report
