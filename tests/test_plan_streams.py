"""The multi-stream launch program (Plan.fork / join / b_off_path) on CPU: its markers are well formed, and — the race check —
ANY execution order that respects stream order and the fork / join edges gives bit-identical results.  The program is run
with the plain-PyTorch kernels in float64 once in program order and then under adversarial schedules (side streams first,
main stream first, random); a missing dependency or a scratch buffer shared between streams shows up as a difference."""
import random

import pytest
import torch

from engine_check import make_inputs, oracle_state
from hcmoco_b200.engine import FORK, JOIN, Engine
from kernel_ref import TorchKernels

CFG = dict(stage=2, width=18, skeleton="coco_reduce", B=2, R=64, K=64, n=300, S=40)


def build_engine():
    K = TorchKernels("cpu", torch.float64)
    layout, P, mom, banks = oracle_state(CFG, torch.float64)
    e = Engine(K, CFG["width"], CFG["stage"], CFG["skeleton"], CFG["B"], CFG["R"], CFG["n"], CFG["K"], num_samples=CFG["S"])
    e.store.load_state_dict(P)
    e.init_banks(banks)
    e.build()
    batch, nce, dense = make_inputs(CFG, 0, torch.float64)
    e.set_batch(batch, nce, dense)
    return e


def schedule(prog, policy, seed=0):
    """A valid total order of `prog` under stream semantics: per-tag FIFO queues, FORK = parent records / children wait,
    JOIN = children record / parent waits."""
    queues, nev = {}, 0
    for fn, args, tag in prog:
        if fn is FORK or fn is JOIN:
            parent, children = args
            for c in children:
                ev = nev
                nev += 1
                src, dst = (parent, c) if fn is FORK else (c, parent)
                queues.setdefault(src, []).append(("rec", ev))
                queues.setdefault(dst, []).append(("wait", ev))
        else:
            queues.setdefault(tag, []).append(("op", fn, args))
    heads = {t: 0 for t in queues}
    done, order, rng = set(), [], random.Random(seed)
    while any(heads[t] < len(q) for t, q in queues.items()):
        ready = [t for t, q in queues.items() if heads[t] < len(q) and (q[heads[t]][0] != "wait" or q[heads[t]][1] in done)]
        assert ready, "deadlock: a stream waits for an event nobody records"
        t = {"low": min, "high": max}.get(policy, rng.choice)(ready)
        item = queues[t][heads[t]]
        heads[t] += 1
        if item[0] == "rec":
            done.add(item[1])
        elif item[0] == "op":
            order.append((item[1], item[2]))
    return order


def test_markers_are_well_formed():
    e = build_engine()
    for prog in (e.plan.fwd, e.plan.bwd):
        active, seen = {0}, set()
        for fn, args, tag in prog:
            if fn is FORK:
                parent, children = args
                assert parent in active
                for c in children:
                    assert c not in active or c >= 8, "re-fork of a live branch stream"   # off-path streams are re-forked per launch
                    active.add(c)
                    seen.add(c)
            elif fn is JOIN:
                parent, children = args
                assert parent in active
                for c in children:
                    assert c in active, "join of a stream that was never forked"
                    active.discard(c)
            else:
                assert tag in active, "launch on a stream outside its fork / join region"
        assert active == {0}, "streams left unjoined at the end of the program: %s" % sorted(active - {0})
        assert seen, "the program has no side streams"
    # branches 1..3 of both encoders, the second encoder, and the off-path weight-gradient streams all occur in the backward
    tags = {tag for fn, _, tag in e.plan.bwd if fn is not FORK and fn is not JOIN}
    assert {0, 1} <= tags and any(2 <= t <= 7 for t in tags) and any(t >= 8 for t in tags)


def run_fresh(policy, mutate=None):
    """Forward + backward of a freshly built engine (its `empty` buffers are NaN-filled, so reading a buffer before its
    producer ran cannot go unnoticed) under one schedule; policy None = program order."""
    e = build_engine()
    fwd, bwd = e.plan.fwd, e.plan.bwd
    if mutate is not None:
        fwd, bwd = mutate(fwd, bwd)
    if policy is None:
        of = [(fn, args) for fn, args, _ in fwd if fn is not FORK and fn is not JOIN]
        ob = [(fn, args) for fn, args, _ in bwd if fn is not FORK and fn is not JOIN]
    else:
        seed = {"random1": 1, "random2": 2}.get(policy, 0)
        of, ob = schedule(fwd, policy, seed), schedule(bwd, policy, seed)
    for fn, args in of:
        fn(*args)
    e.K.zero(e.store.g, e.store.n * e.store.g.element_size())
    for fn, args in ob:
        fn(*args)
    return (e.losses.clone(), e.accs.clone(), e.f.clone(), e.store.g.clone(),
            torch.cat([b.reshape(-1).double() for b in e.store.buffers.values()]))


@pytest.fixture(scope="module")
def reference_run():
    ref = run_fresh(None)
    assert float(ref[3].abs().sum()) > 0 and not torch.isnan(ref[3]).any()
    return ref


@pytest.mark.parametrize("policy", ["high", "low", "random1", "random2"])
def test_any_valid_schedule_gives_identical_results(policy, reference_run):
    got = run_fresh(policy)
    for a, b in zip(reference_run, got):
        assert torch.equal(a, b)


def test_the_check_notices_a_missing_join(reference_run):
    """Sensitivity of the check itself: without the join of the first HR module's branches the fuse layers read a branch
    output before it is written (main stream first)."""
    def drop_first_branch_join(fwd, bwd):
        i = next(k for k, (fn, args, _) in enumerate(fwd) if fn is JOIN and args[1] and min(args[1]) >= 2)
        return fwd[:i] + fwd[i + 1:], bwd

    got = run_fresh("low", drop_first_branch_join)
    assert not all(torch.equal(a, b) for a, b in zip(reference_run, got))
