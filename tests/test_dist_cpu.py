"""Two ranks on CPU (gloo): the data-parallel contract of the step —
gradient all-reduce (mean over ranks), embedding/index all-gather feeding the identical bank update on every rank,
bank broadcast of all three banks — executed by the real trainer with the PyTorch statement of the kernels."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = dict(stage=2, width=18, skeleton="mpii", B=2, R=64, K=64, n=300, S=50)


def _worker(rank, world, port, q):
    for p in (os.path.dirname(HERE), HERE, os.path.join(HERE, "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from kernel_ref import TorchKernels
    from oracle import hcmoco_oracle as O
    from engine_check import make_inputs, oracle_state, rel
    from test_api_cpu import make_opt
    from hcmoco_b200 import api
    dt = torch.float64
    K = TorchKernels("cpu", dt)
    opt = make_opt(CFG)
    trainer = api.build_contrast(opt)
    trainer.init_ddp_environment(0, 1)
    assert dist.get_world_size() == world and opt.rank == rank
    layout, P, mom, banks = oracle_state(CFG, dt)
    model, _ = api.build_model(opt, kernels=K)
    model.store.load_state_dict(P)
    mem = api.build_mem(opt, CFG["n"], kernels=K, )
    if rank == 0:
        for i in range(3):
            getattr(mem, "memory_%d" % (i + 1)).copy_(banks[i])
    trainer.broadcast_memory(mem)                      # ranks != 0 start with different banks
    for i in range(3):
        assert rel(getattr(mem, "memory_%d" % (i + 1)), banks[i]) == 0.0
    _, _, optimizer = trainer.wrap_up(model, None, torch.optim.SGD(model.parameters(), lr=0.03, momentum=0.9, weight_decay=1e-4))
    batch, nce, dense = make_inputs(CFG, 10 * rank, dt)          # rank-local shard
    data = [batch["x"], batch["index"], batch["skeleton"], None, batch["joints_yx"], batch["joints_vis"], batch["use_depth"],
            batch["depth_mask"], None]
    mem.injected_idx = nce.clone()
    trainer.injected_dense_idx = dense
    res = trainer.train_step(model, mem, optimizer, data, world)()
    # ---- oracle: every rank's step, gradients averaged, banks updated with the rank-ordered gather
    outs, fs, ys = [], [], []
    for r in range(world):
        b, n_, d_ = make_inputs(CFG, 10 * r, dt)
        Pr = type(P)((k, v.clone()) for k, v in P.items())
        o = O.train_step(Pr, O.make_momentum(Pr), [bk.clone() for bk in banks], b, n_, d_, width=CFG["width"],
                         skeleton=CFG["skeleton"], stage=CFG["stage"], first=True, apply_update=False)
        outs.append(o)
        fs.append(o["f"])
        ys.append(b["index"])
    assert rel(res["loss"], outs[rank]["loss"]) < 1e-9              # per-rank loss = oracle on that rank's shard
    gmean = {k: sum(o["grads"][k] for o in outs) / world for k in outs[0]["grads"]}
    O.sgd_step(P, gmean, mom, first=True)
    sd = model.state_dict()
    for k in layout:
        if O.is_param(k):
            assert rel(sd[k], P[k]) < 1e-7, k
    all_f, all_y = torch.cat(fs), torch.cat(ys)
    for i in range(3):
        O.bank_update(banks[i], all_f[:, 128 * i:128 * (i + 1)], all_y)
        assert rel(getattr(mem, "memory_%d" % (i + 1)), banks[i]) < 1e-9
    # ---- replicas stay bit-identical
    flat = torch.cat([model.store.p, mem.memory_1.reshape(-1), mem.memory_3.reshape(-1)])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], g) for g in gathered)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, "ok"))


@pytest.mark.timeout(900)
def test_two_rank_step_contract():
    world, port = 2, 29000 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(850)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(0, "ok"), (1, "ok")]
