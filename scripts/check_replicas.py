"""Diagnose replica consistency under N ranks: checksums of parameters at construction, of the all-reduced gradient, and of the
parameters / momentum after each of a few steps.  torchrun --nproc-per-node N scripts/check_replicas.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world)
from hcmoco_b200.kernels import CudaKernels  # noqa: E402
from hcmoco_b200.pretrain import PretrainStep  # noqa: E402
from hcmoco_b200.synthetic import make_batch  # noqa: E402


def same(t, what):
    v = t.reshape(-1).view(torch.int32).to(torch.int64)
    mine = torch.stack([v.sum(), (v * (torch.arange(v.numel(), device=v.device) % 8191 + 1)).sum()])
    allv = torch.empty(world, 2, dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(allv, mine)
    ok = bool((allv == allv[0:1]).all())
    if rank == 0:
        print("%-40s identical=%s" % (what, ok), flush=True)
    return ok


K = CudaKernels()
B = int(os.environ.get("BATCH", 16))
R = int(os.environ.get("RES", 128))
NDATA = int(os.environ.get("NDATA", 20000))
NCEK = int(os.environ.get("NCEK", 1024))
SYNC = int(os.environ.get("SYNC", 1))
step = PretrainStep(K, width=18, stage=1, skeleton="mpii", B=B, R=R, n_data=NDATA, nce_k=NCEK, world_size=world, rank=rank,
                    use_graph=True, seed=0)
e = step.eng
same(e.store.p, "params at construction")
# a bare all-reduce of rank-dependent data
t = torch.randn(1 << 22, device="cuda", generator=torch.Generator(device="cuda").manual_seed(rank))
dist.all_reduce(t)
same(t, "all_reduce(randn) result")
t2 = torch.randn(19_600_000, device="cuda", generator=torch.Generator(device="cuda").manual_seed(100 + rank))
h = dist.all_reduce(t2, async_op=True)
h.wait()
same(t2, "async all_reduce(78 MB) result")
batches = [[x.cuda() for x in make_batch(B, R, 16, NDATA, seed=1234 + rank + 17 * i)] for i in range(2)]


def maxdiff(t, what):
    ref = t.clone()
    dist.broadcast(ref, 0)
    d = (t - ref).abs().max()
    allv = torch.empty(world, device="cuda")
    dist.all_gather_into_tensor(allv, d.reshape(1))
    nf = (~torch.isfinite(t)).sum()
    if rank == 0:
        print("   %s: max |x - x(rank 0)| per rank %s   (|x| max %.3e, non-finite on rank 0: %d)" % (
            what, ["%.2e" % v for v in allv.tolist()], float(t.abs().max()), int(nf)), flush=True)


for s in range(int(os.environ.get("NSTEPS", 4))):
    step.run(batches[s % 2])
    if SYNC:
        torch.cuda.synchronize()
        ok = same(e.store.g, "step %d: store.g after the step" % s)
        ok &= same(e.store.p, "step %d: params" % s)
        if not ok:
            maxdiff(e.store.g, "g")
            maxdiff(e.store.p, "p")
torch.cuda.synchronize()
same(e.store.g, "end: store.g")
if not same(e.store.p, "end: params"):
    maxdiff(e.store.g, "g")
    maxdiff(e.store.p, "p")
same(e.store.m, "end: momentum")
same(e.banks[0], "end: memory_1")
dist.destroy_process_group()
