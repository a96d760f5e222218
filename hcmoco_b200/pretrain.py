"""One data-parallel pre-train step on top of the engine: random draws, CUDA-graph replay of
forward+losses+backward, embedding all-gather -> memory-bank update, gradient all-reduce, fused SGD.

Follows learning/contrast_trainer.py:532-640 (`_train_mem_skeleton3d`) and :894-1039
(`_train_bank_joints_pri3d_cmc3`); collectives as in SURVEY.md §8(e): NCCL all-reduce of the flat
gradient buffer (one message instead of DDP's buckets), all-gather of [B,384] embeddings + [B] indices
so that every rank applies the identical bank update (contrast_trainer.py:578-579, mem_bank.py:195-199).
"""
import math

import torch

from . import layout as L
from .engine import Engine, Plan


def init_parameters(store, seed=0):
    """The reference's initialisers: HRNet convs N(0, 0.001^2), BN gamma 1 / beta 0
    (official_hrnet.py:456-463); nn.Linear / nn.Conv2d defaults for the heads and the 1x1 projections;
    SemGraphConv: xavier_uniform(gain 1.414) W, e = 1, bias U(+-1/sqrt(out)) (sem_graph_conv.py:19-30)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in store.keys.items():
        if L.is_buffer(k):
            continue
        if k.endswith(".W"):
            bound = 1.414 * math.sqrt(6.0 / (shp[1] + shp[2]))
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif k.endswith(".e"):
            t = torch.ones(shp)
        elif len(shp) == 4 and "_linear" in k:
            t = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(shp[1])
        elif len(shp) == 4:
            t = 0.001 * torch.randn(shp, generator=g)
        elif len(shp) == 2:
            t = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(shp[1])
        elif k.endswith(".weight"):                    # BatchNorm gamma
            t = torch.ones(shp)
        elif ".gconv" in k or k.startswith("head") or "_linear" in k:      # linear / gconv / projection biases
            fan = {"head1": store.keys.get("head1.0.weight", (0, 1))[1], "head2": store.keys.get("head2.0.weight", (0, 1))[1],
                   "head3": 128}.get(k.split(".")[0], 128)
            if "_linear" in k:
                fan = store.keys[k.replace("bias", "weight")][1]
            t = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan)
        else:                                          # BatchNorm beta
            t = torch.zeros(shp)
        sd[k] = t
    for k in store.keys:
        if k.endswith("running_mean"):
            sd[k] = torch.zeros(store.keys[k])
        elif k.endswith("running_var"):
            sd[k] = torch.ones(store.keys[k])
        elif k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.int64)
    store.load_state_dict(sd)


class InputStager:
    """Double-buffered host->device staging of a step's inputs on a copy stream, so that the copies of step i+1 (the [B,6,R,R]
    RGB-D tensor alone is 100 MB at B=64, ~2 ms over PCIe) overlap the kernels of step i; the step itself then starts with
    device-to-device copies into the static buffers the CUDA graph reads.  `like` = one device tensor or a list of them."""

    def __init__(self, like):
        self.single = not isinstance(like, (list, tuple))
        like = [like] if self.single else list(like)
        self.buf = [[torch.empty_like(t) for t in like] for _ in range(2)]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]        # slot consumed (its D2D copies are enqueued)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]       # H2D into the slot done
        self.stream = torch.cuda.Stream()
        self.slot = 0
        for ev in self.free:
            ev.record()

    def stage(self, host):
        host = [host] if self.single else host
        k = self.slot
        self.slot ^= 1
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.free[k])
            for b, t in zip(self.buf[k], host):
                b.copy_(t, non_blocking=True)
            self.ready[k].record(self.stream)
        return k

    def consume(self, k, dst):
        dst = [dst] if self.single else dst
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready[k])
        for d, b in zip(dst, self.buf[k]):
            d.copy_(b, non_blocking=True)
        self.free[k].record(cur)


class PretrainStep:
    def __init__(self, K, width=18, stage=1, skeleton="mpii", B=64, R=256, n_data=165894, nce_k=16384, nce_t=0.07,
                 nce_m=0.5, temperature=0.07, num_samples=400, world_size=1, rank=0, use_graph=True, seed=0,
                 lr=0.03, momentum=0.9, weight_decay=1e-4):
        self.K, self.world, self.rank = K, world_size, rank
        self.lr, self.momentum, self.wd = lr, momentum, weight_decay
        self.eng = Engine(K, width, stage, skeleton, B, R, n_data, nce_k, nce_t, nce_m, temperature, num_samples,
                          world_size=world_size)
        init_parameters(self.eng.store, seed)
        self.eng.init_banks(seed=seed)               # identical on every rank (all three banks)
        self.eng.build()
        self.stager, self._staged = None, {}
        self.injected = None
        self.gen = torch.Generator(device="cuda").manual_seed(1000 + rank)
        self.use_graph = use_graph
        if use_graph:
            self.eng.capture()
        if world_size > 1:
            B = self.eng.B
            self.all_f = torch.empty(world_size * B, 384, device="cuda")
            self.all_y = torch.empty(world_size * B, dtype=torch.int64, device="cuda")

    @property
    def launches_per_step(self):
        return self.eng.launches_per_step

    def draw(self, batch):
        """Negative indices (memory/alias_multinomial.py:49-65 with uniform probabilities = uniform draw,
        idx[:,0] = own row, mem_bank.py:176-177) and the dense pixel samples (contrast_trainer.py:674-685)."""
        e = self.eng
        if self.injected is not None:           # tests: (nce_idx [B,K+1], dense_idx [B,S] or None)
            e.nce_idx.copy_(self.injected[0])
            if e.stage == 2:
                e.dense_idx.copy_(self.injected[1])
            return
        idx = torch.randint(0, e.n_data, (e.B, e.K1), device="cuda", generator=self.gen)
        idx[:, 0] = e.index
        e.nce_idx.copy_(idx)
        if e.stage == 2:
            step = e.R // e.h
            m = e.depth_mask[:, ::step, ::step][:, :e.h, :e.h].reshape(e.B, -1)
            has = m.sum(1, keepdim=True) > 0
            w = torch.where(has, m, torch.ones_like(m))          # rows of dropped samples are ignored downstream
            e.dense_idx.copy_(torch.multinomial(w, e.S, replacement=True, generator=self.gen))

    def _targets(self):
        """(tuple position in the reference's batch layout, static device buffer) of every input the step reads: the depth mask
        (16.8 MB at B=64) only feeds the dense objective, so the first stage never moves it."""
        e = self.eng
        t = [(0, e.x), (1, e.index), (2, e.skel), (4, e.joints_yx), (5, e.joints_vis), (6, e.use_depth)]
        if e.stage == 2:
            t.append((7, e.depth_mask))
        return t

    def h2d_bytes(self, batch):
        return sum(batch[i].numel() * batch[i].element_size() for i, _ in self._targets())

    def prefetch(self, batch):
        """Start the host->device copies of the NEXT step's inputs on the copy stream while the current step computes;
        `run(batch)` recognises a prefetched batch by identity.  Only meaningful for pinned host batches."""
        x = batch[0]
        if x.is_cuda:
            return
        tg = self._targets()
        if self.stager is None:
            self.stager = InputStager([d for _, d in tg])
        self._staged[id(x)] = (x, self.stager.stage([batch[i] for i, _ in tg]))

    def run(self, batch, next_batch=None):
        e = self.eng
        data = batch
        tg = self._targets()
        hit = self._staged.pop(id(data[0]), None)
        if hit is not None and hit[0] is data[0]:
            self.stager.consume(hit[1], [d for _, d in tg])
        else:
            for i, d in tg:
                d.copy_(data[i], non_blocking=True)
        if next_batch is not None:
            self.prefetch(next_batch)
        self.draw(batch)
        if self.use_graph:
            e.graph.replay()
        else:
            e.forward()
            e.backward()
        if self.world > 1:
            import torch.distributed as dist
            # Two collectives of one communicator must never be in flight on different streams.  The first version issued the async
            # all-reduce first and the synchronous all-gathers after it: the all-gathers could run on the compute stream while the
            # all-reduce was still executing on the process group's stream — measured on 8 GPUs, the replicas' parameters drifted
            # apart (max |dp| 0.03 after 24 steps; 2 and 4 GPUs were unaffected) although every call "completed" on every rank
            # (bench.py `replicas_identical`, profiles/r02_bench_8gpu.json).
            # All three are issued synchronously (one at a time whatever stream the process group uses for them); what the async
            # variant overlapped with the all-reduce was three tiny bank-update kernels.
            dist.all_gather_into_tensor(self.all_f, e.f)
            dist.all_gather_into_tensor(self.all_y, e.index)
            dist.all_reduce(e.store.g)
            e.update_banks(self.all_f, self.all_y)
            e.sgd(self.lr, self.momentum, self.wd, 1.0 / self.world)
        else:
            e.update_banks()
            e.sgd(self.lr, self.momentum, self.wd)

    def results(self):
        return self.eng.results()

    def results_async(self):
        return self.eng.results_async()

    def profile_families(self, batch):
        """Device time of every C-ABI launch of one real step (CUDA events on the launching stream around
        each call, no graph), summed per entry point."""
        e = self.eng
        saved = e.store.save_buffers()
        g_saved = e.store.g.clone()
        self.run(batch)                       # make inputs current
        torch.cuda.synchronize()
        ev, names, shapes = [], [], []

        def go(prog):
            for fn, args, _tag in prog:
                if isinstance(fn, str):          # FORK / JOIN markers
                    continue
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn(*args)
                b.record()
                ev.append((a, b))
                n = getattr(fn, "__name__", "fn")
                n = {"fwd_logits": "nce_logits", "bwd": "nce_bwd"}.get(n, n[4:] if n.startswith("hcm_") else n)
                names.append(n)
                o = {"tc_conv": 4, "conv2d_fwd": 4, "conv2d_dgrad": 3, "conv2d_wgrad": 3, "tc_wgrad": 4, "tc_dgrad_s2": 3}.get(n)
                shapes.append(None if o is None else "%s %dx%d %d->%d k%d%s" % (
                    n, args[o + 1], args[o + 2], args[o + 3], args[o + 4], args[o + 5],
                    "" if n == "tc_dgrad_s2" else " s%d" % args[o + 6]))

        go(e.plan.fwd)
        e.K.zero(e.store.g, e.store.n * 4)
        go(e.plan.bwd)
        torch.cuda.synchronize()
        fam = {}
        self.detail = {}
        for (a, b), n, sh in zip(ev, names, shapes):
            d = fam.setdefault(n, {"ms": 0.0, "calls": 0})
            t = a.elapsed_time(b)
            d["ms"] += t
            d["calls"] += 1
            if sh is not None:
                dd = self.detail.setdefault(sh, {"ms": 0.0, "calls": 0})
                dd["ms"] += t
                dd["calls"] += 1
        e.store.restore_buffers(saved)
        e.store.g.copy_(g_saved)
        return fam
