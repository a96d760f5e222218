"""Host engine of the HCMoCo pre-train step: builds, once per (batch, resolution, model) shape, a static
launch program over preallocated device buffers — forward, losses, backward, bank update, SGD — whose
every arithmetic op is a C-ABI call into libhcmoco_sm100.so (include/hcmoco.h).  The program only
enqueues kernels on the current stream, so it replays under a CUDA graph.

What it restates (paths relative to pycontrast/ in the reference):
  model      networks/build_backbone.py:186-303 (CMC3HRNetSGCNSingleHead), official_hrnet.py:32-454,
             SGCN/sem_gcn.py:8-95, SGCN/sem_graph_conv.py:34-48
  NCE        memory/mem_bank.py:157-205, learning/contrast_trainer.py:212-253
  dense      learning/contrast_trainer.py:642-723     sparse  :744-828     SCL  :830-892
  step       learning/contrast_trainer.py:532-640 (first stage), :894-1039 (second stage)
  optimiser  main_contrast.py:78-81 (SGD momentum 0.9, wd 1e-4)

Layout: activations channels-last fp32 [B,H,W,C]; conv weights in checkpoint layout (OIHW) inside one
flat parameter buffer (+ flat gradient and momentum buffers of the same shape, so the optimiser and
the gradient all-reduce are single launches).  A BatchNorm is never applied as a separate pass when
its consumer is a convolution: the consumer applies scale/shift(+ReLU) while loading (`Act.scale`).
"""
from collections import OrderedDict

import torch

from . import layout as L

BN2D_MOMENTUM = 0.01    # official_hrnet.py:22-23
BN1D_MOMENTUM = 0.1     # nn.BatchNorm1d default (sem_gcn.py:13)
BN_EPS = 1e-5


class Act:
    """Activation value = relu?(data * scale + shift); scale None -> value = data (materialised)."""
    __slots__ = ("data", "B", "H", "W", "C", "scale", "shift", "relu", "grad", "grad_ready", "needs_grad", "spec",
                 "consumers")

    def __init__(self, data, B, H, W, C, scale=None, shift=None, relu=False, needs_grad=True):
        self.data, self.B, self.H, self.W, self.C = data, B, H, W, C
        self.scale, self.shift, self.relu = scale, shift, relu
        self.grad, self.grad_ready, self.needs_grad = None, False, needs_grad
        self.spec = None        # lazy acts: how the single consumer hands the gradient over
        self.consumers = 0

    @property
    def P(self):
        return self.B * self.H * self.W

    @property
    def lazy(self):
        return self.scale is not None


class ParamStore:
    """Flat fp32 parameter / gradient / momentum buffers with per-key views in checkpoint layout."""

    def __init__(self, K, keys):
        self.keys = keys
        self.K = K
        off, self.off = 0, {}
        for k, shp in keys.items():
            if L.is_buffer(k):
                continue
            n = 1
            for s in shp:
                n *= s
            self.off[k] = (off, n)
            off += (n + 3) // 4 * 4
        self.n = off
        self.p = K.zeros(off)
        self.g = K.zeros(off)
        self.m = K.zeros(off)
        # BatchNorm buffers: ONE flat fp32 tensor (running_mean / running_var) and ONE int64 tensor (num_batches_tracked) with
        # per-key views — no per-key allocation / fill launches (619 BN layers), and save / restore is two copies
        nf = sum(shp[0] for k, shp in keys.items() if k.endswith(("running_mean", "running_var")))
        nt = sum(1 for k in keys if k.endswith("num_batches_tracked"))
        init = torch.zeros(nf)
        spans, o, t = [], 0, 0
        for k, shp in keys.items():
            if k.endswith("num_batches_tracked"):
                spans.append((k, None, t))
                t += 1
            elif k.endswith(("running_mean", "running_var")):
                if k.endswith("running_var"):
                    init[o:o + shp[0]] = 1.0
                spans.append((k, o, shp[0]))
                o += shp[0]
        self.bflat = init.to(device=K.device, dtype=K.dtype)
        self.nbt = torch.zeros(nt, dtype=torch.int64, device=K.device)
        self.buffers = OrderedDict()
        for k, o, n in spans:
            self.buffers[k] = self.nbt[n] if o is None else self.bflat[o:o + n]

    def save_buffers(self):
        return self.bflat.clone(), self.nbt.clone()

    def restore_buffers(self, saved):
        self.bflat.copy_(saved[0])
        self.nbt.copy_(saved[1])

    def view(self, flat, key):
        o, n = self.off[key]
        return flat[o:o + n]

    def param(self, key):
        return self.view(self.p, key)

    def grad(self, key):
        return self.view(self.g, key)

    # ---- every tensor is stored in exactly the checkpoint layout (OIHW convs, [out,in] linears, [2,Cin,Cout] SemGCN)
    def export(self, flat, key):
        return self.view(flat, key).reshape(self.keys[key]).clone()

    def load(self, flat, key, t):
        v = self.view(flat, key)
        v.copy_(t.to(device=v.device, dtype=v.dtype).reshape(-1))

    def state_dict(self, prefix=""):
        out = OrderedDict()
        for k in self.keys:
            out[prefix + k] = self.buffers[k].clone() if L.is_buffer(k) else self.export(self.p, k)
        return out

    def load_state_dict(self, sd, strict=True):
        missing = []
        for k in self.keys:
            src = sd.get(k, sd.get("module." + k))
            if src is None:
                missing.append(k)
                continue
            if L.is_buffer(k):
                self.buffers[k].copy_(src.to(self.buffers[k].device))
            else:
                if tuple(src.shape) != tuple(self.keys[k]):
                    raise ValueError("shape mismatch for %s: %s vs %s" % (k, tuple(src.shape), self.keys[k]))
                self.load(self.p, k, src)
        if strict and missing:
            raise KeyError("missing keys: %s ..." % missing[:5])
        return missing

    def grads_dict(self):
        return OrderedDict((k, self.export(self.g, k)) for k in self.keys if not L.is_buffer(k))


FORK, JOIN = "fork", "join"      # program markers: the side stream starts from / is merged back into the main stream


class Plan:
    """Recorded launch lists.  `f` appends a forward launch; ops register a backward *builder* that is
    invoked in reverse op order by `finish()`, so that first-writer / accumulate flags of gradient
    buffers are resolved statically.

    Every launch carries a stream tag.  Tag 0 is the stream the program is run on; `fork(parent, children)` lets the
    children's streams start from the parent's current position, `join(parent, children)` makes the parent wait for them
    (events; everything is captured into the same CUDA graph).  Both register the mirrored marker for the backward program.
    Used at two levels: encoder2 (+ the skeleton encoder) beside encoder1, and inside every HR module the branches 1..3
    beside branch 0 — the low-resolution branches have 50-160 tiles for 148 SMs and are latency-bound, so their kernels
    fill the SMs (and the gaps between the kernels) of the high-resolution branch instead of running alone."""

    def __init__(self, K):
        self.K = K
        self.fwd, self.bwd, self._builders = [], [], []
        self.tag = 0
        self.streams = {}
        self.off_path_streams = True
        self._wtags = set()

    def f(self, fn, *args):
        self.fwd.append((fn, args, self.tag))

    def b(self, fn, *args):
        self.bwd.append((fn, args, self.tag))

    def b_off_path(self, fn, *args):
        """A backward launch nothing later in the program depends on (a weight gradient: it only feeds the flat gradient
        buffer): it goes to the side stream `8 + tag`, ordered after what has been recorded on `tag` so far, and is joined at
        the very end of the backward program.  The critical chain of a layer is then BN backward -> data gradient, and
        the weight-gradient kernels fill the SMs and the gaps those leave."""
        if not self.off_path_streams:
            return self.b(fn, *args)
        wtag = 8 + self.tag
        self._wtags.add(wtag)
        self.bwd.append((FORK, (self.tag, (wtag,)), self.tag))
        self.bwd.append((fn, args, wtag))

    def fork(self, parent, children):
        children = tuple(children)
        self.fwd.append((FORK, (parent, children), parent))
        self._builders.append((lambda: self.bwd.append((JOIN, (parent, children), parent)), parent))

    def join(self, parent, children):
        children = tuple(children)
        self.fwd.append((JOIN, (parent, children), parent))
        self._builders.append((lambda: self.bwd.append((FORK, (parent, children), parent)), parent))

    def on_backward(self, builder):
        self._builders.append((builder, self.tag))

    def finish(self, mark=0):
        """Run the builders in reverse registration order; returns len(bwd) after the builders >= `mark` ran."""
        for bld, tag in reversed(self._builders[mark:]):
            self.tag = tag
            bld()
        n = len(self.bwd)
        for bld, tag in reversed(self._builders[:mark]):
            self.tag = tag
            bld()
        self.tag = 0
        self._builders = []
        if self._wtags:
            self.bwd.append((JOIN, (0, tuple(sorted(self._wtags))), 0))
        return n

    def grad(self, act):
        """Gradient buffer of a materialised act + whether the next writer must accumulate."""
        if act.grad is None:
            act.grad = self.K.empty(act.B, act.H, act.W, act.C)
        acc = act.grad_ready
        act.grad_ready = True
        return act.grad, int(acc)

    def zeroed_grad(self, act):
        """Gradient buffer that scatter-style writers (atomics) can add into."""
        if act.grad is None:
            act.grad = self.K.empty(act.B, act.H, act.W, act.C)
        if not act.grad_ready:
            self.b(self.K.zero, act.grad, act.grad.numel() * act.grad.element_size())
            act.grad_ready = True
        return act.grad

    def run(self, prog, two_streams=True):
        """Enqueue a program.  With a CUDA device and two_streams, launches go to the stream of their tag (tag 0 = the
        current stream); otherwise everything runs in program order on the current stream."""
        use_side = two_streams and self.K.device == "cuda" and any(e[0] is FORK for e in prog)
        if not use_side:
            for fn, args, _ in prog:
                if fn is not FORK and fn is not JOIN:
                    fn(*args)
            return
        main = torch.cuda.current_stream()

        def stream_of(tag):
            if tag == 0:
                return main
            if tag not in self.streams:
                self.streams[tag] = torch.cuda.Stream()
            return self.streams[tag]

        for fn, args, tag in prog:
            if fn is FORK:
                parent, children = args
                ev = torch.cuda.Event()
                ev.record(stream_of(parent))
                for c in children:
                    stream_of(c).wait_event(ev)
            elif fn is JOIN:
                parent, children = args
                for c in children:
                    ev = torch.cuda.Event()
                    ev.record(stream_of(c))
                    stream_of(parent).wait_event(ev)
            elif tag == 0:
                fn(*args)
            else:
                with torch.cuda.stream(stream_of(tag)):
                    fn(*args)


class Engine:
    def __init__(self, K, width=18, stage=1, skeleton="mpii", B=2, R=224, n_data=20000, nce_k=16384, nce_t=0.07,
                 nce_m=0.5, temperature=0.07, num_samples=400, feat_dim=128, world_size=1, train=True, use_tc=True,
                 store=None, two_streams=True, fuse_bn_finalize=False, branch_streams=True,
                 wgrad_streams=True):
        assert feat_dim == 128, "the NCE / loss kernels are specialised for feat_dim=128"
        assert R % 32 == 0, "HRNet needs the input side to be a multiple of 32"
        assert B >= 2, "the reference collapses B=1 (mem_bank.py:39 out.squeeze())"
        self.K, self.width, self.stage, self.skeleton = K, width, stage, skeleton
        self.B, self.R, self.h = B, R, R // 4
        self.J, rows, cols = L.graph_edges(skeleton)
        self.nnz = len(rows)
        self.n_data, self.K1, self.T_nce, self.m_nce = n_data, nce_k + 1, nce_t, nce_m
        self.T, self.S = temperature, num_samples
        assert width in (18, 32), ("HRNet-w%d: the BatchNorm / pooling / tensor-core kernels tile at most 256 channels (w18: 144, "
                                   "w32: 256); w48's 384-channel stage-4 branch is not built (no shipped pre-train script uses it)" % width)
        self.world = world_size
        self.two_streams = two_streams   # encoder2 on a side stream (see Plan)
        self.fuse_bn_finalize = fuse_bn_finalize     # see _bn_stats
        self.branch_streams = branch_streams         # HR-module branches 1..3 on their own streams (see Plan)
        self.wgrad_streams = wgrad_streams           # weight gradients off the critical path (Plan.b_off_path)
        self.use_tc = use_tc     # tensor-core path for the stride-1 convs (SIMT fp32 implicit GEMM otherwise)
        self.ch = L.WIDTHS[width]
        self.cm = sum(self.ch)
        # `store`: parameter storage shared with an api.HCMoCoModel (several engines = several batch shapes, one model)
        self.store = store if store is not None else ParamStore(K, L.model_keys(width, stage, skeleton, feat_dim))
        self.edge_rows = torch.tensor(rows, dtype=torch.int32, device=K.device)
        self.edge_cols = torch.tensor(cols, dtype=torch.int32, device=K.device)
        self.banks = None
        self.first_step = True
        self.built = False

    # ------------------------------------------------------------------ memory bank (mem_bank.py:157-170)
    def init_banks(self, banks=None, seed=0):
        if banks is None:
            g = torch.Generator().manual_seed(seed)
            banks = [torch.nn.functional.normalize(torch.randn(self.n_data, 128, generator=g)) for _ in range(3)]
        self.banks = [b.to(device=self.K.device, dtype=self.K.dtype).contiguous().clone() for b in banks]

    # ------------------------------------------------------------------ plan construction
    def build(self):
        K, B, R, J = self.K, self.B, self.R, self.J
        self.plan = p = Plan(K)
        p.off_path_streams = self.wgrad_streams
        self.pack_jobs = []
        p.f(self._run_packs)
        maxc = 4 * max(self.ch[-1], 256)
        # shared scratch (single stream => sequential reuse is safe)
        # scratch per stream tag (launches with the same tag are sequential, so reuse within a tag is safe)
        ntags = 8          # 0 / 1: the encoders' own streams; 2-4 / 5-7: branches 1..3 of their HR modules
        self._part = [K.empty(2 * 4096 * 2 * 256) for _ in range(ntags)]
        self._k = [(K.empty(maxc), K.empty(maxc), K.empty(maxc)) for _ in range(ntags)]
        self._cnt = [K.zeros(4, dtype=torch.int32) for _ in range(ntags)]  # last-CTA tickets of the fused BN statistics kernels
        # step inputs (static buffers; the caller copies each batch in)
        self.x = K.zeros(B, 6, R, R)
        self.skel = K.zeros(B, J, 2)
        self.index = K.zeros(B, dtype=torch.int64)
        self.joints_yx = K.zeros(B, J, 2)
        self.joints_vis = K.zeros(B, J, dtype=torch.int32)
        self.use_depth = K.zeros(B, dtype=torch.int64)
        self.depth_mask = K.zeros(B, R, R)
        self.nce_idx = K.zeros(B, self.K1, dtype=torch.int64)
        self.dense_idx = K.zeros(B, self.S, dtype=torch.int64)
        # model
        xs = []
        for m in range(2):
            if self.use_tc and K.tc_conv_supported(B, R, R, 4, 64, 3, 2) and K.tc_wgrad_supported(B, R, R, 4, 64, 3, 2):
                # the 3-channel planes are stored with a zero 4th channel (16-byte rows, even channel count): the stem convolution
                # and its weight gradient then run on the tensor-core kernels instead of the SIMT ones (B=64, 256^2:
                # 489 + 805 us per encoder before); the weight keeps its [64,3,3,3] checkpoint layout (pack / wgrad with ld = 3)
                xin = K.empty(B, R, R, 4)
                p.f(K.nchw_to_nhwc_pad, self.x, xin, B, 6, R * R, 3 * m, 3, 4)
                xs.append(Act(xin, B, R, R, 4, needs_grad=False))
            else:
                xin = K.empty(B, R, R, 3)
                p.f(K.nchw_to_nhwc, self.x, xin, B, 6, R * R, 3 * m, 3)
                xs.append(Act(xin, B, R, R, 3, needs_grad=False))
        p.fork(0, [1])
        self.feat1 = self._hrnet("encoder1.", xs[0])
        p.tag = 1
        # the skeleton encoder's launches are tiny (B*J rows): on the side stream they hide under encoder1's kernels
        self.feat3 = self._sgcn("encoder3.", self.skel)
        self.feat2 = self._hrnet("encoder2.", xs[1])
        p.tag = 0
        p.join(0, [1])
        self.f = K.empty(B, 384)
        self.df = K.zeros(B, 384)
        self._head("head1.0", self._pool(self.feat1), self.cm, 0)
        self._head("head2.0", self._pool(self.feat2), self.cm, 1)
        self._head("head3.0", self._joint_mean(self.feat3), 128, 2)
        if self.stage == 2:
            self.lm1 = self._projection("encoder1_linear", self.feat1)
            self.lm2 = self._projection("encoder2_linear", self.feat2)
        self.n_model_fwd = len(p.fwd)
        # losses (the builders registered here run first in the backward)
        self.losses = K.zeros(16)     # 0-5 nce, 6-7 dense, 8-9 joint, 10 scl
        self.accs = K.zeros(16)       # 0-5 nce, 6-7 dense, 8-9 joint
        n_model_builders = len(p._builders)
        self._nce()
        if self.stage == 2:
            self._stage2_losses()
        self.n_loss_bwd = p.finish(n_model_builders)      # bwd[:n_loss_bwd] = loss kernels, the rest = model backward
        self._finish_packs()
        self.built = True
        return self

    @property
    def part(self):
        return self._part[self.plan.tag]

    @property
    def cnt(self):
        return self._cnt[self.plan.tag]

    @property
    def k1(self):
        return self._k[self.plan.tag][0]

    @property
    def k2(self):
        return self._k[self.plan.tag][1]

    @property
    def k3(self):
        return self._k[self.plan.tag][2]

    # ---- weight packing for the tensor-core kernels: every job of the step goes into ONE launch at the head of the forward
    def _pack_job(self, w, ldw, wp, geo, cin, cout, ks, mode):
        self.pack_jobs.append((w, ldw, wp, geo, cin, cout, ks, mode))

    def _run_packs(self):
        K = self.K
        if not self.pack_jobs:
            return
        if hasattr(K, "tc_pack_batch") and self.pack_table is not None:
            K.tc_pack_batch(self.pack_table, len(self.pack_jobs), self.pack_steps)
            return
        for w, ldw, wp, (B, H, W), cin, cout, ks, mode in self.pack_jobs:
            if mode == 2:
                if (ldw & 15) == 0:          # one call packs every parity group
                    K.tc_dgrad_s2_pack(w, wp, B, H, W, cin, cout)
            else:
                K.tc_conv_pack(w, ldw, wp, B, H, W, cin, cout, ks, mode)

    def _finish_packs(self):
        self.pack_table, self.pack_steps = None, 0
        if not self.pack_jobs or not hasattr(self.K, "tc_pack_batch"):
            return
        rows, first = [], 0
        for w, ldw, wp, geo, cin, cout, ks, mode in self.pack_jobs:
            rows.append([w.data_ptr(), wp.data_ptr(), cin, cout, ks, mode, ldw, first])
            kch = (cout if mode == 2 else cin)
            first += (4 if mode == 2 else ks * ks) * ((kch + 15) // 16)
        self.pack_steps = first
        self.pack_table = torch.tensor(rows, dtype=torch.int64).to(self.K.device)

    # ---- train-mode BN statistics (forward) / gradient sums (backward) + their per-channel finalize.
    # fuse_bn_finalize=True uses the one-launch forms (last CTA reduces the partial rows).  Measured on B200 (B=64 step):
    # 102.9 ms fused vs 99.4 ms with the separate C-CTA finalize kernels — the serial tail of one CTA costs more than the
    # ~5 us launch it saves — so the two-launch form is the default.
    def _bn_stats(self, y, P, C, bk, momentum, scale, shift, mean, invstd):
        K, p, st, bf = self.K, self.plan, self.store, self.store.buffers
        args = (st.param(bk + ".weight"), st.param(bk + ".bias"), bf[bk + ".running_mean"], bf[bk + ".running_var"],
                bf[bk + ".num_batches_tracked"], momentum, BN_EPS, scale, shift, mean, invstd)
        assert K.colstat_rows(P, C) * 2 * C <= self.part.numel()
        if self.fuse_bn_finalize:
            p.f(K.bn_stats_finalize, y, P, C, self.part, self.cnt, *args)
        else:
            p.f(K.bn_stats, y, P, C, self.part)
            p.f(K.bn_finalize, self.part, K.colstat_rows(P, C), C, P, *args)

    def _bn_bwd_reduce(self, dz, mask, msc, msh, y, mean, invstd, P, C, bk):
        K, p, st = self.K, self.plan, self.store
        args = (st.param(bk + ".weight"), st.grad(bk + ".weight"), st.grad(bk + ".bias"), self.k1, self.k2, self.k3)
        if self.fuse_bn_finalize:
            p.b(K.bn_bwd_reduce_finalize, dz, mask, msc, msh, y, mean, invstd, P, C, self.part, self.cnt, *args)
        else:
            p.b(K.bn_bwd_reduce, dz, mask, msc, msh, y, mean, invstd, P, C, self.part)
            p.b(K.bn_bwd_finalize, self.part, K.colstat_rows(P, C), C, P, args[0], mean, invstd, *args[1:])

    # ---- conv + train-mode BN; output is lazy (raw conv output + per-channel affine)
    def _conv_bn(self, x, ck, bk, stride, relu):
        K, p, st = self.K, self.plan, self.store
        cout, cin_w, ks, _ = st.keys[ck + ".weight"]
        cin = x.C                              # channels of the stored input; > cin_w only for the zero-padded stem input (3 -> 4)
        assert cin == cin_w or (cin == 4 and cin_w == 3 and not x.needs_grad), (ck, cin_w, x.C)
        ldw = cin_w if cin != cin_w else 0
        B, H, W = x.B, x.H, x.W
        pad = (ks - 1) // 2
        Ho, Wo = (H + 2 * pad - ks) // stride + 1, (W + 2 * pad - ks) // stride + 1
        P = B * Ho * Wo
        y = K.empty(B, Ho, Wo, cout)
        w = st.param(ck + ".weight")
        tc = self.use_tc and bool(K.tc_conv_supported(B, H, W, cin, cout, ks, stride))
        scale, shift, mean, invstd = K.empty(cout), K.empty(cout), K.empty(cout), K.empty(cout)
        bf = st.buffers
        x.consumers += 1
        if tc:
            nb = K.tc_conv_wpack_bytes(B, H, W, cin, cout, ks)
            wp_f = K.empty((nb + 3) // 4)
            rows = K.colstat_rows(P, cout)
            self._pack_job(w, ldw, wp_f, (B, H, W), cin, cout, ks, 4 * K.tc_conv_rowcat_supported(cout, ks, stride))
            assert rows * 2 * cout <= self.part.numel()
            p.f(K.tc_conv, x.data, wp_f, None, y, B, H, W, cin, cout, ks, stride, x.scale, x.shift, int(x.relu), 0)
            self._bn_stats(y, P, cout, bk, BN2D_MOMENTUM, scale, shift, mean, invstd)
        else:
            rows = K.conv2d_stat_rows(B, H, W, cin, cout, ks, stride)
            assert rows * 2 * cout <= self.part.numel()
            p.f(K.conv2d_fwd, x.data, w, None, y, B, H, W, cin, cout, ks, stride, x.scale, x.shift, int(x.relu), self.part)
            p.f(K.bn_finalize, self.part, rows, cout, P, st.param(bk + ".weight"), st.param(bk + ".bias"),
                bf[bk + ".running_mean"], bf[bk + ".running_var"], bf[bk + ".num_batches_tracked"], BN2D_MOMENTUM, BN_EPS,
                scale, shift, mean, invstd)
        out = Act(y, B, Ho, Wo, cout, scale, shift, relu)

        def backward():
            # gradient wrt the lazy value arrives either in out.grad (consumer was a conv: in place) or via out.spec
            if out.spec is None:
                assert out.grad is not None, "lazy act without consumer: " + ck
                spec = dict(dz=out.grad, mask=None, recompute=relu, dy=out.grad, g_out=None, g_acc=0)
            else:
                spec = out.spec
            msc, msh = (scale, shift) if spec["recompute"] else (None, None)
            self._bn_bwd_reduce(spec["dz"], spec["mask"], msc, msh, y, mean, invstd, P, cout, bk)
            p.b(K.bn_bwd_apply, spec["dz"], spec["mask"], msc, msh, y, self.k1, self.k2, self.k3, spec["dy"],
                spec["g_out"], spec["g_acc"], P, cout)
            dy = spec["dy"]
            if self.use_tc and K.tc_wgrad_supported(B, H, W, cin, cout, ks, stride):
                p.b_off_path(K.tc_wgrad, x.data, dy, st.grad(ck + ".weight"), ldw, B, H, W, cin, cout, ks, stride, x.scale, x.shift,
                             int(x.relu))
            else:
                p.b_off_path(K.conv2d_wgrad, x.data, dy, st.grad(ck + ".weight"), B, H, W, cin, cout, ks, stride, x.scale,
                             x.shift, int(x.relu))
            if x.needs_grad:
                if x.lazy:
                    assert x.consumers == 1 and x.grad is None
                    x.grad = K.empty(B, H, W, cin)
                    gx, acc = x.grad, 0
                else:
                    gx, acc = p.grad(x)
                if tc and stride == 1 and K.tc_conv_supported(B, H, W, cout, cin, ks, 1):
                    # data gradient = the same tensor-core conv run on dy with transposed + flipped weights
                    nbt = K.tc_conv_wpack_bytes(B, H, W, cout, cin, ks)
                    wp_t = K.empty((nbt + 3) // 4)
                    self._pack_job(w, 0, wp_t, (B, H, W), cout, cin, ks, 1 + 4 * K.tc_conv_rowcat_supported(cin, ks, 1))
                    p.b(K.tc_conv, dy, wp_t, None, gx, B, H, W, cout, cin, ks, 1, None, None, 0, acc)
                elif tc and stride == 2 and ks == 3 and K.tc_dgrad_s2_supported(B, H, W, cin, cout):
                    nbt2 = K.tc_dgrad_s2_wpack_bytes(B, H, W, cin, cout)
                    wp_t = K.empty((nbt2 + 3) // 4)
                    nqs = K.tc_dgrad_s2_nqs(B, H, W, cin, cout)
                    per = nbt2 // (4 // nqs)
                    for gi in range(4 // nqs):
                        self._pack_job(w, gi * nqs + 16 * nqs, wp_t[gi * per // 4:], (B, H, W), cin, cout, 3, 2)
                    p.b(K.tc_dgrad_s2, dy, wp_t, gx, B, H, W, cin, cout, acc)
                else:
                    p.b(K.conv2d_dgrad, dy, w, gx, B, H, W, cin, cout, ks, stride, acc)

        p.on_backward(backward)
        return out

    # ---- out = act(bn(y) [+ residual]) as a materialised tensor
    def _materialize(self, a, res=None, relu=None):
        K, p = self.K, self.plan
        relu = a.relu if relu is None else relu
        out = Act(K.empty(a.B, a.H, a.W, a.C), a.B, a.H, a.W, a.C)
        a.consumers += 1
        if res is not None:
            res.consumers += 1
        p.f(K.bn_apply, a.data, a.scale, a.shift, None if res is None else res.data, None if res is None else res.scale,
            None if res is None else res.shift, int(relu), out.data, a.P, a.C)

        def backward():
            assert out.grad is not None, "materialised act without gradient"
            mask = out.data if relu else None
            if res is not None and res.lazy:      # projection shortcut (Bottleneck downsample): needs its own dy
                res.grad = K.empty(a.B, a.H, a.W, a.C)
                res.spec = dict(dz=out.grad, mask=mask, recompute=False, dy=res.grad, g_out=None, g_acc=0)
                # the shortcut's producer runs *after* the main branch in the backward (it was built earlier),
                # so the main branch must not overwrite out.grad in place
                a.grad = K.empty(a.B, a.H, a.W, a.C)
                a.spec = dict(dz=out.grad, mask=mask, recompute=False, dy=a.grad, g_out=None, g_acc=0)
                return
            g_out, g_acc = (None, 0)
            if res is not None and res.needs_grad:
                g_out, g_acc = p.grad(res)
            a.spec = dict(dz=out.grad, mask=mask, recompute=False, dy=out.grad, g_out=g_out, g_acc=g_acc)

        p.on_backward(backward)
        return out

    def _basic_block(self, x, q):          # official_hrnet.py:32-61
        y = self._conv_bn(x, q + "conv1", q + "bn1", 1, True)
        y = self._conv_bn(y, q + "conv2", q + "bn2", 1, False)
        return self._materialize(y, res=x, relu=True)

    def _bottleneck(self, x, q, down):     # official_hrnet.py:64-102
        r = self._conv_bn(x, q + "downsample.0", q + "downsample.1", 1, False) if down else x
        y = self._conv_bn(x, q + "conv1", q + "bn1", 1, True)
        y = self._conv_bn(y, q + "conv2", q + "bn2", 1, True)
        y = self._conv_bn(y, q + "conv3", q + "bn3", 1, False)
        return self._materialize(y, res=r, relu=True)

    # ---- HR-module fuse: out_i = relu(sum_j T_ij(x_j))   (official_hrnet.py:176-220, 232-247)
    def _fuse(self, xs, mp, i):
        K, p = self.K, self.plan
        n = len(xs)
        xi = xs[i]
        terms = []            # (act, log2 factor, kind)
        for j in range(n):
            q = "%sfuse_layers.%d.%d." % (mp, i, j)
            if j == i:
                terms.append((xs[j], 0, "id"))
            elif j > i:
                terms.append((self._conv_bn(xs[j], q + "0", q + "1", 1, False), j - i, "up"))
            else:
                t = xs[j]
                for hop in range(i - j):
                    t = self._conv_bn(t, "%s%d.0" % (q, hop), "%s%d.1" % (q, hop), 2, hop != i - j - 1)
                terms.append((t, 0, "down"))
        out = Act(K.empty(xi.B, xi.H, xi.W, xi.C), xi.B, xi.H, xi.W, xi.C)
        for t, _, _ in terms:
            t.consumers += 1
        log2f = torch.tensor([t[1] for t in terms], dtype=torch.int32)
        p.f(K.fuse_sum, n, [t[0].data for t in terms], [t[0].scale for t in terms], [t[0].shift for t in terms], log2f,
            None, 1, out.data, xi.B, xi.H, xi.W, xi.C)

        def backward():
            assert out.grad is not None
            G = out.grad
            total = G.numel()
            p.b(K.relu_bwd, G, out.data, G, 0, total)          # G <- G * [out > 0], in place
            for t, k, kind in terms:
                if kind == "id":
                    gx, acc = p.grad(t)
                    if acc:
                        p.b(K.axpy, gx, G, 1.0, total)
                    else:
                        p.b(K.relu_bwd, G, out.data, gx, 0, total)   # masked copy (idempotent)
                elif kind == "down":
                    t.grad = K.empty(t.B, t.H, t.W, t.C)
                    t.spec = dict(dz=G, mask=None, recompute=False, dy=t.grad, g_out=None, g_acc=0)
                else:
                    t.grad = K.empty(t.B, t.H, t.W, t.C)
                    p.b(K.upsample_adjoint, G, t.grad, 0, xi.B, xi.H, xi.W, xi.C, k)
                    t.spec = dict(dz=t.grad, mask=None, recompute=False, dy=t.grad, g_out=None, g_acc=0)

        p.on_backward(backward)
        return out

    def _hr_module(self, xs, mp):
        p = self.plan
        xs = list(xs)
        base = p.tag                                           # 0 (encoder1) or 1 (encoder2)
        children = [2 + 3 * base + (i - 1) for i in range(1, len(xs))] if self.branch_streams else []
        if children:
            p.fork(base, children)                             # the branches are independent until the fuse layers
        for i in range(len(xs)):
            p.tag = children[i - 1] if (children and i > 0) else base
            for blk in range(4):
                xs[i] = self._basic_block(xs[i], "%sbranches.%d.%d." % (mp, i, blk))
        p.tag = base
        if children:
            p.join(base, children)
        return [self._fuse(xs, mp, i) for i in range(len(xs))]

    def _hrnet(self, pre, x):              # official_hrnet.py:411-454
        st = self.store
        y = self._conv_bn(x, pre + "conv1", pre + "bn1", 2, True)
        y = self._conv_bn(y, pre + "conv2", pre + "bn2", 2, True)
        x = self._materialize(y)
        for blk in range(4):
            x = self._bottleneck(x, "%slayer1.%d." % (pre, blk), blk == 0)
        ys = [x]
        for s, (nmod, nbr) in enumerate(L.STAGES):
            t = "%stransition%d." % (pre, s + 1)
            xs = []
            for i in range(nbr):
                if i < len(ys):
                    if ("%s%d.0.weight" % (t, i)) in st.keys:
                        xs.append(self._materialize(self._conv_bn(ys[i], "%s%d.0" % (t, i), "%s%d.1" % (t, i), 1, True)))
                    else:
                        xs.append(ys[i])
                else:
                    a = ys[-1]
                    for j in range(i + 1 - len(ys)):
                        a = self._conv_bn(a, "%s%d.%d.0" % (t, i, j), "%s%d.%d.1" % (t, i, j), 2, True)
                    xs.append(self._materialize(a))
            for m in range(nmod):
                xs = self._hr_module(xs, "%sstage%d.%d." % (pre, s + 2, m))
            ys = xs
        return ys

    # ---- SemGCN (sem_gcn.py:60-95; sem_graph_conv.py:34-48): x [B*J, Cin] rows
    def _gconv(self, x, xg_slot, q, cin, cout, bn):
        """x: tensor [B,J,cin]; xg_slot: dict with 'grad' tensor or None (input has no grad), 'ready' flag.
        Returns (value tensor [B,J,cout] post BN/ReLU if bn else raw, slot)."""
        K, p, st, B, J = self.K, self.plan, self.store, self.B, self.J
        A = K.empty(J, J)
        xa = K.empty(B, J, 2 * cin)
        y = K.empty(B, J, cout)
        W, e, bias = st.param(q + ".gconv.W" if bn else q + ".W"), st.param(q + ".gconv.e" if bn else q + ".e"), \
            st.param(q + ".gconv.bias" if bn else q + ".bias")
        gk = (q + ".gconv") if bn else q
        M = B * J
        p.f(K.sgcn_adj, e, self.edge_rows, self.edge_cols, self.nnz, J, A)
        p.f(K.sgcn_aggregate, x, A, B, J, cin, xa)
        p.f(K.gemm, xa, W, bias, y, 1, M, cout, 2 * cin, 2 * cin, 1, cout, 1, cout, 0, 0, 0, 1.0, 0)
        slot = dict(grad=None, ready=False)
        if bn:
            bk = q + ".bn"
            scale, shift, mean, invstd = K.empty(cout), K.empty(cout), K.empty(cout), K.empty(cout)
            out = K.empty(B, J, cout)
            self._bn_stats(y, M, cout, bk, BN1D_MOMENTUM, scale, shift, mean, invstd)
            p.f(K.bn_apply, y, scale, shift, None, None, None, 1, out, M, cout)
        else:
            out = y

        def backward():
            G = slot["grad"]
            assert G is not None and slot["ready"], "gconv output without gradient: " + q
            if bn:
                self._bn_bwd_reduce(G, out, None, None, y, mean, invstd, M, cout, bk)
                p.b(K.bn_bwd_apply, G, out, None, None, y, self.k1, self.k2, self.k3, G, None, 0, M, cout)
            dy = G
            # bias, W ([2*cin, cout] stacked), aggregated input
            p.b(K.colsum_small, dy, M, cout, cout, st.grad(gk + ".bias"), 0)
            # dW[k][n] = sum_m xa[m][k] * dy[m][n]
            p.b(K.gemm, xa, dy, None, st.grad(gk + ".W"), 1, 2 * cin, cout, M, 1, 2 * cin, cout, 1, cout, 0, 0, 0, 1.0, 0)
            dxa = K.empty(B, J, 2 * cin)
            # dxa[m][k] = sum_n dy[m][n] * W[k][n]
            p.b(K.gemm, dy, W, None, dxa, 1, M, 2 * cin, cout, cout, 1, 1, cout, 2 * cin, 0, 0, 0, 1.0, 0)
            dA = K.empty(J, J)
            if xg_slot is not None:
                if xg_slot["grad"] is None:
                    xg_slot["grad"] = K.empty(B, J, cin)
                acc = int(xg_slot["ready"])
                xg_slot["ready"] = True
                p.b(K.sgcn_aggregate_bwd, dxa, x, A, B, J, cin, xg_slot["grad"], acc, dA)
            else:
                p.b(K.sgcn_aggregate_bwd, dxa, x, A, B, J, cin, None, 0, dA)
            p.b(K.sgcn_adj_bwd, A, dA, self.edge_rows, self.edge_cols, self.nnz, J, st.grad(gk + ".e"), 0)

        p.on_backward(backward)
        return out, slot

    def _sgcn(self, pre, s):
        K, p, B, J = self.K, self.plan, self.B, self.J
        x, xs = self._gconv(s, None, pre + "gconv_input.0", 2, 128, True)
        for layer in range(4):
            q = "%sgconv_layers.%d." % (pre, layer)
            y1, s1 = self._gconv(x, xs, q + "gconv1", 128, 128, True)
            y2, s2 = self._gconv(y1, s1, q + "gconv2", 128, 128, True)
            out = K.empty(B, J, 128)
            outs = dict(grad=None, ready=False)
            n = B * J * 128
            p.f(K.bn_apply, x, None, None, y2, None, None, 0, out, B * J, 128)       # out = x + y2

            def backward(xs=xs, s2=s2, outs=outs, n=n):
                G = outs["grad"]
                assert G is not None and outs["ready"]
                s2["grad"], s2["ready"] = G, True                 # d y2 = G (shared, gconv2's BN bwd runs in place last)
                if xs["grad"] is None:
                    xs["grad"] = K.empty(B, J, 128)
                if xs["ready"]:
                    p.b(K.axpy, xs["grad"], G, 1.0, n)
                else:
                    p.b(K.bn_apply, G, None, None, None, None, None, 0, xs["grad"], B * J, 128)   # copy
                    xs["ready"] = True

            p.on_backward(backward)
            x, xs = out, outs
        y, ys = self._gconv(x, xs, pre + "gconv_output", 128, 128, False)
        self.feat3_slot = ys
        return y

    def _slot_grad(self, slot, shape):
        if slot["grad"] is None:
            slot["grad"] = self.K.empty(*shape)
        acc = int(slot["ready"])
        slot["ready"] = True
        return slot["grad"], acc

    # ---- global average pool of the four branches, concatenated (build_backbone.py:267-278)
    def _pool(self, feats):
        K, p, B = self.K, self.plan, self.B
        a = K.empty(B, self.cm)
        da = K.empty(B, self.cm)
        off = 0
        offs = []
        for ft in feats:
            ft.consumers += 1
            p.f(K.avgpool, ft.data, a, B, ft.H * ft.W, ft.C, self.cm, off)
            offs.append(off)
            off += ft.C

        def backward():
            for ft, o in zip(feats, offs):
                g, acc = p.grad(ft)
                p.b(K.avgpool_bwd, da, g, acc, B, ft.H * ft.W, ft.C, self.cm, o)

        p.on_backward(backward)
        return a, da

    def _joint_mean(self, feat3):          # build_backbone.py:279
        K, p, B, J = self.K, self.plan, self.B, self.J
        a, da = K.empty(B, 128), K.empty(B, 128)
        p.f(K.joint_mean, feat3, B, J, 128, a)

        def backward():
            g, acc = self._slot_grad(self.feat3_slot, (B, J, 128))
            p.b(K.joint_mean_bwd, da, B, J, 128, g, acc)

        p.on_backward(backward)
        return a, da

    # ---- Linear + L2 head (build_backbone.py:226-241, networks/util.py:74-81) writing f[:, 128*m : 128*(m+1)]
    def _head(self, key, pooled, cin, m):
        K, p, st, B = self.K, self.plan, self.store, self.B
        a, da = pooled
        W, bias = st.param(key + ".weight"), st.param(key + ".bias")
        lin, inv, dlin = K.empty(B, 128), K.empty(B), K.empty(B, 128)
        fm = self.f[:, 128 * m:128 * (m + 1)]
        dfm = self.df[:, 128 * m:128 * (m + 1)]
        p.f(K.gemm, a, W, bias, lin, 1, B, 128, cin, cin, 1, 1, cin, 128, 0, 0, 0, 1.0, 0)
        p.f(K.gather_l2norm, lin, 128, None, 0, 1, B, 128, fm, 384, inv)

        def backward():
            p.b(K.gather_l2norm_bwd, dfm, 384, fm, 384, inv, None, 0, 1, B, 128, dlin, 128, 0)
            p.b(K.colsum_small, dlin, B, 128, 128, st.grad(key + ".bias"), 0)
            # dW[n][k] = sum_b dlin[b][n] * a[b][k]
            p.b(K.gemm, dlin, a, None, st.grad(key + ".weight"), 1, 128, cin, B, 1, 128, cin, 1, cin, 0, 0, 0, 1.0, 0)
            # da[b][k] = sum_n dlin[b][n] * W[n][k]
            p.b(K.gemm, dlin, W, None, da, 1, B, cin, 128, 128, 1, cin, 1, cin, 0, 0, 0, 1.0, 0)

        p.on_backward(backward)

    # ---- encoder{1,2}_linear over the bilinear merge of the four branches (build_backbone.py:247-254, 289-294).
    # A 1x1 convolution commutes with bilinear interpolation (the interpolation weights sum to one), so each
    # branch is projected at its own resolution (K = C_j) and the four 128-channel results are upsampled and
    # summed by fuse_sum: the [B,Cm,h,h] merged map is never materialised.
    def _projection(self, key, feats):
        K, p, st, B, h = self.K, self.plan, self.store, self.B, self.h
        wfull, bias = st.param(key + ".weight"), st.param(key + ".bias")
        gfull = st.grad(key + ".weight")
        cm = self.cm
        ws, gs, ys, o = [], [], [], 0
        for j, ft in enumerate(feats):
            # column block j of the [128, Cm] matrix: element (n, c) at n*Cm + o + c  (row stride ldw = Cm)
            ws.append(wfull[o:])
            gs.append(gfull[o:])
            o += ft.C
            ft.consumers += 1
            y = K.empty(B, ft.H, ft.W, 128)
            # (the projection always runs on the tensor-core kernel, also in the exact-fp32 conv mode: its weight is a
            #  strided column block of the checkpoint tensor, which only the tc pack / wgrad kernels address)
            if not K.tc_conv_supported(B, ft.H, ft.W, ft.C, 128, 1, 1):
                raise NotImplementedError("1x1 projection: unsupported geometry")
            wp = K.empty((K.tc_conv_wpack_bytes(B, ft.H, ft.W, ft.C, 128, 1) + 3) // 4)
            self._pack_job(ws[j], cm, wp, (B, ft.H, ft.W), ft.C, 128, 1, 0)
            p.f(K.tc_conv, ft.data, wp, None, y, B, ft.H, ft.W, ft.C, 128, 1, 1, None, None, 0, 0)
            ys.append(y)
        out = Act(K.empty(B, h, h, 128), B, h, h, 128)
        log2f = torch.tensor([0, 1, 2, 3], dtype=torch.int32)
        p.f(K.fuse_sum, 4, ys, None, None, log2f, bias, 0, out.data, B, h, h, 128)

        def backward():
            G = out.grad
            assert G is not None and out.grad_ready
            P0 = B * h * h
            nparts = K.colstat_rows(P0, 128)
            p.b(K.bn_stats, G, P0, 128, self.part)
            p.b(K.colsum_finalize, self.part, nparts, 128, st.grad(key + ".bias"), 0)
            for j, ft in enumerate(feats):
                if j == 0:
                    Gj = G
                else:
                    Gj = K.empty(B, ft.H, ft.W, 128)
                    p.b(K.upsample_adjoint, G, Gj, 0, B, h, h, 128, j)
                p.b(K.tc_wgrad, ft.data, Gj, gs[j], cm, B, ft.H, ft.W, ft.C, 128, 1, 1, None, None, 0)
                gx, acc = p.grad(ft)
                wpt = K.empty((K.tc_conv_wpack_bytes(B, ft.H, ft.W, 128, ft.C, 1) + 3) // 4)
                self._pack_job(ws[j], cm, wpt, (B, ft.H, ft.W), 128, ft.C, 1, 1)
                p.b(K.tc_conv, Gj, wpt, None, gx, B, ft.H, ft.W, 128, ft.C, 1, 1, None, None, 0, acc)

        p.on_backward(backward)
        return out

    # ---- sample-level memory-bank NCE (mem_bank.py:172-205; contrast_trainer.py:212-253)
    def _nce(self):
        K, p, B, K1 = self.K, self.plan, self.B, self.K1
        self.logits = K.empty(6, B, K1)
        lse, l0, hit, coef = K.empty(6, B), K.empty(6, B), K.empty(6, B), K.empty(6, B)
        f = self.f
        x1, x2, x3 = f[:, 0:128], f[:, 128:256], f[:, 256:384]
        self._nce_args = (x1, x2, x3)

        def fwd_logits():
            b = self.banks
            K.nce_logits(b[0], b[1], b[2], x1, x2, x3, 384, self.nce_idx, B, K1, 128, self.T_nce, self.logits)

        def bwd():
            b = self.banks
            K.nce_bwd(b[0], b[1], b[2], x1, x2, x3, 384, self.nce_idx, B, K1, 128, self.T_nce, self.logits, lse, coef,
                      1.0, self.df, 384)

        p.f(fwd_logits)
        p.f(K.nce_loss, self.logits, B, K1, self.use_depth, None, lse, l0, hit, coef, self.losses[0:6], self.accs[0:6])

        def backward():
            p.b(K.zero, self.df, self.df.numel() * self.df.element_size())
            p.b(bwd)
            if not self.feat3_slot["ready"]:
                # first stage: no objective writes d(loss)/d(feat3) before the joint mean does.  Its slot is zeroed HERE (the loss
                # part of the program) and marked ready, so the model part accumulates into it: the autograd bridge, which skips the
                # loss part and seeds the slot with the caller's gradient (`seed_output_grads`), then keeps that gradient
                g, _ = self._slot_grad(self.feat3_slot, (B, self.J, 128))
                p.b(K.zero, g, g.numel() * g.element_size())

        p.on_backward(backward)

    # ---- dense intra-sample, sparse joint<->pixel and cross-subject SCL (contrast_trainer.py:642-892)
    def _stage2_losses(self):
        K, p, B, J, h, S, T = self.K, self.plan, self.B, self.J, self.h, self.S, self.T
        G1, G2 = self.lm1, self.lm2
        HW = h * h
        G1.consumers += 3
        G2.consumers += 3
        # --- dense
        kept = K.empty(B)
        stat = K.zeros(B, 2, S, 4)
        fin_d = K.zeros(8)
        p.f(K.dense_kept, self.depth_mask, B, self.R, h, kept)
        # one fused kernel: gather + L2-norm + S x S x 128 affinity (tcgen05) + soft-target statistics; L[b][i][j] = <d_i, a_j>/T
        # stays on chip (the unfused chain gather_l2norm -> gemm -> dense_stats remains in the C-ABI and the kernel tests)
        dwork = K.empty((K.dense_affinity_work_bytes(B, S) + 3) // 4)     # normalised bf16 hi/lo operand slabs (forward -> backward)
        p.f(K.dense_affinity_fwd, G1.data, G2.data, self.dense_idx, kept, self.use_depth, B, S, h, 128, 1.0 / T, stat, fin_d, dwork)
        p.f(K.bn_apply, fin_d, None, None, None, None, None, 0, self.losses[6:8], 1, 2)
        p.f(K.bn_apply, fin_d[2:4], None, None, None, None, None, 0, self.accs[6:8], 1, 2)
        # --- joints: F rows [0,BJ) rgb pixels, [BJ,2BJ) depth pixels (also the SCL feature matrix)
        N = 2 * B * J
        pixj = K.zeros(B, J, dtype=torch.int64)
        Fm, inv_f, dF = K.empty(N, 128), K.empty(N), K.empty(N, 128)
        Sk, inv_s, dSk = K.empty(B * J, 128), K.empty(B * J), K.empty(B * J, 128)
        Lr, Ldj = K.empty(B, J, J), K.empty(B, J, J)
        rs, lsej, fin_j = K.empty(B, 2, 3), K.empty(B, 2, J), K.zeros(8)
        Fa, Fd = Fm[:B * J], Fm[B * J:]
        dFa, dFd = dF[:B * J], dF[B * J:]
        p.f(K.joint_pixel_index, self.joints_yx, B * J, h, pixj)
        p.f(K.gather_l2norm, G1.data, 0, pixj, HW, J, B * J, 128, Fa, 128, inv_f[:B * J])
        p.f(K.gather_l2norm, G2.data, 0, pixj, HW, J, B * J, 128, Fd, 128, inv_f[B * J:])
        p.f(K.gather_l2norm, self.feat3, 128, None, 0, 1, B * J, 128, Sk, 128, inv_s)
        # Lr[b][k][j] = <s_k, a_j>/T
        p.f(K.gemm, Sk, Fa, None, Lr, B, J, J, 128, 128, 1, 1, 128, J, J * 128, J * 128, J * J, 1.0 / T, 0)
        p.f(K.gemm, Sk, Fd, None, Ldj, B, J, J, 128, 128, 1, 1, 128, J, J * 128, J * 128, J * J, 1.0 / T, 0)
        p.f(K.joint_stats, Lr, Ldj, self.joints_vis, self.use_depth, B, J, rs, lsej, fin_j)
        p.f(K.bn_apply, fin_j, None, None, None, None, None, 0, self.losses[8:10], 1, 2)
        p.f(K.bn_apply, fin_j[2:4], None, None, None, None, None, 0, self.accs[8:10], 1, 2)
        # --- SCL
        Z = K.empty(N, N)
        rowstat, fin_s = K.empty(N, 3), K.zeros(4)
        p.f(K.gemm, Fm, Fm, None, Z, 1, N, N, 128, 128, 1, 1, 128, N, 0, 0, 0, 1.0 / T, 0)
        p.f(K.scl_stats, Z, B, J, None, self.use_depth, rowstat, fin_s)
        p.f(K.bn_apply, fin_s, None, None, None, None, None, 0, self.losses[10:11], 1, 1)
        self.stage2_debug = dict(stat=stat, Lr=Lr, Ldj=Ldj, Z=Z, fin_d=fin_d, fin_j=fin_j, fin_s=fin_s, kept=kept, pixj=pixj)

        def backward():
            iT = 1.0 / T
            g1, g2 = p.zeroed_grad(G1), p.zeroed_grad(G2)
            # joints: logits -> gradients in place; dF (rgb | depth rows), dSk
            p.b(K.joint_grad, Lr, Ldj, self.joints_vis, self.use_depth, lsej, fin_j, B, J, 1.0)
            # dSk[b][k][:] = iT * (sum_j dLr[k][j] a_j + sum_j dLd[k][j] d_j)
            p.b(K.gemm, Lr, Fa, None, dSk, B, J, 128, J, J, 1, 128, 1, 128, J * J, J * 128, J * 128, iT, 0)
            p.b(K.gemm, Ldj, Fd, None, dSk, B, J, 128, J, J, 1, 128, 1, 128, J * J, J * 128, J * 128, iT, 1)
            # dFa[b][j][:] = iT * sum_k dLr[k][j] s_k
            p.b(K.gemm, Lr, Sk, None, dFa, B, J, 128, J, 1, J, 128, 1, 128, J * J, J * 128, J * 128, iT, 0)
            p.b(K.gemm, Ldj, Sk, None, dFd, B, J, 128, J, 1, J, 128, 1, 128, J * J, J * 128, J * 128, iT, 0)
            # SCL: dF += iT * (dZ + dZ^T) F
            p.b(K.scl_grad, Z, B, J, None, self.use_depth, rowstat, fin_s, 1.0)
            p.b(K.gemm, Z, Fm, None, dF, 1, N, 128, N, N, 1, 128, 1, 128, 0, 0, 0, iT, 1)
            p.b(K.gemm, Z, Fm, None, dF, 1, N, 128, N, 1, N, 128, 1, 128, 0, 0, 0, iT, 1)
            p.b(K.gather_l2norm_bwd, dFa, 128, Fa, 128, inv_f[:B * J], pixj, HW, J, B * J, 128, g1, 0, 1)
            p.b(K.gather_l2norm_bwd, dFd, 128, Fd, 128, inv_f[B * J:], pixj, HW, J, B * J, 128, g2, 0, 1)
            gs, acc = self._slot_grad(self.feat3_slot, (B, J, 128))
            p.b(K.gather_l2norm_bwd, dSk, 128, Sk, 128, inv_s, None, 0, 1, B * J, 128, gs, 128, acc)
            # dense: affinity recomputed on chip, logit gradient -> second MMA -> L2-norm backward -> atomic scatter
            p.b(K.dense_affinity_bwd, G1.data, G2.data, self.dense_idx, stat, kept, fin_d, B, S, h, 128, iT, 1.0, 1.0, g1, g2, dwork, 1)

        p.on_backward(backward)

    # ------------------------------------------------------------------ running
    def set_batch(self, batch, nce_idx, dense_idx=None):
        """batch: the reference's DataLoader tuple (see hcmoco_b200/synthetic.py) or a dict with the same fields."""
        if not isinstance(batch, dict):
            batch = dict(x=batch[0], index=batch[1], skeleton=batch[2], joints_yx=batch[4], joints_vis=batch[5],
                         use_depth=batch[6], depth_mask=batch[7])
        self.x.copy_(batch["x"], non_blocking=True)
        self.index.copy_(batch["index"], non_blocking=True)
        self.skel.copy_(batch["skeleton"], non_blocking=True)
        self.joints_yx.copy_(batch["joints_yx"], non_blocking=True)
        self.joints_vis.copy_(batch["joints_vis"], non_blocking=True)
        self.use_depth.copy_(batch["use_depth"], non_blocking=True)
        self.depth_mask.copy_(batch["depth_mask"], non_blocking=True)
        self.nce_idx.copy_(nce_idx, non_blocking=True)
        if dense_idx is not None:
            self.dense_idx.copy_(dense_idx, non_blocking=True)

    def forward(self):
        self.plan.run(self.plan.fwd, self.two_streams)

    def backward(self):
        self.K.zero(self.store.g, self.store.n * self.store.g.element_size())
        self.plan.run(self.plan.bwd, self.two_streams)

    # ---- model-only programs for the autograd bridge (api.HCMoCoModel): the caller owns the losses
    def forward_model(self):
        self.plan.run(self.plan.fwd[:self.n_model_fwd], self.two_streams)

    def seed_output_grads(self, gf, g_feat3=None, g_lm1=None, g_lm2=None):
        """Write d(loss)/d(outputs) into the buffers the model-backward program starts from."""
        B, J, h = self.B, self.J, self.h
        self.df.copy_(gf) if gf is not None else self.df.zero_()
        slot = self.feat3_slot
        if slot["grad"] is None:
            slot["grad"] = self.K.empty(B, J, 128)
        if g_feat3 is not None:
            slot["grad"].copy_(g_feat3)
        else:
            slot["grad"].zero_()
        if self.stage == 2:
            for act, g in ((self.lm1, g_lm1), (self.lm2, g_lm2)):
                if g is not None:
                    act.grad.copy_(g.permute(0, 2, 3, 1))
                else:
                    act.grad.zero_()

    def backward_model(self):
        self.K.zero(self.store.g, self.store.n * self.store.g.element_size())
        self.plan.run(self.plan.bwd[self.n_loss_bwd:], self.two_streams)

    def draw_dense(self, injected=None):
        """Dense pixel samples: S draws with replacement from each sample's nearest-resized depth mask
        (contrast_trainer.py:674-685); rows of samples the mask drops are ignored downstream."""
        if injected is not None:
            self.dense_idx.copy_(injected)
            return
        step = self.R // self.h
        m = self.depth_mask[:, ::step, ::step][:, :self.h, :self.h].reshape(self.B, -1)
        has = m.sum(1, keepdim=True) > 0
        w = torch.where(has, (m != 0).to(m.dtype), torch.ones_like(m))
        self.dense_idx.copy_(torch.multinomial(w, self.S, replacement=True))

    def update_banks(self, all_f=None, all_index=None):
        """mem_bank.py:195-203 with the all-gathered embeddings (contrast_trainer.py:578-579)."""
        f = self.f if all_f is None else all_f
        y = self.index if all_index is None else all_index
        N = f.shape[0]
        for m in range(3):
            self.K.bank_update(self.banks[m], f[:, 128 * m:128 * (m + 1)], 384, y, N, 128, self.m_nce)

    def sgd(self, lr=0.03, momentum=0.9, wd=1e-4, gscale=1.0):
        # the momentum buffer starts at zero, so torch.optim.SGD's "first step: buf = grad" is the same update
        st = self.store
        self.K.sgd_step(st.p, st.g, st.m, st.n, lr, momentum, wd, 0, gscale)
        self.first_step = False

    def step(self, lr=0.03, momentum=0.9, wd=1e-4):
        """Single-rank step: forward, losses, backward, bank update, SGD."""
        self.forward()
        self.backward()
        self.update_banks()
        self.sgd(lr, momentum, wd)

    # ---- CUDA graph: forward + losses + backward are one graph launch (every buffer is static)
    def capture(self):
        saved = self.store.save_buffers()
        self.forward()                     # warm-up outside capture (module loading), then undo its BN side effects
        self.backward()
        self.store.restore_buffers(saved)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.forward()
            self.backward()
        self.store.restore_buffers(saved)
        return self

    def capture_parts(self):
        """Three CUDA graphs instead of one — forward (model + loss kernels), loss part of the backward (writes d(loss)/d(outputs):
        embedding, skeleton-feature and projection-map gradients), model part of the backward — so that a caller can add its own
        gradient to the projection maps between the second and the third (the segmentation head of `segment.SegTrainer`, whose
        batch of labelled samples changes size from step to step and therefore launches eagerly)."""
        saved = self.store.save_buffers()
        self.forward()
        self.backward()
        self.store.restore_buffers(saved)
        torch.cuda.synchronize()
        K, st, p = self.K, self.store, self.plan
        self.graph_parts = [torch.cuda.CUDAGraph() for _ in range(3)]
        with torch.cuda.graph(self.graph_parts[0]):
            self.forward()
        with torch.cuda.graph(self.graph_parts[1]):
            K.zero(st.g, st.n * st.g.element_size())
            p.run(p.bwd[:self.n_loss_bwd], self.two_streams)
        with torch.cuda.graph(self.graph_parts[2]):
            p.run(p.bwd[self.n_loss_bwd:], self.two_streams)
        self.store.restore_buffers(saved)
        return self

    def step_graph(self, lr=0.03, momentum=0.9, wd=1e-4):
        self.graph.replay()
        self.update_banks()
        self.sgd(lr, momentum, wd)

    @property
    def launches_per_step(self):
        """C-ABI calls in one step (each enqueues at least one kernel)."""
        return sum(1 for e in self.plan.fwd + self.plan.bwd if e[0] is not FORK and e[0] is not JOIN) + 1 + 3 + 1

    def total_loss(self):
        n = 11 if self.stage == 2 else 6
        return self.losses[:n].sum()

    def results_async(self):
        """Enqueue the D2H read of this step's losses / accuracies (32 floats into pinned memory) behind the step and return
        a callable that waits for it and unpacks: a training loop calls it one step later, so the host never idles the GPU."""
        if not hasattr(self, "_res_host"):
            self._res_host = [torch.empty(32, pin_memory=True) for _ in range(2)]
            self._res_slot = 0
        host = self._res_host[self._res_slot]
        self._res_slot ^= 1
        host[:16].copy_(self.losses, non_blocking=True)
        host[16:].copy_(self.accs, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()

        def wait():
            ev.synchronize()
            return self._unpack(host.clone())

        return wait

    def results(self):
        """Host copy of the step's losses and accuracies (one D2H read)."""
        return self._unpack(torch.cat([self.losses, self.accs]).cpu())

    def _unpack(self, la):
        ls, ac = la[:16], la[16:]
        out = dict(nce_losses=ls[0:6], nce_accs=ac[0:6])
        if self.stage == 2:
            out.update(dense_losses=ls[6:8], dense_accs=ac[6:8], joint_losses=ls[8:10], joint_accs=ac[8:10],
                       scl_loss=ls[10])
        out["loss"] = ls[:11 if self.stage == 2 else 6].sum()
        return out

    # NCHW views for the drop-in surface / tests
    @staticmethod
    def nchw(act):
        return act.data.permute(0, 3, 1, 2)
