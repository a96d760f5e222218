"""The reference's call surface for the pre-train path, backed by the sm_100a engine.

`main_contrast.py` (pycontrast/main_contrast.py:36-106) uses exactly these objects; same names, argument
meaning, return shapes and checkpoint layout, so the scripts/ launchers work unchanged when
`hcmoco_b200/compat/` is first on sys.path:

    build_model(opt)            -> (model, None)          networks/build_backbone.py:525-566
    model(x, s, mode=0, return_fm=False)                  networks/build_backbone.py:256-303
    build_mem(opt, n_data)      -> CMCMem3-like module    memory/build_memory.py:5-17, memory/mem_bank.py:157-205
    ContrastTrainer(opt)        -> trainer                learning/contrast_trainer.py, learning/base_trainer.py
    build_contrast(opt)         -> the same trainer (factory named by the north star; absent in the reference)

The model and the memory are torch.nn.Modules whose parameters / buffers ALIAS the engine's flat device
buffers: `state_dict()` has the reference's 3741 / 3745 keys and shapes, `loss.backward()` works (one
autograd.Function around the engine's launch programs), `torch.optim.SGD(model.parameters())` works.
The trainer's own step loop does not go through autograd: it replays the engine's fused program
(forward + the loss kernels + backward as one CUDA graph), all-gathers embeddings, updates the banks,
all-reduces the flat gradient buffer and applies the fused SGD kernel.

There is no CPU fallback: without a CUDA device / the built library, `build_model` raises.
"""
import math
import os
import sys
import time
from collections import OrderedDict

import torch
import torch.nn as nn

from . import layout as L
from .engine import Engine, ParamStore
from .pretrain import init_parameters


def _kernels(kernels=None):
    if kernels is not None:
        return kernels
    from .kernels import CudaKernels
    return CudaKernels()


def _stage_of(opt):
    return 2 if int(getattr(opt, "linear_feat_map", 0)) else 1


# ---------------------------------------------------------------------------------------------- model
class _Node(nn.Module):
    """Empty container; parameters / buffers are attached under the reference's attribute names."""


class _ModelFn(torch.autograd.Function):
    """autograd bridge: forward = engine model program, backward = engine model-backward program."""

    @staticmethod
    def forward(ctx, model, x, s, return_fm, *params):
        eng = model.engine_for(x.shape[0], x.shape[-1])
        eng.x.copy_(x)
        eng.skel.copy_(s)
        eng.forward_model()
        ctx.model, ctx.eng, ctx.return_fm = model, eng, return_fm
        outs = [eng.f.clone()]
        if return_fm:
            outs.append(eng.feat3.clone())
            if eng.stage == 2:
                outs += [eng.nchw(eng.lm1).clone(), eng.nchw(eng.lm2).clone()]
        return tuple(outs)

    @staticmethod
    def backward(ctx, gf, *gouts):
        eng, model = ctx.eng, ctx.model
        eng.seed_output_grads(gf, gouts[0] if len(gouts) > 0 else None,
                              gouts[1] if len(gouts) > 1 else None, gouts[2] if len(gouts) > 2 else None)
        eng.backward_model()
        # hand autograd views of a COPY of the flat gradient buffer (one device-to-device copy): AccumulateGrad keeps the
        # tensors it is given as `p.grad`, and the engine rewrites `store.g` on the next backward — aliasing the two would
        # double the gradient from the second `zero_grad(set_to_none=False); backward()` on
        st = model.store
        flat = st.g.clone()
        grads = tuple(st.view(flat, k).view(st.keys[k]) for k in model.param_keys)
        return (None, None, None, None) + grads


class HCMoCoModel(nn.Module):
    """CMC3HRNetSGCNSingleHead (RGBD2S + HRNet) with engine-owned storage."""

    def __init__(self, opt, kernels=None):
        super().__init__()
        assert getattr(opt, "modal", "RGBD2S") == "RGBD2S" and getattr(opt, "arch", "HRNet") == "HRNet", \
            "hcmoco_b200 implements the RGBD2S / HRNet path (build_backbone.py:516-523 key 'RGBD2SHRNetSin')"
        assert not getattr(opt, "jigsaw", False) and getattr(opt, "head", "linear") == "linear"
        assert getattr(opt, "pool_method", "mean") == "mean"
        self.opt = opt
        self.K = _kernels(kernels)
        self.width = int(getattr(opt, "width", 18))
        self.stage = _stage_of(opt)
        self.skeleton = getattr(opt, "skeleton_meta_name", "mpii")
        self.in_channel_list = list(getattr(opt, "in_channel_list", [3, 3]))
        self.linear_feat_map = bool(self.stage == 2)
        self.store = ParamStore(self.K, L.model_keys(self.width, self.stage, self.skeleton, int(getattr(opt, "feat_dim", 128))))
        init_parameters(self.store, seed=int(getattr(opt, "seed", None) or 0))
        self.param_keys = []
        for k, shp in self.store.keys.items():
            node, parts = self, k.split(".")
            for a in parts[:-1]:
                if not hasattr(node, a):
                    node.add_module(a, _Node())
                node = getattr(node, a)
            if L.is_buffer(k):
                node.register_buffer(parts[-1], self.store.buffers[k])
            else:
                prm = nn.Parameter(self.store.view(self.store.p, k).view(shp))
                node.register_parameter(parts[-1], prm)
                self.param_keys.append(k)
        self._engines = {}
        self._mem = None
        self._load_imagenet(opt)

    # IN_Pretrain / depth_Pretrain: partial load of ImageNet HRNet weights (build_backbone.py:531-560)
    def _load_imagenet(self, opt):
        for attr, enc in (("IN_Pretrain", "encoder1"), ("depth_Pretrain", "encoder2")):
            path = getattr(opt, attr, None)
            if path:
                sd = torch.load(path, map_location="cpu")
                own = getattr(self, enc).state_dict()
                sel = {k: v for k, v in sd.items() if k in own and tuple(v.shape) == tuple(own[k].shape)}
                getattr(self, enc).load_state_dict(sel, strict=False)

    def engine_for(self, B, R, mem=None):
        """The static launch program for this (batch, resolution); built on first use."""
        mem = mem or self._mem
        key = (int(B), int(R))
        if key not in self._engines:
            o = self.opt
            n_data = mem.n_data if mem is not None else 2
            eng = Engine(self.K, self.width, self.stage, self.skeleton, int(B), int(R), n_data,
                         int(getattr(o, "nce_k", 16384)), float(getattr(o, "nce_t", 0.07)), float(getattr(o, "nce_m", 0.5)),
                         float(getattr(o, "temperature", 0.07)), int(getattr(o, "pri3d_num_samples_per_image", 400)),
                         store=self.store)
            if mem is None:
                eng.init_banks()          # placeholder rows; only the model programs run without a memory
            eng.build()
            self._engines[key] = eng
        eng = self._engines[key]
        if mem is not None:
            eng.banks = [mem.memory_1, mem.memory_2, mem.memory_3]
        return eng

    def attach_memory(self, mem):
        self._mem = mem
        for eng in self._engines.values():
            eng.banks = [mem.memory_1, mem.memory_2, mem.memory_3]

    def forward(self, x, s, mode=0, return_fm=False):
        assert mode == 0, "mode 1 (momentum encoder) / 2 (test) belong to the MoCo / linear-probe paths"
        params = [p for p in self.parameters()]
        outs = _ModelFn.apply(self, x.to(self.K.dtype), s.to(self.K.dtype), bool(return_fm), *params)
        f = outs[0]
        if not return_fm:
            return f
        eng = self.engine_for(x.shape[0], x.shape[-1])
        feat1 = [eng.nchw(a).detach() for a in eng.feat1]
        feat2 = [eng.nchw(a).detach() for a in eng.feat2]
        if self.stage == 2:
            # merge1 / merge2 (the 270-channel maps) are never materialised by the engine: the projection is applied
            # per branch before upsampling; the reference only consumes linear_merge{1,2}
            return feat1, feat2, outs[1], f, {"merge1": None, "merge2": None, "linear_merge1": outs[2],
                                               "linear_merge2": outs[3]}
        B = x.shape[0]
        avg1 = torch.cat([a.mean((2, 3)) for a in feat1], 1)
        avg2 = torch.cat([a.mean((2, 3)) for a in feat2], 1)
        return feat1, feat2, outs[1], avg1, avg2, outs[1].mean(1), f

    def cuda(self, device=None):          # storage already lives on the device
        return self

    def train(self, mode=True):
        assert mode, "the engine implements train-mode BatchNorm (the pre-train path)"
        return super().train(mode)


def build_model(opt, kernels=None):
    """networks/build_backbone.py:525-566: returns (model, model_ema); model_ema is None for the bank methods."""
    if getattr(opt, "mem", "bank") == "moco":
        raise NotImplementedError("MoCo queues are outside the RGBD2S pre-train path (SURVEY.md section 2)")
    return HCMoCoModel(opt, kernels), None


# ---------------------------------------------------------------------------------------------- memory
class _NceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mem, x, idx):
        K, B, K1 = mem.K_, x.shape[0], idx.shape[1]
        x = x.contiguous()
        logits = K.empty(6, B, K1)
        K.nce_logits(mem.memory_1, mem.memory_2, mem.memory_3, x[:, 0:128], x[:, 128:256], x[:, 256:384], 384, idx, B, K1,
                     128, mem.T, logits)
        ctx.mem, ctx.x, ctx.idx = mem, x, idx
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        mem, x, idx = ctx.mem, ctx.x, ctx.idx
        K, B, K1 = mem.K_, x.shape[0], idx.shape[1]
        df = K.zeros(B, 384)
        banks = (mem.memory_1, mem.memory_2, mem.memory_3)
        # swap the pre-update rows back in for the re-gather (see HCMoCoMem.forward), then restore the updated ones
        cur = [b.index_select(0, mem.touched) for b in banks]
        for b, old in zip(banks, mem.old_rows):
            b.index_copy_(0, mem.touched, old)
        # lse = None: `logits` carries d(loss)/d(logits) itself
        K.nce_bwd(banks[0], banks[1], banks[2], x[:, 0:128], x[:, 128:256], x[:, 256:384], 384, idx, B, K1, 128, mem.T,
                  dlogits.contiguous(), None, None, 1.0, df, 384)
        for b, c in zip(banks, cur):
            b.index_copy_(0, mem.touched, c)
        return None, df, None


class HCMoCoMem(nn.Module):
    """CMCMem3 (memory/mem_bank.py:157-205): three [n_data,128] banks, uniform negatives, momentum update."""

    def __init__(self, n_dim, n_data, K=16384, T=0.07, m=0.5, kernels=None, seed=0):
        super().__init__()
        assert n_dim == 128
        self.K_ = _kernels(kernels)
        self.K, self.T, self.m, self.n_data = K, T, m, n_data
        g = torch.Generator().manual_seed(seed)
        for i in (1, 2, 3):
            bank = torch.nn.functional.normalize(torch.randn(n_data, n_dim, generator=g))
            self.register_buffer("memory_%d" % i, bank.to(self.K_.device, self.K_.dtype))
        self.gen = None
        self.injected_idx = None      # tests inject the draw

    def cuda(self, device=None):
        return self

    def draw(self, bsz, y):
        """AliasMethod.draw over uniform probabilities == uniform integers (alias_multinomial.py:49-65); idx[:,0] = y."""
        if self.injected_idx is not None:
            idx = self.injected_idx.to(y.device)
        else:
            idx = torch.randint(0, self.n_data, (bsz, self.K + 1), device=y.device, generator=self.gen)
        idx[:, 0] = y
        return idx

    def update(self, x, y):
        for i in range(3):
            self.K_.bank_update(getattr(self, "memory_%d" % (i + 1)), x[:, 128 * i:128 * (i + 1)], x.stride(0), y,
                                x.shape[0], 128, self.m)

    def forward(self, x1, x2, x3, y, all_x1=None, all_x2=None, all_x3=None, all_y=None):
        bsz = x1.shape[0]
        assert bsz >= 2, "the reference collapses bsz=1 (mem_bank.py:39 out.squeeze())"
        idx = self.draw(bsz, y)
        x = torch.cat((x1, x2, x3), 1)
        logits = _NceFn.apply(self, x, idx)
        if all_x1 is not None and all_x2 is not None and all_x3 is not None and all_y is not None:
            ux, uy = torch.cat((all_x1, all_x2, all_x3), 1).detach().contiguous(), all_y
        else:
            ux, uy = x.detach().contiguous(), y
        # the backward re-gathers bank rows and must see them as they were when the logits were computed (autograd in the
        # reference holds the gathered copies): keep the pre-update contents of the few rows the update touches
        self.touched = uy.clone()
        self.old_rows = [getattr(self, "memory_%d" % i).index_select(0, uy) for i in (1, 2, 3)]
        self.update(ux, uy)
        labels = torch.zeros(bsz, dtype=torch.long, device=x.device)
        return (logits[0], logits[1], logits[2], logits[3], logits[4], logits[5], labels)


def build_mem(opt, n_data, kernels=None):
    """memory/build_memory.py:5-17."""
    if not str(getattr(opt, "mem", "bank")).startswith("bank"):
        raise NotImplementedError("mem not supported: {}".format(opt.mem))
    return HCMoCoMem(int(getattr(opt, "feat_dim", 128)), n_data, int(opt.nce_k), float(opt.nce_t), float(opt.nce_m), kernels)


# ---------------------------------------------------------------------------------------------- optimiser
class FusedSGD(torch.optim.Optimizer):
    """torch.optim.SGD semantics (momentum, coupled weight decay) as ONE kernel over the engine's flat buffers.
    state_dict() / load_state_dict() use torch's format (per-parameter `momentum_buffer`), so checkpoints interchange."""

    def __init__(self, model, lr=0.03, momentum=0.9, weight_decay=1e-4):
        self.model = model
        super().__init__(list(model.parameters()), dict(lr=lr, momentum=momentum, weight_decay=weight_decay,
                                                        dampening=0, nesterov=False))
        st = model.store
        for k, p in zip(model.param_keys, self.param_groups[0]["params"]):
            self.state[p]["momentum_buffer"] = st.view(st.m, k).view(st.keys[k])

    @classmethod
    def from_torch(cls, model, optimizer):
        g = optimizer.param_groups[0]
        return cls(model, g["lr"], g.get("momentum", 0.0), g.get("weight_decay", 0.0))

    def zero_grad(self, set_to_none=True):
        """torch.optim.Optimizer.zero_grad semantics for `p.grad` (the autograd path), plus the engine's flat buffer."""
        st = self.model.store
        self.model.K.zero(st.g, st.n * st.g.element_size())
        flat = self._autograd_flat()
        for p in self.param_groups[0]["params"]:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                elif flat is None:
                    p.grad.zero_()
        if flat is not None and not set_to_none:
            flat.zero_()

    def _autograd_flat(self):
        """The flat tensor all `p.grad`s are views of, in store layout (what `_ModelFn.backward` hands to autograd), or None."""
        st, ps = self.model.store, self.param_groups[0]["params"]
        g0 = ps[0].grad
        if g0 is None or g0._base is None:
            return None
        base = g0._base
        if base.numel() != st.n or base.dtype != st.g.dtype:
            return None
        es = base.element_size()
        for k, p in zip(self.model.param_keys, ps):
            if p.grad is None or p.grad._base is not base or p.grad.data_ptr() != base.data_ptr() + st.off[k][0] * es:
                return None
        return base

    @torch.no_grad()
    def step(self, closure=None, gscale=1.0, from_autograd=None):
        """One fused launch.  The trainer's fused step leaves the gradient in `store.g` (from_autograd=False); after a
        `loss.backward()` through `model(...)` the gradients live in `p.grad` (from_autograd=True; default: decided by
        whether any `p.grad` is set) — views of one flat tensor in the common case, gathered per key otherwise."""
        g, st = self.param_groups[0], self.model.store
        ps = self.param_groups[0]["params"]
        if from_autograd is None:
            from_autograd = any(p.grad is not None for p in ps)
        src = st.g
        if from_autograd:
            flat = self._autograd_flat()
            if flat is not None:
                src = flat
            else:
                for k, p in zip(self.model.param_keys, ps):
                    v = st.view(st.g, k)
                    v.zero_() if p.grad is None else v.copy_(p.grad.reshape(-1))
        self.model.K.sgd_step(st.p, src, st.m, st.n, g["lr"], g["momentum"], g["weight_decay"], 0, gscale)

    def load_state_dict(self, sd):
        st = self.model.store
        for i, k in enumerate(self.model.param_keys):
            buf = sd["state"].get(i, {}).get("momentum_buffer")
            if buf is not None:
                st.load(st.m, k, buf)
        for k in ("lr", "momentum", "weight_decay"):
            self.param_groups[0][k] = sd["param_groups"][0][k]


# ---------------------------------------------------------------------------------------------- trainer
class AverageMeter(object):            # learning/util.py:6-21
    def __init__(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class ContrastTrainer(object):
    """learning/contrast_trainer.py + learning/base_trainer.py, same public methods."""

    def __init__(self, args):
        self.args = args
        self.local_group = None
        self.logger = None
        self.graphs = {}
        self.stagers = {}

    # ---- distributed environment (base_trainer.py:20-73).  torchrun / SLURM env; one process per GPU; NVSwitch makes the
    # per-node process groups of the reference (ShuffleBN only) unnecessary.
    def init_ddp_environment(self, gpu=0, ngpus_per_node=1):
        import torch.distributed as dist
        a = self.args
        env = os.environ
        rank = int(env.get("RANK", env.get("SLURM_PROCID", 0)))
        world = int(env.get("WORLD_SIZE", env.get("SLURM_NTASKS", 1)))
        local = int(env.get("LOCAL_RANK", rank % max(1, torch.cuda.device_count() or 1)))
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        if world > 1 and not dist.is_initialized():
            if "MASTER_ADDR" not in env and "SLURM_NODELIST" in env:
                # base_trainer.py:36-44: first host of the SLURM allocation (multi-node srun launches)
                import subprocess
                addr = subprocess.getoutput("scontrol show hostname {} | head -n1".format(env["SLURM_NODELIST"])).strip()
                env["MASTER_ADDR"] = addr if addr and " " not in addr else env["SLURM_NODELIST"].split(",")[0]
            env.setdefault("MASTER_ADDR", "127.0.0.1")
            env.setdefault("MASTER_PORT", "29500")
            dist.init_process_group(backend=getattr(a, "dist_backend", "nccl") if torch.cuda.is_available() else "gloo",
                                    rank=rank, world_size=world)
        # as the reference (base_trainer.py:40-47: local_rank = rank = SLURM_PROCID): `local_rank` is the GLOBAL rank, so that
        # `save()` (gated on local_rank == 0, contrast_trainer.py:120) writes once per job, not once per node
        a.gpu, a.rank, a.local_rank, a.world_size, a.ngpus_per_node = local, rank, rank, world, ngpus_per_node
        a.distributed = world > 1

    def wrap_up(self, model, model_ema, optimizer):
        """No DDP wrapper: gradients live in one flat buffer that is all-reduced in a single NCCL call per step
        (contrast_trainer.py:50-79 wraps in DDP; the state_dict keys written by save() keep its 'module.' prefix)."""
        if not isinstance(optimizer, FusedSGD):
            optimizer = FusedSGD.from_torch(model, optimizer)
        return model, model_ema, optimizer

    def broadcast_memory(self, contrast):
        """All three banks from rank 0 (the reference skips memory_3)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            for i in (1, 2, 3):
                dist.broadcast(getattr(contrast, "memory_%d" % i), 0)

    def resume_model(self, model, model_ema, contrast, optimizer):
        a, start_epoch = self.args, 1
        if getattr(a, "resume", ""):
            if os.path.isfile(a.resume):
                ck = torch.load(a.resume, map_location="cpu")
                start_epoch = ck["epoch"] + 1
                model.store.load_state_dict(ck["model"])
                contrast.load_state_dict(ck["contrast"])
                optimizer.load_state_dict(ck["optimizer"])
                print("=> resume successfully '{}' (epoch {})".format(a.resume, ck["epoch"]))
            else:
                print("=> no checkpoint found at '{}'".format(a.resume))
        return start_epoch

    def save(self, model, model_ema, contrast, optimizer, epoch):
        a = self.args
        if getattr(a, "local_rank", 0) == 0:
            print("==> Saving...")
            state = {"model": model.store.state_dict(prefix="module."), "contrast": contrast.state_dict(),
                     "optimizer": optimizer.state_dict(), "epoch": epoch}
            torch.save(state, os.path.join(a.model_folder, "current.pth"))
            if epoch % a.save_freq == 0:
                torch.save(state, os.path.join(a.model_folder, "ckpt_epoch_{}.pth".format(epoch)))

    def init_tensorboard_logger(self):
        if getattr(self.args, "rank", 0) == 0:
            try:
                import tensorboard_logger as tb_logger
                self.logger = tb_logger.Logger(logdir=self.args.tb_folder, flush_secs=2)
            except Exception:          # optional dependency of the reference; absent here
                self.logger = None

    def logging(self, epoch, logs, lr):
        if getattr(self.args, "rank", 0) == 0 and self.logger is not None:
            for name, v in zip(("loss", "acc", "jig_loss", "jig_acc"), logs):
                self.logger.log_value(name, v, epoch)
            self.logger.log_value("learning_rate", lr, epoch)

    def adjust_learning_rate(self, optimizer, epoch):          # base_trainer.py:80-93
        a = self.args
        lr = a.learning_rate
        if a.cosine:
            eta_min = lr * (a.lr_decay_rate ** 3)
            lr = eta_min + (lr - eta_min) * (1 + math.cos(math.pi * epoch / a.epochs)) / 2
        else:
            steps = sum(1 for e in a.lr_decay_epochs if epoch > e)
            if steps > 0:
                lr = lr * (a.lr_decay_rate ** steps)
        for g in optimizer.param_groups:
            g["lr"] = lr

    def warmup_learning_rate(self, epoch, batch_id, total_batches, optimizer):   # base_trainer.py:95-103
        a = self.args
        if getattr(a, "warm", False) and epoch <= a.warm_epochs:
            p = (batch_id + (epoch - 1) * total_batches) / (a.warm_epochs * total_batches)
            lr = a.warmup_from + p * (a.warmup_to - a.warmup_from)
            for g in optimizer.param_groups:
                g["lr"] = lr

    @staticmethod
    def _global_gather(x):                                     # contrast_trainer.py:160-165
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return x
        out = torch.empty((dist.get_world_size() * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous())
        return out

    # ---- one epoch (contrast_trainer.py:142-158 dispatch; :532-640 first stage; :894-1039 second stage)
    def train(self, epoch, train_loader, model, model_ema, contrast, criterion, optimizer):
        a = self.args
        assert a.mem in ("bank", "bank+jointspri3d"), "mem '%s' is not on the RGBD2S pre-train path" % a.mem
        assert (a.mem == "bank+jointspri3d") == (model.stage == 2), "--linear_feat_map 1 goes with bank+jointspri3d"
        t0 = time.time()
        model.attach_memory(contrast)
        meters = {k: AverageMeter() for k in ("bt", "dt", "loss", "a12", "a23", "a13", "r2d", "d2r", "ar2d", "ad2r", "r2j",
                                               "d2j", "ar2j", "ad2j", "scl")}
        world = getattr(a, "world_size", 1) or 1
        end = time.time()
        n_batches = len(train_loader) if hasattr(train_loader, "__len__") else 0
        for idx, data in enumerate(train_loader):
            meters["dt"].update(time.time() - end)
            self.warmup_learning_rate(epoch, idx, n_batches, optimizer)
            res = self.train_step(model, contrast, optimizer, data, world)
            bsz = data[0].shape[0]
            if (idx + 1) % a.print_freq == 0 or idx + 1 == n_batches:      # the only host sync (the reference syncs 6-13x per step)
                r = res()
                meters["loss"].update(float(r["loss"]), bsz)
                na = r["nce_accs"]
                for k, v in (("a12", 0.5 * (na[0] + na[1])), ("a23", 0.5 * (na[2] + na[3])), ("a13", 0.5 * (na[4] + na[5]))):
                    meters[k].update(float(v), bsz)
                if model.stage == 2:
                    for k, v in (("r2d", r["dense_losses"][0]), ("d2r", r["dense_losses"][1]), ("ar2d", r["dense_accs"][0]),
                                 ("ad2r", r["dense_accs"][1]), ("r2j", r["joint_losses"][0]), ("d2j", r["joint_losses"][1]),
                                 ("ar2j", r["joint_accs"][0]), ("ad2j", r["joint_accs"][1]), ("scl", r["scl_loss"])):
                        meters[k].update(float(v), bsz)
                meters["bt"].update((time.time() - end))
                if getattr(a, "local_rank", 0) == 0:
                    m = meters
                    msg = ("Train: [{0}][{1}/{2}]\t" "BT {3:.3f}\tDT {4:.3f}\tL {5:.3f} ({6:.3f})\ta_I {7:.3f} {8:.3f} {9:.3f}"
                           .format(epoch, idx + 1, n_batches, m["bt"].val, m["dt"].val, m["loss"].val, m["loss"].avg,
                                   m["a12"].avg, m["a23"].avg, m["a13"].avg))
                    if model.stage == 2:
                        msg += ("\tp3d {:.3f} {:.3f} {:.3f} {:.3f}\tj {:.3f} {:.3f} {:.3f} {:.3f}\tscl {:.3f}".format(
                            m["r2d"].avg, m["ar2d"].avg, m["d2r"].avg, m["ad2r"].avg, m["r2j"].avg, m["ar2j"].avg,
                            m["d2j"].avg, m["ad2j"].avg, m["scl"].avg))
                    print(msg)
                    sys.stdout.flush()
            end = time.time()
        print("epoch {}, total time {:.2f}".format(epoch, time.time() - t0))
        return meters["loss"].avg, meters["a12"].avg, 0.0, 0.0

    def train_step(self, model, contrast, optimizer, data, world=1):
        """One fused step on this rank's batch tuple (layout: hcmoco_b200/synthetic.py).  Returns a callable that
        reads the step's losses / accuracies back (one D2H copy) when called."""
        import torch.distributed as dist
        a = self.args
        x = data[0]
        eng = model.engine_for(x.shape[0], x.shape[-1], contrast)
        dev = eng.x.device
        if dev.type == "cuda" and not x.is_cuda and x.is_pinned():
            # pinned loader batches: the H2D copy runs on a copy stream while the previous step (still executing: the loop only
            # syncs every print_freq steps) computes; the step starts with a device-to-device copy
            from .pretrain import InputStager
            stager = self.stagers.get(id(eng))
            if stager is None:
                stager = self.stagers[id(eng)] = InputStager(eng.x)
            stager.consume(stager.stage(x), eng.x)
        else:
            eng.x.copy_(x, non_blocking=True)
        eng.index.copy_(data[1], non_blocking=True)
        eng.skel.copy_(data[2], non_blocking=True)
        eng.joints_yx.copy_(data[4], non_blocking=True)
        eng.joints_vis.copy_(data[5], non_blocking=True)
        if getattr(a, "modality_missing", 0):
            eng.use_depth.copy_(data[6], non_blocking=True)
        else:
            eng.use_depth.fill_(1)
        eng.depth_mask.copy_(data[7], non_blocking=True)
        eng.nce_idx.copy_(contrast.draw(eng.B, eng.index))
        if eng.stage == 2:
            eng.draw_dense(getattr(self, "injected_dense_idx", None))
        key = id(eng)
        if dev.type == "cuda" and getattr(a, "cuda_graph", True):
            if key not in self.graphs:
                eng.capture()
                self.graphs[key] = True
            eng.graph.replay()
        else:
            eng.forward()
            eng.backward()
        if world > 1:
            # synchronous collectives, one at a time (see pretrain.PretrainStep.run: an async all-reduce racing the gathers made
            # the replicas drift on 8 GPUs)
            all_f, all_y = self._global_gather(eng.f), self._global_gather(eng.index)
            dist.all_reduce(eng.store.g)
            eng.update_banks(all_f, all_y)
        else:
            eng.update_banks()
        optimizer.step(gscale=1.0 / world, from_autograd=False)
        return eng.results

    # ---- the three objectives with the reference signatures (contrast_trainer.py:642, 744, 830).  Inputs are the NCHW
    # tensors the reference passes; outputs are ([loss tensors], [accuracy tensors]) that back-propagate into feat_map1 /
    # feat_map2 / skeleton_map through the same loss kernels the fused step uses (hcmoco_b200/losses_api.py).
    def _compute_soft_pri3d_loss_accuracy(self, feat_map1, feat_map2, depth, criterion=None, use_depth=None, depth_mask=None,
                                          scale=None, sample_idx=None, K=None):
        from .losses_api import dense_loss
        return dense_loss(K or _kernels(), feat_map1, feat_map2, depth_mask, use_depth, float(self.args.temperature),
                          int(self.args.pri3d_num_samples_per_image), sample_idx)

    def _compute_joints_pri3d_loss_accuracy(self, feat_map1, feat_map2, skeleton_map, criterion=None, original_joints2d=None,
                                            joints_vis=None, use_depth=None, K=None):
        from .losses_api import joint_loss
        return joint_loss(K or _kernels(), feat_map1, feat_map2, skeleton_map, original_joints2d, joints_vis, use_depth,
                          float(self.args.temperature))

    def _compute_cross_subject_joints_pri3d_loss(self, feat_map1, feat_map2, skeleton_map=None, criterion=None,
                                                 original_joints2d=None, joints_vis=None, use_depth=None, index=None,
                                                 memory=None, use_rgb=None, K=None):
        from .losses_api import scl_loss
        return scl_loss(K or _kernels(), feat_map1, feat_map2, original_joints2d, use_depth, use_rgb,
                        float(self.args.temperature))


def build_contrast(opt):
    """Factory named by the north star (the reference has no `build_contrast`): the object that owns
    the three contrastive objectives and the step loops."""
    return ContrastTrainer(opt)
