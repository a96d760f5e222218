"""Segmentation fine-tuning on the pre-train engine.

Reference: pycontrast/main_segmentor.py:30-128, learning/segment_trainer.py:617-824 (`SegTrainer.train_soft_joint_pri3d`),
networks/fcn.py:35-111 (`FCNHead`), networks/build_linear.py:4-15 (`build_segmentor`: FCNHead(128, 128, n_class, num_convs=1,
kernel_size=1)).  The step is the second-stage pre-train step (identical encoders, identical four contrastive objectives: the
engine's launch programs, untouched) plus

    feat   = max(normalize(linear_merge1), normalize(linear_merge2))   on the samples that carry a label   (supervise_type 0)
    logits = upsample_x4(conv_seg(relu(bn(conv1x1(feat)))))                                               (FCNHead.forward)
    loss  += 10 * CrossEntropyLoss(ignore_index=255, weight=class_weights)(logits, label)

`SegHead` runs that head forward AND backward on the C-ABI kernels (hcm_l2norm_max_*, hcm_tc_conv / hcm_tc_wgrad for the 1x1
convolution, hcm_bn_*, hcm_gemm for the classifier, hcm_fuse_sum / hcm_upsample_adjoint for the x4 resize, hcm_seg_ce_*) on the
engine's channels-last maps; its gradient w.r.t. the two projection maps is added to the gradients the contrastive objectives left
in the engine's map-gradient buffers before the model part of the backward program runs.  `FCNHead` is the nn.Module face with the
reference's state_dict keys; `SegTrainer` mirrors the reference trainer's public methods on this path.
There is no CPU / PyTorch fallback: the kernels object is `CudaKernels` unless a test injects its reference executor."""
import os
import sys
import time
from collections import OrderedDict

import torch
import torch.nn as nn

from .api import AverageMeter, ContrastTrainer, _kernels
from .engine import ParamStore

FCN_BN_MOMENTUM, FCN_BN_EPS = 0.1, 1e-5            # nn.BatchNorm2d defaults (fcn.py:9)
SEG_LOSS_WEIGHT = 10.0                             # segment_trainer.py:745 `loss += loss_seg * 10`
# main_segmentor.py:76: class weights of the 25 NTU body-part classes
NTU_CLASS_WEIGHTS = [1.448, 49.234, 49.483, 48.030, 49.247, 49.492, 48.018, 49.704, 50.052, 49.369, 49.694, 50.090, 49.425, 49.459,
                     45.846, 47.156, 45.868, 47.197, 44.167, 42.789, 44.341, 48.632, 48.873, 48.644, 49.004]


def fcn_keys(n_class, channels=128):
    """state_dict of FCNHead(channels, channels, n_class, num_convs=1, kernel_size=1), in the reference's order (ConvModule registers
    its norm before its conv, fcn.py:9-22)."""
    c = channels
    return OrderedDict([
        ("convs.0.norm_name.weight", (c,)), ("convs.0.norm_name.bias", (c,)), ("convs.0.norm_name.running_mean", (c,)),
        ("convs.0.norm_name.running_var", (c,)), ("convs.0.norm_name.num_batches_tracked", ()),
        ("convs.0.conv.weight", (c, c, 1, 1)), ("convs.0.conv.bias", (c,)),
        ("conv_seg.weight", (n_class, c, 1, 1)), ("conv_seg.bias", (n_class,)),
    ])


class SegHead:
    """FCNHead forward + weighted CE + backward as C-ABI launches on channels-last maps.  Parameters / gradients / momentum live in a
    flat `ParamStore` (checkpoint layout), exactly as the encoder's."""

    def __init__(self, K, n_class=25, channels=128, class_weights=None, ignore_index=255, seed=0):
        assert channels == 128, "the projection maps have 128 channels (build_linear.py:7)"
        self.K, self.C, self.Cn, self.ignore = K, channels, int(n_class), int(ignore_index)
        self.store = ParamStore(K, fcn_keys(self.Cn, channels))
        self.cw = None if class_weights is None else torch.as_tensor(class_weights, dtype=torch.float32).to(K.device, K.dtype)
        self.first_step = True
        self._init(seed)

    def _init(self, seed):
        """nn.Conv2d / nn.BatchNorm2d default initialisers (kaiming_uniform(a=sqrt(5)) = U(+-1/sqrt(fan_in)) for weight and bias)."""
        g = torch.Generator().manual_seed(seed)
        st = self.store
        bound = 1.0 / (self.C ** 0.5)
        for k, shp in st.keys.items():
            if k.endswith("norm_name.weight"):
                st.load(st.p, k, torch.ones(shp))
            elif k.endswith("norm_name.bias"):
                st.load(st.p, k, torch.zeros(shp))
            elif k.endswith(("conv.weight", "conv.bias", "conv_seg.weight", "conv_seg.bias")):
                st.load(st.p, k, (torch.rand(shp, generator=g) * 2 - 1) * bound)

    # -------------------------------------------------------------------------------------------- forward pieces
    def _features(self, m1, m2, supervise_type):
        """segment_trainer.py:722-741: the classifier's input and what the backward of it needs."""
        K, n, h = self.K, m1.shape[0], m1.shape[1]
        P = n * h * h
        a, b = (m1, m2) if supervise_type == 0 else ((m1, None) if supervise_type == 1 else (m2, None))
        feat, inv1 = K.empty(n, h, h, 128), K.empty(P)
        inv2 = K.empty(P) if b is not None else None
        K.l2norm_max_fwd(a, b, P, 128, feat, inv1, inv2)
        return feat, (a, b, inv1, inv2)

    def _fcn_forward(self, feat, update_stats=True):
        """fcn.py:104-110 (train mode): 1x1 conv + bias -> BN -> ReLU -> conv_seg -> bilinear x4.  Returns the upsampled logits
        [n,4h,4h,Cn] (channels-last) and the saved activations."""
        K, st, C, Cn = self.K, self.store, self.C, self.Cn
        n, h = feat.shape[0], feat.shape[1]
        P, R = n * h * h, 4 * h
        if not K.tc_conv_supported(n, h, h, C, C, 1, 1):
            raise NotImplementedError("FCNHead 1x1 convolution: unsupported geometry")
        w1, b1 = st.param("convs.0.conv.weight"), st.param("convs.0.conv.bias")
        wp = K.empty((K.tc_conv_wpack_bytes(n, h, h, C, C, 1) + 3) // 4)
        K.tc_conv_pack(w1, 0, wp, n, h, h, C, C, 1, 0)
        y1 = K.empty(n, h, h, C)
        K.tc_conv(feat, wp, b1, y1, n, h, h, C, C, 1, 1, None, None, 0, 0)
        nparts = K.colstat_rows(P, C)
        part = K.empty(max(nparts, K.colstat_rows(P, Cn)) * 2 * C)
        scale, shift, mean, invstd = K.empty(C), K.empty(C), K.empty(C), K.empty(C)
        bf = st.buffers
        K.bn_stats(y1, P, C, part)
        bk = "convs.0.norm_name."
        if update_stats:
            K.bn_finalize(part, nparts, C, P, st.param(bk + "weight"), st.param(bk + "bias"), bf[bk + "running_mean"],
                          bf[bk + "running_var"], bf[bk + "num_batches_tracked"], FCN_BN_MOMENTUM, FCN_BN_EPS, scale, shift, mean, invstd)
        else:
            K.bn_finalize(part, nparts, C, P, st.param(bk + "weight"), st.param(bk + "bias"), None, None, None, FCN_BN_MOMENTUM,
                          FCN_BN_EPS, scale, shift, mean, invstd)
        z = K.empty(n, h, h, C)
        K.bn_apply(y1, scale, shift, None, None, None, 1, z, P, C)
        logits = K.empty(n, h, h, Cn)
        # logits[p][c] = sum_k z[p][k] * Wseg[c][k] + b[c]
        K.gemm(z, st.param("conv_seg.weight"), st.param("conv_seg.bias"), logits, 1, P, Cn, C, C, 1, 1, C, Cn, 0, 0, 0, 1.0, 0)
        up = K.empty(n, R, R, Cn)
        K.fuse_sum(1, [logits], None, None, torch.tensor([2], dtype=torch.int32), None, 0, up, n, R, R, Cn)
        return up, dict(feat=feat, y1=y1, z=z, part=part, scale=scale, shift=shift, mean=mean, invstd=invstd, n=n, h=h)

    def _fcn_backward(self, g_up, sv):
        """Backward of `_fcn_forward`: parameter gradients are ADDED into store.g (zeroed by the caller once per step); returns
        d(loss)/d(feat) [n,h,h,128]."""
        K, st, C, Cn = self.K, self.store, self.C, self.Cn
        n, h = sv["n"], sv["h"]
        P, R = n * h * h, 4 * h
        feat, y1, z, part = sv["feat"], sv["y1"], sv["z"], sv["part"]
        dlog = K.empty(n, h, h, Cn)
        K.upsample_adjoint(g_up, dlog, 0, n, R, R, Cn, 2)
        # conv_seg: dW[c][k] += sum_p dlog[p][c] z[p][k] (split over nb position blocks, then a column sum); db; dz
        nb = 1
        while nb < 256 and P % (2 * nb) == 0 and P // (2 * nb) >= 256:
            nb *= 2
        Kc = P // nb
        parts = K.empty(nb, Cn * C)
        K.gemm(dlog, z, None, parts, nb, Cn, C, Kc, 1, Cn, C, 1, C, Kc * Cn, Kc * C, Cn * C, 1.0, 0)
        K.colsum_small(parts, nb, Cn * C, Cn * C, st.grad("conv_seg.weight"), 1)
        K.bn_stats(dlog, P, Cn, part)
        K.colsum_finalize(part, K.colstat_rows(P, Cn), Cn, st.grad("conv_seg.bias"), 1)
        dz = K.empty(n, h, h, C)
        K.gemm(dlog, st.param("conv_seg.weight"), None, dz, 1, P, C, Cn, Cn, 1, C, 1, C, 0, 0, 0, 1.0, 0)
        # BN (+ReLU: mask recomputed from the raw conv output) backward
        bk = "convs.0.norm_name."
        nparts = K.colstat_rows(P, C)
        k1, k2, k3 = K.empty(C), K.empty(C), K.empty(C)
        dg, db = K.empty(C), K.empty(C)
        K.bn_bwd_reduce(dz, None, sv["scale"], sv["shift"], y1, sv["mean"], sv["invstd"], P, C, part)
        K.bn_bwd_finalize(part, nparts, C, P, st.param(bk + "weight"), sv["mean"], sv["invstd"], dg, db, k1, k2, k3)
        K.axpy(st.grad(bk + "weight"), dg, 1.0, C)
        K.axpy(st.grad(bk + "bias"), db, 1.0, C)
        dy1 = K.empty(n, h, h, C)
        K.bn_bwd_apply(dz, None, sv["scale"], sv["shift"], y1, k1, k2, k3, dy1, None, 0, P, C)
        # 1x1 conv: bias, weight and data gradients
        K.bn_stats(dy1, P, C, part)
        K.colsum_finalize(part, nparts, C, st.grad("convs.0.conv.bias"), 1)
        K.tc_wgrad(feat, dy1, st.grad("convs.0.conv.weight"), 0, n, h, h, C, C, 1, 1, None, None, 0)
        wpt = K.empty((K.tc_conv_wpack_bytes(n, h, h, C, C, 1) + 3) // 4)
        K.tc_conv_pack(st.param("convs.0.conv.weight"), 0, wpt, n, h, h, C, C, 1, 1)
        dfeat = K.empty(n, h, h, C)
        K.tc_conv(dy1, wpt, None, dfeat, n, h, h, C, C, 1, 1, None, None, 0, 0)
        return dfeat

    # -------------------------------------------------------------------------------------------- the fused loss
    def loss_backward(self, m1, m2, label, supervise_type=0, gscale=SEG_LOSS_WEIGHT):
        """m1, m2 [n,h,h,128] channels-last projection maps of the labelled samples, label [n,4h,4h] int64.
        Returns (out2 = device tensor (loss_seg, aAcc), d m1, d m2) with d = gscale * d(loss_seg)/d(map); parameter gradients
        (x gscale) are added into store.g."""
        K, Cn = self.K, self.Cn
        n, h = m1.shape[0], m1.shape[1]
        R = 4 * h
        feat, fsv = self._features(m1, m2, supervise_type)
        up, sv = self._fcn_forward(feat)
        lab = label.to(K.device).long().contiguous()
        acc, out2 = K.zeros(4, dtype=torch.float64), K.empty(2)
        K.seg_ce_fwd(up, lab, self.cw, n * R * R, Cn, self.ignore, acc, out2)
        g_up = K.empty(n, R, R, Cn)
        K.seg_ce_bwd(up, lab, self.cw, n * R * R, Cn, self.ignore, acc, float(gscale), g_up)
        dfeat = self._fcn_backward(g_up, sv)
        a, b, inv1, inv2 = fsv
        da = K.empty(n, h, h, 128)
        db = K.empty(n, h, h, 128) if b is not None else None
        K.l2norm_max_bwd(dfeat, a, b, inv1, inv2, n * h * h, 128, 1.0, da, db, 0)
        if supervise_type == 0:
            return out2, da, db
        return (out2, da, None) if supervise_type == 1 else (out2, None, da)

    def forward_only(self, m):
        """`tmp = classifier(linear_merge1); loss += (tmp - tmp).mean()` (segment_trainer.py:742-748): no gradient, but the train-mode
        forward updates the BatchNorm running statistics."""
        self._fcn_forward(m)

    def sgd(self, lr, momentum, wd, gscale=1.0):
        st = self.store
        self.K.sgd_step(st.p, st.g, st.m, st.n, lr, momentum, wd, 1 if self.first_step else 0, gscale)
        self.first_step = False

    def zero_grad(self):
        st = self.store
        self.K.zero(st.g, st.n * st.g.element_size())


class _FcnFn(torch.autograd.Function):
    """`classifier(x)` of the reference (NCHW in, upsampled NCHW logits out), differentiable."""

    @staticmethod
    def forward(ctx, head, x, *params):
        K = head.K
        feat = x.detach().to(K.device, K.dtype).permute(0, 2, 3, 1).contiguous()
        up, sv = head._fcn_forward(feat)
        ctx.head, ctx.sv = head, sv
        return up.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        head = ctx.head
        st = head.store
        head.zero_grad()
        dfeat = head._fcn_backward(g.permute(0, 2, 3, 1).contiguous(), ctx.sv)
        flat = st.g.clone()
        grads = tuple(st.view(flat, k).view(st.keys[k]) for k in head.param_keys)
        return (None, dfeat.permute(0, 3, 1, 2)) + grads


class FCNHead(nn.Module):
    """networks/fcn.py:35-111 as built by build_segmentor: same state_dict keys / shapes; parameters alias the head's flat store."""

    def __init__(self, in_channels=128, channels=128, num_classes=25, num_convs=1, kernel_size=1, kernels=None, class_weights=None):
        super().__init__()
        assert in_channels == 128 and channels == 128 and num_convs == 1 and kernel_size == 1, \
            "build_linear.py:8-14 builds FCNHead(128, 128, n_class, num_convs=1, kernel_size=1)"
        self.head = SegHead(_kernels(kernels), num_classes, channels, class_weights)
        st = self.head.store
        self.head.param_keys = []
        from .api import _Node
        for k, shp in st.keys.items():
            node, parts = self, k.split(".")
            for a in parts[:-1]:
                if not hasattr(node, a):
                    node.add_module(a, _Node())
                node = getattr(node, a)
            if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
                node.register_buffer(parts[-1], st.buffers[k])
            else:
                node.register_parameter(parts[-1], nn.Parameter(st.view(st.p, k).view(shp)))
                self.head.param_keys.append(k)

    def cuda(self, device=None):
        return self

    def forward(self, x):
        return _FcnFn.apply(self.head, x, *list(self.parameters()))


def build_segmentor(opt, kernels=None):
    """networks/build_linear.py:4-15."""
    return FCNHead(128, 128, int(opt.n_class), 1, 1, kernels, getattr(opt, "class_weights", NTU_CLASS_WEIGHTS
                                                                         if int(opt.n_class) == 25 else None))


class SegTrainer(ContrastTrainer):
    """learning/segment_trainer.py on the RGBD2S / `bank+jointspri3d` path: `train_soft_joint_pri3d` and what main_segmentor.py calls
    around it."""

    def wrap_up(self, model, classifier, optimizer=None):       # segment_trainer.py:89-102 (no DDP wrappers: flat gradient buffers)
        return model, classifier

    def resume_model(self, model, contrast, classifier, optimizer):          # :171-190
        a, start_epoch = self.args, 1
        if getattr(a, "resume", "") and os.path.isfile(a.resume):
            ck = torch.load(a.resume, map_location="cpu")
            start_epoch = ck["epoch"] + 1
            model.store.load_state_dict(ck["model"])
            contrast.load_state_dict(ck["contrast"])
            classifier.head.store.load_state_dict(ck["classifier"])
        return start_epoch

    def save(self, model, contrast, classifier, optimizer, epoch):            # :192-212
        a = self.args
        if getattr(a, "local_rank", 0) == 0:
            state = {"model": model.store.state_dict(prefix="module."), "contrast": contrast.state_dict(),
                     "classifier": classifier.head.store.state_dict(prefix="module."), "epoch": epoch}
            torch.save(state, os.path.join(a.model_folder, "current.pth"))

    @staticmethod
    def eval_seg_aacc(logits, target):                                        # :375-379
        return (logits.argmax(1) == target).sum().float() / float(target.numel())

    def seg_step(self, model, classifier, contrast, data, lr, momentum, wd, world=1):
        """One fused fine-tuning step on this rank's batch tuple (`data[9]` label [B,R,R], `data[10]` true_label [B]; the other
        fields as in the pre-train step).  Returns a callable that reads the losses back."""
        import torch.distributed as dist
        a = self.args
        assert float(getattr(a, "cmc_loss_weights", 1)) == 1.0 and float(getattr(a, "other_loss_weights", 1)) == 1.0, \
            "the fused step implements the shipped weights (base_options.py:100-101: 1 / 1)"
        head = classifier.head if hasattr(classifier, "head") else classifier
        K = model.K
        x = data[0]
        eng = model.engine_for(x.shape[0], x.shape[-1], contrast)
        assert eng.stage == 2, "train_soft_joint_pri3d runs on the second-stage model (--linear_feat_map 1)"
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
        eng.draw_dense(getattr(self, "injected_dense_idx", None))
        label = data[9].to(K.device).long()
        sel = torch.nonzero(data[10].to(K.device) != 0).reshape(-1)
        st_type = int(getattr(a, "supervise_type", 0))
        # forward: encoders + the four contrastive objectives (the engine's program), then the head
        graphs = None
        if eng.x.is_cuda and getattr(a, "cuda_graph", True):
            if not hasattr(eng, "graph_parts"):
                eng.capture_parts()
            graphs = eng.graph_parts
        if graphs is not None:
            graphs[0].replay()
        else:
            eng.forward()
        head.zero_grad()
        out2 = None
        if sel.numel() > 0 and st_type in (0, 1, 2):
            m1, m2 = eng.lm1.data.index_select(0, sel), eng.lm2.data.index_select(0, sel)
            out2, d1, d2 = head.loss_backward(m1, m2, label.index_select(0, sel), st_type, SEG_LOSS_WEIGHT)
        else:
            head.forward_only(eng.lm1.data)
            d1 = d2 = None
        # backward: loss part (writes the map gradients), + the head's, then the model part
        if graphs is not None:
            graphs[1].replay()
        else:
            K.zero(eng.store.g, eng.store.n * eng.store.g.element_size())
            eng.plan.run(eng.plan.bwd[:eng.n_loss_bwd], eng.two_streams)
        for act, d in ((eng.lm1, d1), (eng.lm2, d2)):
            if d is not None:
                act.grad.index_add_(0, sel, d)
        if graphs is not None:
            graphs[2].replay()
        else:
            eng.plan.run(eng.plan.bwd[eng.n_loss_bwd:], eng.two_streams)
        if world > 1:
            # synchronous collectives, one at a time (see pretrain.PretrainStep.run)
            all_f, all_y = self._global_gather(eng.f), self._global_gather(eng.index)
            dist.all_reduce(eng.store.g)
            dist.all_reduce(head.store.g)
            eng.update_banks(all_f, all_y)
        else:
            eng.update_banks()
        eng.sgd(lr, momentum, wd, 1.0 / world)
        head.sgd(lr, momentum, wd, 1.0 / world)
        res = eng.results

        def read():
            r = res()
            seg = out2.detach().cpu() if out2 is not None else torch.zeros(2)
            r["seg_loss"], r["seg_aacc"] = seg[0], seg[1]
            r["loss"] = r["loss"] + SEG_LOSS_WEIGHT * float(seg[0])
            return r
        return read

    def train_soft_joint_pri3d(self, epoch, train_loader, model, classifier, contrast, criterion_contrast=None, criterion_pri3d=None,
                               criterion_seg=None, optimizer=None):
        """segment_trainer.py:617-824.  `optimizer` supplies lr / momentum / weight_decay (its param_groups[0]); the update itself is
        the fused SGD kernel over the two flat stores."""
        a = self.args
        model.attach_memory(contrast)
        meters = {k: AverageMeter() for k in ("bt", "loss", "seg", "aacc")}
        world = getattr(a, "world_size", 1) or 1
        n_batches = len(train_loader) if hasattr(train_loader, "__len__") else 0
        end = time.time()
        for idx, data in enumerate(train_loader):
            if optimizer is not None:
                self.warmup_learning_rate(epoch, idx, n_batches, optimizer)
                g = optimizer.param_groups[0]
                lr, mom, wd = g["lr"], g.get("momentum", 0.0), g.get("weight_decay", 0.0)
            else:
                lr, mom, wd = a.learning_rate, a.momentum, a.weight_decay
            res = self.seg_step(model, classifier, contrast, data, lr, mom, wd, world)
            if (idx + 1) % a.print_freq == 0 or idx + 1 == n_batches:
                r = res()
                bsz = data[0].shape[0]
                meters["loss"].update(float(r["loss"]), bsz)
                if int(torch.as_tensor(data[10]).sum()) != 0:
                    meters["seg"].update(float(r["seg_loss"]), bsz)
                    meters["aacc"].update(float(r["seg_aacc"]), bsz)
                meters["bt"].update(time.time() - end)
                if getattr(a, "local_rank", 0) == 0:
                    print("Train: [{0}][{1}/{2}] BT {3:.3f} L {4:.3f} ({5:.3f}) seg {6:.3f} {7:.3f}".format(
                        epoch, idx + 1, n_batches, meters["bt"].val, meters["loss"].val, meters["loss"].avg, meters["seg"].avg,
                        meters["aacc"].avg))
                    sys.stdout.flush()
            end = time.time()
        return meters["seg"].avg, meters["aacc"].avg
