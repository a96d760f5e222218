"""CPU oracle for the HCMoCo contrastive pre-train step.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch, functional restatement (plain fp32 PyTorch on CPU, autograd for
the backward) of the algorithm in the reference's hot path, written against the reference's
*checkpoint key layout* instead of its nn.Module classes.  Every function cites the reference
lines it follows (paths relative to /root/reference/pycontrast).

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the reference itself, executed in the build container under
tests/golden/ref_shim.py by tests/golden/make_golden.py; the resulting fixtures live in
tests/golden/*.pt and tests/test_oracle_golden.py checks this file against them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under hcmoco_b200/ does: the product path is CUDA-only and fails
loudly without its extension.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

BN_MOMENTUM_2D = 0.01   # networks/official_hrnet/official_hrnet.py:22-23
BN_MOMENTUM_1D = 0.1    # nn.BatchNorm1d default, networks/SGCN/sem_gcn.py:13
BN_EPS = 1e-5

# networks/official_hrnet/seg_hrnet_w{18,32,48}*.yaml: NUM_CHANNELS of stage 4; stages 2,3 are prefixes
HRNET_WIDTHS = {18: (18, 36, 72, 144), 32: (32, 64, 128, 256), 48: (48, 96, 192, 384)}
# (num_modules, num_branches) of stage2..4; 4 BasicBlocks per branch everywhere
HRNET_STAGES = ((1, 2), (4, 3), (3, 4))

# networks/SGCN/skeleton_meta.py:3-23
SKELETON_PARENTS = {
    "mpii": [1, 2, 6, 6, 3, 4, -1, 6, 7, 8, 11, 12, 8, 8, 13, 14],
    "coco_reduce": [1, 2, 9, 10, 3, 4, -1, 8, 9, 6, 6, 10, 11],
}


# --------------------------------------------------------------------------------------
# checkpoint layout: ordered key -> shape, identical to the reference's
# model.state_dict() for modal=RGBD2S, arch=HRNet
# --------------------------------------------------------------------------------------
def _bn_keys(out, pre, c):
    out[pre + ".weight"] = (c,)
    out[pre + ".bias"] = (c,)
    out[pre + ".running_mean"] = (c,)
    out[pre + ".running_var"] = (c,)
    out[pre + ".num_batches_tracked"] = ()


def hrnet_layout(pre, width):
    """Key layout of one HighResolutionNet (official_hrnet.py:258-327)."""
    C = HRNET_WIDTHS[width]
    o = OrderedDict()
    o[pre + "conv1.weight"] = (64, 3, 3, 3)
    _bn_keys(o, pre + "bn1", 64)
    o[pre + "conv2.weight"] = (64, 64, 3, 3)
    _bn_keys(o, pre + "bn2", 64)
    for b in range(4):  # layer1: 4 Bottlenecks (official_hrnet.py:64-102, 365-380)
        p = "%slayer1.%d." % (pre, b)
        cin = 64 if b == 0 else 256
        o[p + "conv1.weight"] = (64, cin, 1, 1)
        _bn_keys(o, p + "bn1", 64)
        o[p + "conv2.weight"] = (64, 64, 3, 3)
        _bn_keys(o, p + "bn2", 64)
        o[p + "conv3.weight"] = (256, 64, 1, 1)
        _bn_keys(o, p + "bn3", 256)
        if b == 0:
            o[p + "downsample.0.weight"] = (256, 64, 1, 1)
            _bn_keys(o, p + "downsample.1", 256)
    prev = [256]
    for si, (nmod, nbr) in enumerate(HRNET_STAGES):
        cur = list(C[:nbr])
        # transition (official_hrnet.py:329-363)
        tp = "%stransition%d." % (pre, si + 1)
        for i in range(nbr):
            if i < len(prev):
                if cur[i] != prev[i]:
                    o["%s%d.0.weight" % (tp, i)] = (cur[i], prev[i], 3, 3)
                    _bn_keys(o, "%s%d.1" % (tp, i), cur[i])
            else:
                for j in range(i + 1 - len(prev)):
                    cout = cur[i] if j == i - len(prev) else prev[-1]
                    o["%s%d.%d.0.weight" % (tp, i, j)] = (cout, prev[-1], 3, 3)
                    _bn_keys(o, "%s%d.%d.1" % (tp, i, j), cout)
        # stage modules (official_hrnet.py:105-249)
        for m in range(nmod):
            mp = "%sstage%d.%d." % (pre, si + 2, m)
            for i in range(nbr):
                for b in range(4):
                    bp = "%sbranches.%d.%d." % (mp, i, b)
                    o[bp + "conv1.weight"] = (cur[i], cur[i], 3, 3)
                    _bn_keys(o, bp + "bn1", cur[i])
                    o[bp + "conv2.weight"] = (cur[i], cur[i], 3, 3)
                    _bn_keys(o, bp + "bn2", cur[i])
            for i in range(nbr):
                for j in range(nbr):
                    fp = "%sfuse_layers.%d.%d." % (mp, i, j)
                    if j > i:
                        o[fp + "0.weight"] = (cur[i], cur[j], 1, 1)
                        _bn_keys(o, fp + "1", cur[i])
                    elif j < i:
                        for k in range(i - j):
                            cout = cur[i] if k == i - j - 1 else cur[j]
                            o["%s%d.0.weight" % (fp, k)] = (cout, cur[j], 3, 3)
                            _bn_keys(o, "%s%d.1" % (fp, k), cout)
        prev = cur
    return o


def skeleton_edges(name):
    """Row-major list of (row, col) non-zeros of the SemGCN adjacency (graph_utils.py:27-45)."""
    parents = SKELETON_PARENTS[name]
    J = len(parents)
    nz = set((i, i) for i in range(J))
    for i, p in enumerate(parents):
        if p >= 0:
            nz.add((i, p))
            nz.add((p, i))
    return J, sorted(nz)


def sgcn_layout(pre, skeleton, hid=128):
    """Key layout of SemGCN (sem_gcn.py:60-89, sem_graph_conv.py:14-32)."""
    _, nz = skeleton_edges(skeleton)
    o = OrderedDict()

    def gconv(p, cin, cout):
        o[p + ".W"] = (2, cin, cout)
        o[p + ".e"] = (1, len(nz))
        o[p + ".bias"] = (cout,)

    gconv(pre + "gconv_input.0.gconv", 2, hid)
    _bn_keys(o, pre + "gconv_input.0.bn", hid)
    for l in range(4):
        for g in (1, 2):
            gconv("%sgconv_layers.%d.gconv%d.gconv" % (pre, l, g), hid, hid)
            _bn_keys(o, "%sgconv_layers.%d.gconv%d.bn" % (pre, l, g), hid)
    gconv(pre + "gconv_output", hid, hid)
    return o


def model_layout(width=18, stage=1, skeleton="mpii", feat_dim=128):
    """state_dict layout of CMC3HRNetSGCNSingleHead (build_backbone.py:186-245)."""
    cm = sum(HRNET_WIDTHS[width])
    o = OrderedDict()
    o.update(hrnet_layout("encoder1.", width))
    o.update(hrnet_layout("encoder2.", width))
    o.update(sgcn_layout("encoder3.", skeleton))
    for i, cin in ((1, cm), (2, cm), (3, 128)):
        o["head%d.0.weight" % i] = (feat_dim, cin)
        o["head%d.0.bias" % i] = (feat_dim,)
    if stage == 2:
        for i in (1, 2):
            o["encoder%d_linear.weight" % i] = (128, cm, 1, 1)
            o["encoder%d_linear.bias" % i] = (128,)
    return o


def is_param(key):
    return not (key.endswith("running_mean") or key.endswith("running_var")
                or key.endswith("num_batches_tracked"))


# --------------------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------------------
class _Ctx:
    """Parameter dict + train flag; BN running statistics are updated in place like nn.BatchNorm."""

    def __init__(self, P, train=True):
        self.P = P
        self.train = train

    def conv(self, x, key, stride=1):
        w = self.P[key + ".weight"]
        return F.conv2d(x, w, None, stride, (w.shape[-1] - 1) // 2)

    def bn(self, x, key, momentum=BN_MOMENTUM_2D):
        P = self.P
        if self.train:
            P[key + ".num_batches_tracked"] += 1
        return F.batch_norm(x, P[key + ".running_mean"], P[key + ".running_var"],
                            P[key + ".weight"], P[key + ".bias"], self.train, momentum, BN_EPS)

    def cb(self, x, ck, bk, stride=1, relu=False):
        y = self.bn(self.conv(x, ck, stride), bk)
        return F.relu(y) if relu else y


def _basic_block(c, x, p):
    # official_hrnet.py:32-61
    y = c.cb(x, p + "conv1", p + "bn1", relu=True)
    y = c.cb(y, p + "conv2", p + "bn2")
    return F.relu(y + x)


def _bottleneck(c, x, p, has_down):
    # official_hrnet.py:64-102
    y = c.cb(x, p + "conv1", p + "bn1", relu=True)
    y = c.cb(y, p + "conv2", p + "bn2", relu=True)
    y = c.cb(y, p + "conv3", p + "bn3")
    r = c.cb(x, p + "downsample.0", p + "downsample.1") if has_down else x
    return F.relu(y + r)


def _hr_module(c, xs, mp):
    # official_hrnet.py:225-249 (+ fuse layer construction :176-220)
    n = len(xs)
    xs = list(xs)
    for i in range(n):
        for b in range(4):
            xs[i] = _basic_block(c, xs[i], "%sbranches.%d.%d." % (mp, i, b))
    out = []
    for i in range(n):
        acc = None
        for j in range(n):
            fp = "%sfuse_layers.%d.%d." % (mp, i, j)
            if j == i:
                t = xs[j]
            elif j > i:
                t = c.cb(xs[j], fp + "0", fp + "1")
                t = F.interpolate(t, size=xs[i].shape[-2:], mode="bilinear", align_corners=False)
            else:
                t = xs[j]
                for k in range(i - j):
                    t = c.cb(t, "%s%d.0" % (fp, k), "%s%d.1" % (fp, k), stride=2,
                             relu=(k != i - j - 1))
            acc = t if acc is None else acc + t
        out.append(F.relu(acc))
    return out


def hrnet_forward(P, pre, x, width=18, train=True):
    """HighResolutionNet.forward (official_hrnet.py:411-454): returns the 4 branch maps."""
    c = _Ctx(P, train)
    x = c.cb(x, pre + "conv1", pre + "bn1", stride=2, relu=True)
    x = c.cb(x, pre + "conv2", pre + "bn2", stride=2, relu=True)
    for b in range(4):
        x = _bottleneck(c, x, "%slayer1.%d." % (pre, b), b == 0)
    ys = [x]
    nprev = 1
    for si, (nmod, nbr) in enumerate(HRNET_STAGES):
        tp = "%stransition%d." % (pre, si + 1)
        xs = []
        for i in range(nbr):
            if i < nprev:
                if (tp + "%d.0.weight" % i) in P:
                    xs.append(c.cb(ys[i], "%s%d.0" % (tp, i), "%s%d.1" % (tp, i), relu=True))
                else:
                    xs.append(ys[i])
            else:
                t = ys[-1]
                for j in range(i + 1 - nprev):
                    t = c.cb(t, "%s%d.%d.0" % (tp, i, j), "%s%d.%d.1" % (tp, i, j), stride=2, relu=True)
                xs.append(t)
        for m in range(nmod):
            xs = _hr_module(c, xs, "%sstage%d.%d." % (pre, si + 2, m))
        ys = xs
        nprev = nbr
    return ys


def sgcn_adjacency_mask(skeleton):
    J, nz = skeleton_edges(skeleton)
    m = torch.zeros(J, J, dtype=torch.bool)
    for r, cc in nz:
        m[r, cc] = True
    return m


def _sem_gconv(P, p, x, mask):
    # sem_graph_conv.py:34-48
    W, e, b = P[p + ".W"], P[p + ".e"], P[p + ".bias"]
    h0 = x @ W[0]
    h1 = x @ W[1]
    J = mask.shape[0]
    mask = mask.to(x.device)
    rows, cols = mask.nonzero(as_tuple=True)
    A = torch.full((J, J), -9e15, dtype=x.dtype, device=x.device).index_put((rows, cols), e.reshape(-1))
    A = torch.softmax(A, dim=1)
    eye = torch.eye(J, dtype=x.dtype, device=x.device)
    return (A * eye) @ h0 + (A * (1 - eye)) @ h1 + b.view(1, 1, -1)


def _graph_conv(P, p, x, mask, train):
    # sem_gcn.py:8-28 : gconv -> BatchNorm1d over channels -> ReLU
    y = _sem_gconv(P, p + ".gconv", x, mask).transpose(1, 2)
    if train:
        P[p + ".bn.num_batches_tracked"] += 1
    y = F.batch_norm(y, P[p + ".bn.running_mean"], P[p + ".bn.running_var"], P[p + ".bn.weight"],
                     P[p + ".bn.bias"], train, BN_MOMENTUM_1D, BN_EPS)
    return F.relu(y.transpose(1, 2))


def sgcn_forward(P, pre, s, skeleton="mpii", train=True):
    """SemGCN.forward (sem_gcn.py:91-95) for create_sgcn(name,128,4) (create_SGCN.py:6-14)."""
    mask = sgcn_adjacency_mask(skeleton)
    x = _graph_conv(P, pre + "gconv_input.0", s, mask, train)
    for l in range(4):
        lp = "%sgconv_layers.%d." % (pre, l)
        y = _graph_conv(P, lp + "gconv1", x, mask, train)
        y = _graph_conv(P, lp + "gconv2", y, mask, train)
        x = x + y
    return _sem_gconv(P, pre + "gconv_output", x, mask)


def merge_all_res(feats):
    # build_backbone.py:247-254
    size = feats[0].shape[-2:]
    ups = [feats[0]] + [F.interpolate(f, size=size, mode="bilinear", align_corners=False)
                        for f in feats[1:]]
    return torch.cat(ups, 1)


def model_forward(P, x, s, width=18, skeleton="mpii", stage=1, train=True):
    """CMC3HRNetSGCNSingleHead.forward (build_backbone.py:256-303), mode=0.

    Returns dict(f [B,384], feat1, feat2 (lists of 4), feat3 [B,J,128], and for stage 2
    linear_merge1/2 [B,128,h,h]).
    """
    x1, x2 = x[:, :3], x[:, 3:6]
    feat1 = hrnet_forward(P, "encoder1.", x1, width, train)
    feat2 = hrnet_forward(P, "encoder2.", x2, width, train)
    feat3 = sgcn_forward(P, "encoder3.", s, skeleton, train)
    a1 = torch.cat([f.mean((2, 3)) for f in feat1], 1)
    a2 = torch.cat([f.mean((2, 3)) for f in feat2], 1)
    a3 = feat3.mean(1)
    f1 = F.normalize(F.linear(a1, P["head1.0.weight"], P["head1.0.bias"]), dim=1)
    f2 = F.normalize(F.linear(a2, P["head2.0.weight"], P["head2.0.bias"]), dim=1)
    f3 = F.normalize(F.linear(a3, P["head3.0.weight"], P["head3.0.bias"]), dim=1)
    out = dict(f=torch.cat((f1, f2, f3), 1), feat1=feat1, feat2=feat2, feat3=feat3)
    if stage == 2:
        out["linear_merge1"] = F.conv2d(merge_all_res(feat1), P["encoder1_linear.weight"],
                                        P["encoder1_linear.bias"])
        out["linear_merge2"] = F.conv2d(merge_all_res(feat2), P["encoder2_linear.weight"],
                                        P["encoder2_linear.bias"])
    return out


# --------------------------------------------------------------------------------------
# memory bank NCE (memory/mem_bank.py)
# --------------------------------------------------------------------------------------
NCE_PAIRS = ((0, 1), (1, 0), (1, 2), (2, 1), (0, 2), (2, 0))  # (query modality, bank modality)


def nce_logits(banks, xs, idx, T=0.07):
    """CMCMem3.forward logits (mem_bank.py:179-191): six [B,K+1] tensors, order 12,21,23,32,13,31.

    idx [B,K+1] int64 with idx[:,0] == y (mem_bank.py:176-177).
    """
    B, K1 = idx.shape
    w = [bk.index_select(0, idx.reshape(-1)).view(B, K1, -1) for bk in banks]
    return [torch.bmm(w[q], xs[p].unsqueeze(2)).squeeze(2) / T for p, q in NCE_PAIRS]


def bank_update(bank, x, y, m=0.5):
    """BaseMem._update_memory (mem_bank.py:15-28); duplicates in y: every duplicate reads the old
    row and the last writer wins (index_copy_ on CPU)."""
    with torch.no_grad():
        w = bank.index_select(0, y.view(-1)) * m + x.detach() * (1 - m)
        bank.index_copy_(0, y, F.normalize(w))


def _top1(logit):
    # learning/util.py:24-38 with target 0: fraction (in %) of rows whose arg-max is column 0
    if logit.shape[0] == 0:
        return torch.tensor(float("nan"), device=logit.device)
    return (logit.argmax(1) == 0).to(logit.dtype).mean() * 100.0


def nce_losses(logits, use_depth=None, use_rgb=None):
    """ContrastTrainer._compute_loss_accuracy (contrast_trainer.py:212-253) with target 0."""
    def ce(l):
        return F.cross_entropy(l, torch.zeros(l.shape[0], dtype=torch.long, device=l.device))

    if use_rgb is not None:
        sel = (use_depth == 1) & (use_rgb == 1)
        if sel.sum() == 0:
            losses = [(l - l).sum() for l in logits[:-2]] + [ce(l) for l in logits[-2:]]
            accs = [torch.zeros((), device=logits[0].device)] * 4 + [_top1(l) for l in logits[-2:]]
            return losses, accs
        return [ce(l[sel]) for l in logits], [_top1(l[sel]) for l in logits]
    if use_depth is not None:
        sel = use_depth == 1
        if use_depth.sum() == 0:
            losses = [(l - l).sum() for l in logits[:-2]] + [ce(l) for l in logits[-2:]]
            accs = [torch.zeros((), device=logits[0].device)] * 4 + [_top1(l) for l in logits[-2:]]
            return losses, accs
        losses = [ce(l[sel]) if i <= 3 else ce(l) for i, l in enumerate(logits)]
        accs = [_top1(l[sel]) if i <= 3 else _top1(l) for i, l in enumerate(logits)]
        return losses, accs
    return [ce(l) for l in logits], [_top1(l) for l in logits]


# --------------------------------------------------------------------------------------
# dense / sparse / SCL objectives (learning/contrast_trainer.py)
# --------------------------------------------------------------------------------------
def dense_kept_samples(depth_mask, h):
    """Samples whose nearest-resized mask is non-empty (contrast_trainer.py:674-682)."""
    m = F.interpolate(depth_mask.unsqueeze(1), size=(h, h), mode="nearest").reshape(depth_mask.shape[0], -1)
    return m, m.sum(-1) > 0


def dense_loss(G1, G2, depth_mask, sample_idx, use_depth=None, T=0.07):
    """_compute_soft_pri3d_loss_accuracy (contrast_trainer.py:642-723).

    sample_idx [B,S] int64: the injected multinomial draw for *every* sample (rows of samples that
    are dropped by the mask are ignored).  Returns ([loss_r2d, loss_d2r], [acc_r2d, acc_d2r]).
    """
    if use_depth is not None and use_depth.sum() == 0:
        z = (G1 - G1 + G2 - G2).mean()
        return [z, z], [torch.zeros((), device=z.device), torch.zeros((), device=z.device)]
    B, C, h, w = G1.shape
    _, keep = dense_kept_samples(depth_mask, h)
    idx = sample_idx[keep]
    g1 = G1.reshape(B, C, h * w)[keep]
    g2 = G2.reshape(B, C, h * w)[keep]
    S = idx.shape[1]
    gi = idx.unsqueeze(1).expand(-1, C, -1)
    a = F.normalize(torch.gather(g1, 2, gi), dim=1)      # [B',C,S]  rgb
    d = F.normalize(torch.gather(g2, 2, gi), dim=1)      # [B',C,S]  depth
    L = torch.matmul(d.permute(0, 2, 1), a) / T          # rgb2depth_logits[b,i,j] = <d_i, a_j>/T
    Lt = torch.matmul(a.permute(0, 2, 1), d) / T         # depth2rgb_logits = L^T
    xy = torch.stack([idx // w, idx % w], -1).to(G1.dtype)
    dist = ((xy.unsqueeze(2) - xy.unsqueeze(1)) ** 2).sum(-1).sqrt()
    soft = torch.softmax(-dist, 1)
    losses = [-(soft * F.log_softmax(L, 1)).sum(-2).mean(),
              -(soft * F.log_softmax(Lt, 1)).sum(-2).mean()]
    tgt = torch.arange(S, device=L.device).unsqueeze(0)
    accs = [((L.argmax(-2) == tgt).sum(-1).to(L.dtype) / S).mean(),
            ((Lt.argmax(-2) == tgt).sum(-1).to(L.dtype) / S).mean()]
    return losses, accs


def joint_pixel_index(joints_yx, h):
    """clamp(floor(joint/4), 0, h-1) -> y*h+x (contrast_trainer.py:754-760)."""
    q = (joints_yx // 4).long().clamp(0, h - 1)
    return q[:, :, 0] * h + q[:, :, 1]


def joint_loss(G1, G2, feat3, joints_yx, joints_vis, use_depth=None, T=0.07):
    """_compute_joints_pri3d_loss_accuracy (contrast_trainer.py:744-828)."""
    B, C, h, w = G1.shape
    J = joints_vis.shape[1]
    p = joint_pixel_index(joints_yx, h).unsqueeze(1).expand(-1, C, -1)
    a = F.normalize(torch.gather(G1.reshape(B, C, h * w), 2, p), dim=1)   # [B,C,J]
    d = F.normalize(torch.gather(G2.reshape(B, C, h * w), 2, p), dim=1)
    s = F.normalize(feat3, dim=-1)                                        # [B,J,C]
    Lr = torch.matmul(s, a) / T                                           # [B,J(skel),J(pixel)]
    Ld = torch.matmul(s, d) / T
    tgt = torch.arange(J, device=Lr.device).unsqueeze(0).repeat(B, 1)
    tgt[joints_vis == 0] = -100
    dtgt = tgt.clone()
    if use_depth is not None:
        dtgt[use_depth == 0] = -100
    losses = [F.cross_entropy(Lr, tgt), F.cross_entropy(Ld, dtgt)]
    accs = []
    for Lx, t in ((Lr, tgt), (Ld, dtgt)):
        cnt = (t != -100).sum(-1)
        hit = (Lx.argmax(-2) == t).sum(-1).to(Lx.dtype) / cnt.clamp(min=1)
        accs.append(hit[cnt != 0].mean())
    return losses, accs


def scl_loss(G1, G2, joints_yx, use_depth, use_rgb=None, T=0.07):
    """_compute_cross_subject_joints_pri3d_loss (contrast_trainer.py:830-892) with the
    `use_rgb is None -> all ones` semantics of segment_trainer.py:601-606."""
    if use_depth is not None and use_depth.sum() == 0:
        return (G1 - G1 + G2 - G2).mean()
    B, C, h, w = G1.shape
    J = joints_yx.shape[1]
    if use_rgb is None:
        use_rgb = torch.ones(B, dtype=torch.long, device=G1.device)
    if use_depth is None:
        use_depth = torch.ones(B, dtype=torch.long, device=G1.device)
    p = joint_pixel_index(joints_yx, h).unsqueeze(1).expand(-1, C, -1)
    a = F.normalize(torch.gather(G1.reshape(B, C, h * w), 2, p), dim=1).permute(0, 2, 1).reshape(B * J, C)
    d = F.normalize(torch.gather(G2.reshape(B, C, h * w), 2, p), dim=1).permute(0, 2, 1).reshape(B * J, C)
    Fm = torch.cat([a, d], 0)
    N = 2 * B * J
    logp = F.log_softmax(Fm @ Fm.t() / T, 1)
    r = torch.arange(N, device=G1.device)
    pos = ((r.view(-1, 1) % J) == (r.view(1, -1) % J)) & (r.view(-1, 1) != r.view(1, -1))
    off = torch.cat([(use_rgb == 0).view(B, 1).expand(B, J).reshape(-1),
                     (use_depth == 0).view(B, 1).expand(B, J).reshape(-1)])
    pos = pos & ~off.view(-1, 1) & ~off.view(1, -1)
    pos = pos.to(logp.dtype)
    return (-(logp * pos).sum(-1) / pos.sum(-1).clamp(min=1)).mean()


# --------------------------------------------------------------------------------------
# full step (contrast_trainer.py:532-640 first stage, :894-1039 second stage; main_contrast.py:78-81)
# --------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------ segmentation fine-tuning head
def fcn_layout(n_class=25, channels=128):
    """state_dict of networks/build_linear.py:4-15 `FCNHead(128, 128, n_class, num_convs=1, kernel_size=1)` (networks/fcn.py:5-33: the
    ConvModule registers its BatchNorm as `norm_name` BEFORE its conv)."""
    c = channels
    return OrderedDict([
        ("convs.0.norm_name.weight", (c,)), ("convs.0.norm_name.bias", (c,)), ("convs.0.norm_name.running_mean", (c,)),
        ("convs.0.norm_name.running_var", (c,)), ("convs.0.norm_name.num_batches_tracked", ()),
        ("convs.0.conv.weight", (c, c, 1, 1)), ("convs.0.conv.bias", (c,)),
        ("conv_seg.weight", (n_class, c, 1, 1)), ("conv_seg.bias", (n_class,)),
    ])


def fcn_forward(C, x, train=True):
    """networks/fcn.py:104-110: conv -> BN(momentum 0.1) -> ReLU -> conv_seg -> bilinear x4 (align_corners=False)."""
    bk = "convs.0.norm_name."
    y = F.conv2d(x, C["convs.0.conv.weight"], C["convs.0.conv.bias"])
    if train:
        with torch.no_grad():
            C[bk + "num_batches_tracked"] += 1
    y = F.batch_norm(y, C[bk + "running_mean"], C[bk + "running_var"], C[bk + "weight"], C[bk + "bias"], train, 0.1, BN_EPS)
    y = F.relu(y)
    logits = F.conv2d(y, C["conv_seg.weight"], C["conv_seg.bias"])
    return F.interpolate(logits, size=(logits.shape[2] * 4, logits.shape[3] * 4), mode="bilinear", align_corners=False)


def seg_loss(C, G1, G2, label, true_label, supervise_type=0, class_weights=None, ignore_index=255):
    """learning/segment_trainer.py:722-748 + main_segmentor.py:76-79 + :375-379.  Returns (loss_seg, aAcc) — (0, 0) with the
    BN-statistics-only forward of :742-748 when no sample carries a label."""
    sel = true_label.bool()
    if int(sel.sum()) == 0 or supervise_type not in (0, 1, 2):
        tmp = fcn_forward(C, G1)
        return (tmp - tmp).mean(), torch.zeros(())
    if supervise_type == 0:
        a, b = F.normalize(G1[sel], dim=1), F.normalize(G2[sel], dim=1)
        feat = torch.max(torch.stack([a, b]), 0)[0]
    else:
        feat = F.normalize((G1 if supervise_type == 1 else G2)[sel], dim=1)
    out = fcn_forward(C, feat)
    lab = label[sel]
    loss = F.cross_entropy(out, lab, weight=class_weights, ignore_index=ignore_index)
    aacc = (out.argmax(1) == lab).sum().float() / float(lab.numel())
    return loss, aacc


def make_momentum(P):
    return {k: torch.zeros_like(v) for k, v in P.items() if is_param(k)}


def sgd_step(P, grads, mom, lr=0.03, momentum=0.9, wd=1e-4, first=False):
    """torch.optim.SGD semantics: g += wd*p; buf = g (first step) or momentum*buf + g; p -= lr*buf."""
    with torch.no_grad():
        for k, g in grads.items():
            g = g + wd * P[k]
            if first:
                mom[k].copy_(g)
            else:
                mom[k].mul_(momentum).add_(g)
            P[k].sub_(lr * mom[k])


def train_step(P, mom, banks, batch, nce_idx, dense_idx=None, *, width=18, skeleton="mpii", stage=1,
               T=0.07, nce_m=0.5, lr=0.03, momentum=0.9, wd=1e-4, first=False, all_gather=None,
               apply_update=True, seg=None):
    """One pre-train step on one rank.  batch = dict(x, index, skeleton, joints_yx, joints_vis,
    use_depth, depth_mask).  Returns dict of losses/accs/f/grads."""
    for k, v in P.items():
        if is_param(k):
            v.requires_grad_(True)
            v.grad = None
    out = model_forward(P, batch["x"], batch["skeleton"], width, skeleton, stage, True)
    f = out["f"]
    f1, f2, f3 = torch.chunk(f, 3, dim=1)
    logits = nce_logits(banks, (f1, f2, f3), nce_idx, T)
    use_depth = batch.get("use_depth")
    losses, accs = nce_losses(logits, use_depth)
    res = dict(f=f.detach().clone(), nce_losses=[l.detach() for l in losses], nce_accs=accs)
    loss = sum(losses)
    if stage == 2:
        G1, G2 = out["linear_merge1"], out["linear_merge2"]
        dl, da = dense_loss(G1, G2, batch["depth_mask"], dense_idx, use_depth, T)
        jl, ja = joint_loss(G1, G2, out["feat3"], batch["joints_yx"], batch["joints_vis"], use_depth, T)
        sl = scl_loss(G1, G2, batch["joints_yx"], use_depth, None, T)
        loss = loss + sum(dl) + sum(jl) + sl
        res.update(dense_losses=[l.detach() for l in dl], dense_accs=da,
                   joint_losses=[l.detach() for l in jl], joint_accs=ja, scl_loss=sl.detach(),
                   linear_merge1=G1.detach(), linear_merge2=G2.detach(), feat3=out["feat3"].detach())
    if seg is not None:
        # SegTrainer.train_soft_joint_pri3d (segment_trainer.py:617-824): the same step + 10 * the FCN-head loss.
        # seg = dict(C=classifier state, mom=its momentum, label [B,R,R], true_label [B], supervise_type, class_weights)
        Cc = seg["C"]
        for k, v in Cc.items():
            if is_param(k):
                v.requires_grad_(True)
                v.grad = None
        ls, aacc = seg_loss(Cc, out["linear_merge1"], out["linear_merge2"], seg["label"], seg["true_label"],
                            seg.get("supervise_type", 0), seg.get("class_weights"))
        loss = loss + 10.0 * ls
        res.update(seg_loss=ls.detach(), seg_aacc=aacc)
    res["loss"] = loss.detach()
    loss.backward()
    if seg is not None:
        Cc = seg["C"]
        cg = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Cc.items() if is_param(k)}
        res["seg_grads"] = {k: g.clone() for k, g in cg.items()}
        if apply_update:
            sgd_step(Cc, cg, seg["mom"], lr, momentum, wd, first)
        for k, v in Cc.items():
            if is_param(k):
                v.requires_grad_(False)
    grads = {k: v.grad for k, v in P.items() if is_param(k) and v.grad is not None}
    res["grads"] = {k: g.clone() for k, g in grads.items()}
    # bank update with the (all-gathered) features (mem_bank.py:195-199)
    if all_gather is None:
        all_f, all_y = f.detach(), batch["index"]
    else:
        all_f, all_y = all_gather
    a1, a2, a3 = torch.chunk(all_f, 3, dim=1)
    for bk, ax in zip(banks, (a1, a2, a3)):
        bank_update(bk, ax, all_y, nce_m)
    if apply_update:
        sgd_step(P, grads, mom, lr, momentum, wd, first)
    for k, v in P.items():
        if is_param(k):
            v.requires_grad_(False)
    return res


def lr_at_epoch(epoch, epochs, lr=0.03, cosine=True, decay_rate=0.1, decay_epochs=()):
    """BaseTrainer.adjust_learning_rate (learning/base_trainer.py:80-93)."""
    if cosine:
        eta_min = lr * decay_rate ** 3
        return eta_min + (lr - eta_min) * (1 + math.cos(math.pi * epoch / epochs)) / 2
    steps = sum(1 for e in decay_epochs if epoch > e)
    return lr * decay_rate ** steps if steps > 0 else lr
