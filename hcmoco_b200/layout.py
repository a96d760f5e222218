"""Checkpoint key layout of the pre-train model (`CMC3HRNetSGCNSingleHead`, modal RGBD2S, arch HRNet).

The engine keeps its parameters in flat device buffers; this module defines, in the reference's
`state_dict()` order, which keys exist and what shape each has, so that checkpoints written by the
reference load into the engine and vice versa (networks/build_backbone.py:186-245,
networks/official_hrnet/official_hrnet.py:258-327, networks/SGCN/sem_gcn.py:60-89).
"""
from collections import OrderedDict

# stage-4 NUM_CHANNELS of networks/official_hrnet/seg_hrnet_w{18,32,48}*.yaml (stages 2,3 use prefixes)
WIDTHS = {18: (18, 36, 72, 144), 32: (32, 64, 128, 256), 48: (48, 96, 192, 384)}
# stage2..4: (number of HR modules, number of branches); every branch = 4 BasicBlocks
STAGES = ((1, 2), (4, 3), (3, 4))
# networks/SGCN/skeleton_meta.py:3-23 : parent joint of every joint (-1 = root)
SKELETONS = {
    "mpii": (1, 2, 6, 6, 3, 4, -1, 6, 7, 8, 11, 12, 8, 8, 13, 14),
    "coco_reduce": (1, 2, 9, 10, 3, 4, -1, 8, 9, 6, 6, 10, 11),
}
BN_SUFFIXES = ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")


def graph_edges(skeleton):
    """(J, rows, cols): non-zeros of the symmetric adjacency + self loops, row-major
    (graph_utils.py:27-45; sem_graph_conv.py:23-25 orders `e` by `(adj > 0).nonzero()`)."""
    parents = SKELETONS[skeleton]
    J = len(parents)
    cells = {(j, j) for j in range(J)}
    for child, parent in enumerate(parents):
        if parent >= 0:
            cells.update(((child, parent), (parent, child)))
    cells = sorted(cells)
    return J, [r for r, _ in cells], [c for _, c in cells]


class _Keys:
    def __init__(self):
        self.d = OrderedDict()

    def conv(self, name, cout, cin, ks):
        self.d[name + ".weight"] = (cout, cin, ks, ks)

    def bn(self, name, c):
        for s in BN_SUFFIXES:
            self.d["%s.%s" % (name, s)] = () if s == "num_batches_tracked" else (c,)

    def conv_bn(self, cname, bname, cout, cin, ks):
        self.conv(cname, cout, cin, ks)
        self.bn(bname, cout)


def _hrnet(k, p, width):
    ch = WIDTHS[width]
    k.conv_bn(p + "conv1", p + "bn1", 64, 3, 3)
    k.conv_bn(p + "conv2", p + "bn2", 64, 64, 3)
    for blk in range(4):
        q = "%slayer1.%d." % (p, blk)
        k.conv_bn(q + "conv1", q + "bn1", 64, 256 if blk else 64, 1)
        k.conv_bn(q + "conv2", q + "bn2", 64, 64, 3)
        k.conv_bn(q + "conv3", q + "bn3", 256, 64, 1)
        if blk == 0:
            k.conv_bn(q + "downsample.0", q + "downsample.1", 256, 64, 1)
    before = (256,)
    for s, (nmod, nbr) in enumerate(STAGES):
        now = ch[:nbr]
        t = "%stransition%d." % (p, s + 1)
        for i in range(nbr):
            if i < len(before):
                if now[i] != before[i]:
                    k.conv_bn("%s%d.0" % (t, i), "%s%d.1" % (t, i), now[i], before[i], 3)
            else:
                hops = i + 1 - len(before)
                for j in range(hops):
                    cout = now[i] if j == hops - 1 else before[-1]
                    k.conv_bn("%s%d.%d.0" % (t, i, j), "%s%d.%d.1" % (t, i, j), cout, before[-1], 3)
        for m in range(nmod):
            mp = "%sstage%d.%d." % (p, s + 2, m)
            for i in range(nbr):
                for blk in range(4):
                    q = "%sbranches.%d.%d." % (mp, i, blk)
                    k.conv_bn(q + "conv1", q + "bn1", now[i], now[i], 3)
                    k.conv_bn(q + "conv2", q + "bn2", now[i], now[i], 3)
            for i in range(nbr):
                for j in range(nbr):
                    q = "%sfuse_layers.%d.%d." % (mp, i, j)
                    if j > i:
                        k.conv_bn(q + "0", q + "1", now[i], now[j], 1)
                    elif j < i:
                        for hop in range(i - j):
                            cout = now[i] if hop == i - j - 1 else now[j]
                            k.conv_bn("%s%d.0" % (q, hop), "%s%d.1" % (q, hop), cout, now[j], 3)
        before = now


def _sgcn(k, p, skeleton, hid=128):
    nnz = len(graph_edges(skeleton)[1])

    def gconv(name, cin, cout):
        k.d[name + ".W"] = (2, cin, cout)
        k.d[name + ".e"] = (1, nnz)
        k.d[name + ".bias"] = (cout,)

    gconv(p + "gconv_input.0.gconv", 2, hid)
    k.bn(p + "gconv_input.0.bn", hid)
    for layer in range(4):
        for g in (1, 2):
            q = "%sgconv_layers.%d.gconv%d" % (p, layer, g)
            gconv(q + ".gconv", hid, hid)
            k.bn(q + ".bn", hid)
    gconv(p + "gconv_output", hid, hid)


def model_keys(width=18, stage=1, skeleton="mpii", feat_dim=128):
    """Ordered {key: shape} identical to the reference model's state_dict()."""
    k = _Keys()
    cm = sum(WIDTHS[width])
    _hrnet(k, "encoder1.", width)
    _hrnet(k, "encoder2.", width)
    _sgcn(k, "encoder3.", skeleton)
    for i, cin in ((1, cm), (2, cm), (3, 128)):
        k.d["head%d.0.weight" % i] = (feat_dim, cin)
        k.d["head%d.0.bias" % i] = (feat_dim,)
    if stage == 2:
        for i in (1, 2):
            k.d["encoder%d_linear.weight" % i] = (128, cm, 1, 1)
            k.d["encoder%d_linear.bias" % i] = (128,)
    return k.d


def is_buffer(key):
    return key.endswith(("running_mean", "running_var", "num_batches_tracked"))
