"""Plain-PyTorch fp32 statement of every C-ABI entry point (include/hcmoco.h).  TEST INFRASTRUCTURE.

Two uses, both only from tests/:
  * `-m gpu` kernel tests run each CUDA kernel and the function of the same name here on the same
    inputs (the "plain PyTorch fp32 reference of the same op");
  * `-m "not gpu"` host-logic tests build the engine's launch programs with `TorchKernels` in place
    of `CudaKernels`, so the plan wiring (buffers, accumulate flags, backward order) is checked
    against the oracle on CPU.  The product never selects this class.

Semantics follow the header exactly: outputs are written in place into the caller's tensors,
activations are channels-last [B,H,W,C] (any leading shape, channels last), weights OIHW.
"""
import torch
import torch.nn.functional as F


def _nchw(x, B, H, W, C):
    return x.reshape(B, H, W, C).permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _tf(x, sc, sh, relu, C):
    v = x.reshape(-1, C)
    if sc is not None:
        v = v * sc.reshape(1, C) + sh.reshape(1, C)
        if relu:
            v = v.clamp(min=0)
    return v


def _pc(t, P, C):
    return t.reshape(-1)[:P * C].reshape(P, C)


def _up(x, f):
    return F.interpolate(x, scale_factor=f, mode="bilinear", align_corners=False) if f > 1 else x


class TorchKernels:
    name = "torch-ref"

    def __init__(self, device="cpu", dtype=torch.float32):
        # dtype=float64 turns the host-logic tests into an exact wiring check (the fp32 gradient of this
        # network is only accurate to ~1e-2 against fp64 for any implementation, see tests/engine_check.py)
        self.device, self.dtype = device, dtype
        self.launches = 0

    def empty(self, *shape, dtype=None):
        # NaN-fill so that reading an unwritten buffer cannot go unnoticed
        dtype = dtype or self.dtype
        t = torch.empty(*shape, dtype=dtype, device=self.device)
        if dtype.is_floating_point:
            t.fill_(float("nan"))
        return t

    def zeros(self, *shape, dtype=None):
        return torch.zeros(*shape, dtype=dtype or self.dtype, device=self.device)

    @staticmethod
    def _put_part(part, rows, C, s0, s1):
        v = part.reshape(-1)[:rows * 2 * C].reshape(rows, 2, C)
        v.zero_()
        v[rows - 1, 0] = s0
        v[rows - 1, 1] = s1

    # ---------------------------------------------------------------- igemm.cu
    def conv2d_stat_rows(self, B, H, W, Cin, Cout, ks, stride):
        return 3

    def conv2d_fwd(self, x, w, bias, y, B, H, W, Cin, Cout, ks, stride, sc, sh, relu, stat_part):
        xin = _tf(x, sc, sh, relu, Cin).reshape(B, H, W, Cin).permute(0, 3, 1, 2)
        out = F.conv2d(xin, w.reshape(Cout, Cin, ks, ks), bias, stride, (ks - 1) // 2)
        o = _nhwc(out)
        y.reshape(-1).copy_(o.reshape(-1))
        if stat_part is not None:
            o2 = o.reshape(-1, Cout)
            self._put_part(stat_part, 3, Cout, o2.sum(0), (o2 * o2).sum(0))
        return 0

    def conv2d_dgrad(self, dy, w, dx, B, H, W, Cin, Cout, ks, stride, accumulate):
        pad = (ks - 1) // 2
        Ho, Wo = (H + 2 * pad - ks) // stride + 1, (W + 2 * pad - ks) // stride + 1
        g = torch.nn.grad.conv2d_input((B, Cin, H, W), w.reshape(Cout, Cin, ks, ks), _nchw(dy, B, Ho, Wo, Cout),
                                       stride, pad)
        g = _nhwc(g).reshape(-1)
        if accumulate:
            dx.reshape(-1).add_(g)
        else:
            dx.reshape(-1).copy_(g)
        return 0

    def conv2d_wgrad(self, x, dy, dw, B, H, W, Cin, Cout, ks, stride, sc, sh, relu):
        pad = (ks - 1) // 2
        Ho, Wo = (H + 2 * pad - ks) // stride + 1, (W + 2 * pad - ks) // stride + 1
        xin = _tf(x, sc, sh, relu, Cin).reshape(B, H, W, Cin).permute(0, 3, 1, 2)
        g = torch.nn.grad.conv2d_weight(xin, (Cout, Cin, ks, ks), _nchw(dy, B, Ho, Wo, Cout), stride, pad)
        dw.reshape(-1).add_(g.reshape(-1))
        return 0

    # ---------------------------------------------------------------- tc_conv.cu (exact fp32 statement)
    def tc_conv_supported(self, B, H, W, Cin, Cout, ks, stride):
        ok = (ks == 3 and stride in (1, 2)) or (ks == 1 and stride == 1)
        return int(ok and Cin % 2 == 0 and Cout % 2 == 0 and Cin <= 256 and Cout <= 256 and (stride == 1 or (H % 2 == 0 and W % 2 == 0)))

    def tc_conv_wpack_bytes(self, B, H, W, Cin, Cout, ks):
        return Cin * Cout * ks * ks * 8          # room for fp64 in the exact-wiring tests

    def tc_conv_rowcat_supported(self, Cout, ks, stride):
        return 0          # a layout detail of the CUDA kernel; the reference "pack" is the plain fp32 weight

    def tc_conv_pack(self, w, ldw, wpack, B, H, W, Cin, Cout, ks, flags):
        transpose = flags & 1
        # the "packed" form of the reference is simply the effective OIHW weight of the GEMM being run
        O, I = (Cin, Cout) if transpose else (Cout, Cin)                         # dims of the OIHW slice
        ld = ldw if ldw > 0 else I
        Iv = min(I, ld) if not transpose else I                                  # forward, ldw < Cin: K zero-padded to Cin
        wv = torch.as_strided(w, (O, Iv, ks, ks), (ld * ks * ks, ks * ks, ks, 1), w.storage_offset())
        if Iv < I:
            wv = torch.cat([wv, torch.zeros(O, I - Iv, ks, ks, dtype=w.dtype, device=w.device)], 1)
        if transpose:
            weff = wv.transpose(0, 1).flip(2, 3)                                 # w is [Cout(w)=Cin'][Cin(w)=Cout']
        else:
            weff = wv
        wpack.reshape(-1).view(w.dtype)[:weff.numel()].copy_(weff.reshape(-1))
        return 0

    def tc_conv(self, x, wpack, bias, y, B, H, W, Cin, Cout, ks, stride, sc, sh, relu, accumulate):
        weff = wpack.reshape(-1).view(x.dtype)[:Cin * Cout * ks * ks].reshape(Cout, Cin, ks, ks)
        xin = _tf(x, sc, sh, relu, Cin).reshape(B, H, W, Cin).permute(0, 3, 1, 2)
        o = _nhwc(F.conv2d(xin, weff, bias, stride, (ks - 1) // 2)).reshape(-1)
        if accumulate:
            y.reshape(-1).add_(o)
        else:
            y.reshape(-1).copy_(o)
        return 0

    def tc_dgrad_s2_supported(self, B, H, W, Cin, Cout):
        return int(H % 2 == 0 and W % 2 == 0 and Cin % 2 == 0 and Cout % 2 == 0 and Cin <= 256 and Cout <= 256)

    def tc_dgrad_s2_nqs(self, B, H, W, Cin, Cout):
        return 4

    def tc_dgrad_s2_wpack_bytes(self, B, H, W, Cin, Cout):
        return Cin * Cout * 9 * 8

    def tc_dgrad_s2_pack(self, w, wpack, B, H, W, Cin, Cout):
        wpack.reshape(-1).view(w.dtype)[:Cin * Cout * 9].copy_(w.reshape(-1)[:Cin * Cout * 9])
        return 0

    def tc_dgrad_s2(self, dy, wpack, dx, B, H, W, Cin, Cout, accumulate):
        w = wpack.reshape(-1).view(dy.dtype)[:Cin * Cout * 9]
        return self.conv2d_dgrad(dy, w, dx, B, H, W, Cin, Cout, 3, 2, accumulate)

    def tc_wgrad_supported(self, B, H, W, Cin, Cout, ks, stride):
        return self.tc_conv_supported(B, H, W, Cin, Cout, ks, stride)

    def tc_wgrad(self, x, dy, dw, lddw, B, H, W, Cin, Cout, ks, stride, sc, sh, relu):
        ld = lddw if lddw > 0 else Cin
        tmp = torch.zeros(Cout, Cin, ks, ks, dtype=dw.dtype, device=dw.device)
        self.conv2d_wgrad(x, dy, tmp, B, H, W, Cin, Cout, ks, stride, sc, sh, relu)
        Cw = min(Cin, ld)                                                        # lddw < Cin: dw has only lddw input channels
        torch.as_strided(dw, (Cout, Cw, ks, ks), (ld * ks * ks, ks * ks, ks, 1), dw.storage_offset()).add_(tmp[:, :Cw])
        return 0

    def gemm(self, A, Bm, bias, C, batch, M, N, K, sAm, sAk, sBk, sBn, sCm, bsA, bsB, bsC, alpha, accumulate):
        a = torch.as_strided(A.reshape(-1), (batch, M, K), (bsA, sAm, sAk))
        b = torch.as_strided(Bm.reshape(-1), (batch, K, N), (bsB, sBk, sBn))
        c = torch.as_strided(C.reshape(-1), (batch, M, N), (bsC, sCm, 1))
        r = alpha * torch.bmm(a, b)
        if bias is not None:
            r = r + bias.reshape(1, 1, N)
        if accumulate:
            r = r + c
        c.copy_(r)
        return 0

    # ---------------------------------------------------------------- bn.cu
    def colstat_rows(self, P, C):
        return 2

    def bn_stats(self, y, P, C, part):
        v = y.reshape(-1)[:P * C].reshape(P, C)
        self._put_part(part, 2, C, v.sum(0), (v * v).sum(0))
        return 0

    def bn_finalize(self, part, nparts, C, count, gamma, beta, rm, rv, nbt, momentum, eps, scale, shift, mean, invstd):
        p = part.reshape(-1)[:nparts * 2 * C].reshape(nparts, 2, C).double().sum(0)
        m = p[0] / count
        var = (p[1] / count - m * m).clamp(min=0)
        inv = 1.0 / torch.sqrt(var + eps)
        g = gamma.double() if gamma is not None else torch.ones_like(m)
        b = beta.double() if beta is not None else torch.zeros_like(m)
        scale.copy_(g * inv)
        shift.copy_(b - m * g * inv)
        mean.copy_(m)
        invstd.copy_(inv)
        if rm is not None:
            unb = var * count / (count - 1.0) if count > 1 else var
            rm.mul_(1 - momentum).add_(momentum * m.to(rm.dtype))
            rv.mul_(1 - momentum).add_(momentum * unb.to(rv.dtype))
        if nbt is not None:
            nbt += 1
        return 0

    def bn_stats_finalize(self, y, P, C, part, counter, gamma, beta, rm, rv, nbt, momentum, eps, scale, shift, mean, invstd):
        self.bn_stats(y, P, C, part)
        return self.bn_finalize(part, self.colstat_rows(P, C), C, P, gamma, beta, rm, rv, nbt, momentum, eps, scale, shift,
                                mean, invstd)

    def bn_bwd_reduce_finalize(self, dz, mask, msc, msh, y, mean, invstd, P, C, part, counter, gamma, dgamma, dbeta, k1, k2, k3):
        self.bn_bwd_reduce(dz, mask, msc, msh, y, mean, invstd, P, C, part)
        return self.bn_bwd_finalize(part, self.colstat_rows(P, C), C, P, gamma, mean, invstd, dgamma, dbeta, k1, k2, k3)

    def bn_apply(self, y, scale, shift, res, rs, rh, relu, out, P, C):
        v = _pc(y, P, C)
        if scale is not None:
            v = v * scale[:C] + (shift[:C] if shift is not None else 0)
        if res is not None:
            r = _pc(res, P, C)
            if rs is not None:
                r = r * rs[:C] + rh[:C]
            v = v + r
        if relu:
            v = v.clamp(min=0)
        _pc(out, P, C).copy_(v)
        return 0

    @staticmethod
    def _masked(dz, mask, msc, msh, y, P, C):
        g = _pc(dz, P, C)
        if mask is not None:
            g = g * (_pc(mask, P, C) > 0)
        elif msc is not None:
            g = g * ((_pc(y, P, C) * msc[:C] + msh[:C]) > 0)
        return g

    def bn_bwd_reduce(self, dz, mask, msc, msh, y, mean, invstd, P, C, part):
        g = self._masked(dz, mask, msc, msh, y, P, C)
        yh = (_pc(y, P, C) - mean[:C]) * invstd[:C]
        self._put_part(part, 2, C, g.sum(0), (g * yh).sum(0))
        return 0

    def bn_bwd_finalize(self, part, nparts, C, count, gamma, mean, invstd, dgamma, dbeta, k1, k2, k3):
        p = part.reshape(-1)[:nparts * 2 * C].reshape(nparts, 2, C).double().sum(0)
        s, q = p[0], p[1]
        g = gamma.double() if gamma is not None else torch.ones_like(s)
        is_, mu = invstd[:C].double(), mean[:C].double()
        if dgamma is not None:
            dgamma.copy_(q)
        if dbeta is not None:
            dbeta.copy_(s)
        a = g * is_
        k1[:C] = a
        k2[:C] = -a * is_ * q / count
        k3[:C] = -a * s / count + a * is_ * mu * q / count
        return 0

    def bn_bwd_apply(self, dz, mask, msc, msh, y, k1, k2, k3, dy, g_out, g_acc, P, C):
        g = self._masked(dz, mask, msc, msh, y, P, C)
        o = k1[:C] * g + k2[:C] * _pc(y, P, C) + k3[:C]
        if g_out is not None:
            if g_acc:
                _pc(g_out, P, C).add_(g)
            else:
                _pc(g_out, P, C).copy_(g)
        _pc(dy, P, C).copy_(o)
        return 0

    def relu_bwd(self, dout, out, g, accumulate, total):
        v = dout.reshape(-1)[:total] * (out.reshape(-1)[:total] > 0)
        if accumulate:
            g.reshape(-1)[:total].add_(v)
        else:
            g.reshape(-1)[:total].copy_(v)
        return 0

    def axpy(self, dst, src, alpha, total):
        dst.reshape(-1)[:total].add_(alpha * src.reshape(-1)[:total])
        return 0

    # ---------------------------------------------------------------- resample.cu
    def nchw_to_nhwc_pad(self, x, out, B, Ctot, HW, coff, Cn, Cpad):
        o = out.reshape(B, HW, Cpad)
        o.zero_()
        o[:, :, :Cn] = x.reshape(B, Ctot, HW)[:, coff:coff + Cn].permute(0, 2, 1)
        return 0

    def nchw_to_nhwc(self, x, out, B, Ctot, HW, coff, Cn):
        v = x.reshape(B, Ctot, HW)[:, coff:coff + Cn].permute(0, 2, 1)
        out.reshape(B, HW, Cn).copy_(v)
        return 0

    def fuse_sum(self, nterms, ptrs, scales, shifts, log2f, bias, relu, out, B, H, W, C):
        acc = torch.zeros(B, C, H, W, device=out.device, dtype=out.dtype)
        for i in range(nterms):
            k = int(log2f[i])
            t = _up(_nchw(ptrs[i], B, H >> k, W >> k, C), 1 << k)
            if scales is not None and scales[i] is not None:
                t = t * scales[i].reshape(1, C, 1, 1)
                if shifts is not None and shifts[i] is not None:
                    t = t + shifts[i].reshape(1, C, 1, 1)
            acc = acc + t
        if bias is not None:
            acc = acc + bias.reshape(1, C, 1, 1)
        if relu:
            acc = acc.clamp(min=0)
        out.reshape(-1).copy_(_nhwc(acc).reshape(-1))
        return 0

    def upsample_adjoint(self, g, out, accumulate, B, H, W, C, log2f):
        f = 1 << log2f
        with torch.enable_grad():          # may be called from inside an autograd.Function.backward (grad mode off)
            src = torch.zeros(B, C, H // f, W // f, device=g.device, dtype=g.dtype, requires_grad=True)
            up = _up(src, f)
            (gr,) = torch.autograd.grad(up, src, _nchw(g, B, H, W, C).detach())
        r = _nhwc(gr).reshape(-1)
        if accumulate:
            out.reshape(-1).add_(r)
        else:
            out.reshape(-1).copy_(r)
        return 0

    def avgpool(self, x, out, B, HW, C, ldo, coff):
        out.reshape(B, ldo)[:, coff:coff + C] = x.reshape(B, HW, C).mean(1)
        return 0

    def avgpool_bwd(self, dout, dx, accumulate, B, HW, C, ldo, coff):
        v = (dout.reshape(B, ldo)[:, coff:coff + C] / HW).reshape(B, 1, C).expand(B, HW, C)
        if accumulate:
            dx.reshape(B, HW, C).add_(v)
        else:
            dx.reshape(B, HW, C).copy_(v)
        return 0

    # ---------------------------------------------------------------- nce.cu
    PAIRS = ((0, 1), (1, 0), (1, 2), (2, 1), (0, 2), (2, 0))

    def nce_logits(self, b1, b2, b3, x1, x2, x3, ldx, idx, B, K1, dim, T, logits):
        banks = (b1, b2, b3)
        xs = [torch.as_strided(x, (B, dim), (ldx, 1), x.storage_offset()) for x in (x1, x2, x3)]
        w = [bk.reshape(-1, dim).index_select(0, idx.reshape(-1)).view(B, K1, dim) for bk in banks]
        for q, (p, bq) in enumerate(self.PAIRS):
            logits.reshape(6, B, K1)[q] = torch.bmm(w[bq], xs[p].unsqueeze(2)).squeeze(2) / T
        return 0

    def nce_loss(self, logits, B, K1, use_depth, use_rgb, lse, l0, hit, coef, loss6, acc6):
        L = logits.reshape(6, B, K1)
        lse.reshape(6, B).copy_(torch.logsumexp(L, 2))
        l0.reshape(6, B).copy_(L[:, :, 0])
        hit.reshape(6, B).copy_((L[:, :, 0] >= L.max(2).values).float())
        d = (use_depth == 1) if use_depth is not None else torch.ones(B, dtype=torch.bool, device=L.device)
        r = (use_rgb == 1) if use_rgb is not None else torch.ones(B, dtype=torch.bool, device=L.device)
        both = d & r
        for pair in range(6):
            if use_rgb is not None:
                masked = bool(both.sum() > 0) or pair < 4
            else:
                masked = use_depth is not None and pair < 4
            sel = both if masked else torch.ones(B, dtype=torch.bool, device=L.device)
            cs = sel.sum().to(L.dtype)
            inv = (1.0 / cs) if cs > 0 else torch.zeros((), dtype=L.dtype)
            loss6[pair] = ((lse.reshape(6, B)[pair] - l0.reshape(6, B)[pair]) * sel).sum() * inv
            acc6[pair] = 100.0 * (hit.reshape(6, B)[pair] * sel).sum() * inv
            coef.reshape(6, B)[pair] = sel.to(L.dtype) * inv
        return 0

    def nce_bwd(self, b1, b2, b3, x1, x2, x3, ldx, idx, B, K1, dim, T, logits, lse, coef, gscale, df, lddf):
        banks = (b1, b2, b3)
        w = [bk.reshape(-1, dim).index_select(0, idx.reshape(-1)).view(B, K1, dim) for bk in banks]
        L = logits.reshape(6, B, K1)
        dfv = torch.as_strided(df, (B, 3, dim), (lddf, dim, 1), df.storage_offset())
        for q, (p, bq) in enumerate(self.PAIRS):
            if lse is None:          # generic: `logits` holds d(loss)/d(logits)
                s = L[q] * (gscale / T)
            else:
                s = torch.exp(L[q] - lse.reshape(6, B)[q].unsqueeze(1))
                s[:, 0] -= 1.0
                s = s * (coef.reshape(6, B)[q] * gscale / T).unsqueeze(1)
            dfv[:, p] += torch.bmm(s.unsqueeze(1), w[bq]).squeeze(1)
        return 0

    def bank_update(self, bank, x, ldx, y, N, dim, m):
        xs = torch.as_strided(x, (N, dim), (ldx, 1), x.storage_offset())
        bk = bank.reshape(-1, dim)
        w = bk.index_select(0, y) * m + xs * (1 - m)
        w = F.normalize(w)
        for i in range(N):      # last writer wins
            bk[y[i]] = w[i]
        return 0

    # ---------------------------------------------------------------- losses.cu
    def gather_l2norm(self, src, lds, pix, HW, rows_per_b, nrows, dim, out, ldo, inv_norm):
        if pix is not None:
            r = torch.arange(nrows, device=src.device)
            rows = src.reshape(-1, dim)[(r // rows_per_b) * HW + pix.reshape(-1)]
        else:
            rows = torch.as_strided(src, (nrows, dim), (lds, 1), src.storage_offset())
        inv = 1.0 / rows.norm(dim=1).clamp(min=1e-12)
        torch.as_strided(out, (nrows, dim), (ldo, 1), out.storage_offset()).copy_(rows * inv.unsqueeze(1))
        if inv_norm is not None:
            inv_norm.reshape(-1)[:nrows] = inv
        return 0

    def gather_l2norm_bwd(self, dout, lddo, out, ldo, inv_norm, pix, HW, rows_per_b, nrows, dim, dsrc, lds, accumulate):
        g = torch.as_strided(dout, (nrows, dim), (lddo, 1), dout.storage_offset())
        o = torch.as_strided(out, (nrows, dim), (ldo, 1), out.storage_offset())
        d = (g - o * (g * o).sum(1, keepdim=True)) * inv_norm.reshape(-1)[:nrows].unsqueeze(1)
        if pix is not None:
            r = torch.arange(nrows, device=dout.device)
            dsrc.reshape(-1, dim).index_add_(0, (r // rows_per_b) * HW + pix.reshape(-1), d)
        else:
            t = torch.as_strided(dsrc, (nrows, dim), (lds, 1), dsrc.storage_offset())
            if accumulate:
                t.add_(d)
            else:
                t.copy_(d)
        return 0

    def joint_pixel_index(self, joints_yx, n, h, pix):
        q = torch.floor(joints_yx.reshape(n, 2) / 4.0).long().clamp(0, h - 1)
        pix.reshape(-1)[:n] = q[:, 0] * h + q[:, 1]
        return 0

    def dense_kept(self, depth_mask, B, R, h, kept):
        step = R // h
        m = depth_mask.reshape(B, R, R)[:, ::step, ::step][:, :h, :h].reshape(B, -1)
        kept.reshape(-1)[:B] = ((m != 0).sum(1) > 0).float()
        return 0

    @staticmethod
    def _dense_w(pix, B, S, h, dtype):
        xy = torch.stack([pix.reshape(B, S) // h, pix.reshape(B, S) % h], -1).to(dtype)
        dist = ((xy.unsqueeze(2) - xy.unsqueeze(1)) ** 2).sum(-1).sqrt()
        return torch.exp(-dist)          # [B,S,S] symmetric

    def dense_stats(self, L, pix, kept, use_depth, B, S, h, stat, fin):
        Lb = L.reshape(B, S, S)
        w = self._dense_w(pix, B, S, h, L.dtype)
        st = stat.reshape(B, 2, S, 4)
        tgt = torch.arange(S, device=L.device)
        for d, M in ((0, Lb), (1, Lb.transpose(1, 2))):      # statistics over i (dim 1) for fixed j
            st[:, d, :, 0] = torch.logsumexp(M, 1)
            st[:, d, :, 1] = w.sum(1)
            st[:, d, :, 2] = (w * M).sum(1)
            st[:, d, :, 3] = (M.argmax(1) == tgt).float()
        k = kept.reshape(-1)[:B]
        nk = k.sum()
        nd = (use_depth != 0).sum() if use_depth is not None else torch.tensor(B)
        on = bool(nd > 0) and bool(nk > 0)
        inv = 1.0 / (nk * S) if on else 0.0
        per = st[..., 0] - st[..., 2] / st[..., 1]           # [B,2,S]
        fin[0] = (per[:, 0] * k.unsqueeze(1)).sum() * inv
        fin[1] = (per[:, 1] * k.unsqueeze(1)).sum() * inv
        fin[2] = (st[:, 0, :, 3] * k.unsqueeze(1)).sum() * inv
        fin[3] = (st[:, 1, :, 3] * k.unsqueeze(1)).sum() * inv
        fin[4] = nk if on else 0.0
        return 0

    def dense_grad(self, L, pix, stat, kept, fin, B, S, h, gscale):
        Lb = L.reshape(B, S, S)
        st = stat.reshape(B, 2, S, 4)
        w = self._dense_w(pix, B, S, h, L.dtype)
        nk = float(fin[4])
        coef = gscale / (nk * S) if nk > 0 else 0.0
        col_lse, col_Z = st[:, 0, :, 0].unsqueeze(1), st[:, 0, :, 1].unsqueeze(1)     # indexed by c
        row_lse, row_Z = st[:, 1, :, 0].unsqueeze(2), st[:, 1, :, 1].unsqueeze(2)     # indexed by r
        g = coef * (torch.exp(Lb - col_lse) + torch.exp(Lb - row_lse) - w * (1.0 / col_Z + 1.0 / row_Z))
        g = g * (kept.reshape(-1)[:B] != 0).float().view(B, 1, 1)
        Lb.copy_(g)
        return 0

    def dense_finish(self, stat, kept, use_depth, B, S, fin):
        st = stat.reshape(B, 2, S, 4)
        k = kept.reshape(-1)[:B]
        nk = k.sum()
        nd = (use_depth != 0).sum() if use_depth is not None else torch.tensor(B)
        on = bool(nd > 0) and bool(nk > 0)
        inv = 1.0 / (nk * S) if on else 0.0
        per = st[..., 0] - st[..., 2] / st[..., 1]
        per = torch.where(k.view(B, 1, 1) != 0, per, torch.zeros_like(per))
        hit = torch.where(k.view(B, 1, 1) != 0, st[..., 3], torch.zeros_like(per))
        fin[0], fin[1] = per[:, 0].sum() * inv, per[:, 1].sum() * inv
        fin[2], fin[3] = hit[:, 0].sum() * inv, hit[:, 1].sum() * inv
        fin[4] = nk if on else 0.0
        return 0

    # fused form (dense_affinity.cu) = gather_l2norm x2 -> affinity GEMM -> dense_stats / dense_grad -> GEMMs -> scatter
    def dense_affinity_work_bytes(self, B, S):
        return 16

    def dense_affinity_fwd(self, G1, G2, pix, kept, use_depth, B, S, h, dim, inv_T, stat, fin, work=None):
        HW = h * h
        A, D = torch.empty(B * S, dim, dtype=G1.dtype, device=G1.device), torch.empty(B * S, dim, dtype=G1.dtype, device=G1.device)
        self.gather_l2norm(G1, 0, pix, HW, S, B * S, dim, A, dim, None)
        self.gather_l2norm(G2, 0, pix, HW, S, B * S, dim, D, dim, None)
        L = torch.matmul(D.reshape(B, S, dim), A.reshape(B, S, dim).transpose(1, 2)) * inv_T
        k = kept.reshape(-1)[:B]
        st = torch.zeros(B, 2, S, 4, dtype=G1.dtype, device=G1.device)
        self.dense_stats(L, pix, kept, use_depth, B, S, h, st, fin)
        sv = stat.reshape(B, 2, S, 4)
        sv[k != 0] = st[k != 0]          # the fused kernel skips dropped samples
        return 0

    def dense_affinity_bwd(self, G1, G2, pix, stat, kept, fin, B, S, h, dim, inv_T, gscale_r2d, gscale_d2r, dG1, dG2, work=None,
                           prepared=0):
        HW = h * h
        mk = lambda *s: torch.empty(*s, dtype=G1.dtype, device=G1.device)      # noqa: E731
        A, D, ia, idn = mk(B * S, dim), mk(B * S, dim), mk(B * S), mk(B * S)
        self.gather_l2norm(G1, 0, pix, HW, S, B * S, dim, A, dim, ia)
        self.gather_l2norm(G2, 0, pix, HW, S, B * S, dim, D, dim, idn)
        A3, D3 = A.reshape(B, S, dim), D.reshape(B, S, dim)
        L = torch.matmul(D3, A3.transpose(1, 2)) * inv_T
        k = kept.reshape(-1)[:B]
        st = torch.where(k.view(B, 1, 1, 1) != 0, stat.reshape(B, 2, S, 4), torch.ones_like(stat.reshape(B, 2, S, 4)))
        # dense_grad with one weight per direction: loss_r2d owns the column statistics (st[:,0]), loss_d2r the row ones
        w = self._dense_w(pix, B, S, h, L.dtype)
        nk = float(fin[4])
        c0, c1 = (gscale_r2d / (nk * S), gscale_d2r / (nk * S)) if nk > 0 else (0.0, 0.0)
        col_lse, col_Z = st[:, 0, :, 0].unsqueeze(1), st[:, 0, :, 1].unsqueeze(1)
        row_lse, row_Z = st[:, 1, :, 0].unsqueeze(2), st[:, 1, :, 1].unsqueeze(2)
        g = c0 * (torch.exp(L - col_lse) - w / col_Z) + c1 * (torch.exp(L - row_lse) - w / row_Z)
        L = g * (k != 0).to(L.dtype).view(B, 1, 1)
        dD = (torch.matmul(L, A3) * inv_T).reshape(B * S, dim)
        dA = (torch.matmul(L.transpose(1, 2), D3) * inv_T).reshape(B * S, dim)
        self.gather_l2norm_bwd(dA.contiguous(), dim, A, dim, ia, pix, HW, S, B * S, dim, dG1, 0, 1)
        self.gather_l2norm_bwd(dD.contiguous(), dim, D, dim, idn, pix, HW, S, B * S, dim, dG2, 0, 1)
        return 0

    # ---- input staging (stage_input.cu): integer indexing / mask / mm sums exact, fp32 bilinear for the RGB planes
    def stage_input(self, rgb, depth, crop, flip, has_depth, B, Hs, Ws, R, sums, x, mask):
        dev = rgb.device
        yy = torch.arange(R, device=dev)
        mean_c = torch.tensor([0.485, 0.456, 0.406], device=dev, dtype=torch.float32)
        std_c = torch.tensor([0.229, 0.224, 0.225], device=dev, dtype=torch.float32)
        for b in range(B):
            i, j, h, w = [int(v) for v in crop[b]]
            fl = flip is not None and int(flip[b]) != 0
            hd = has_depth is None or int(has_depth[b]) != 0
            xs = (R - 1 - yy) if fl else yy                                   # source column of each output column
            # nearest source pixel (exact integer arithmetic)
            sy = i + torch.clamp(((2 * yy + 1) * h) // (2 * R), max=h - 1)
            sx = j + torch.clamp(((2 * xs + 1) * w) // (2 * R), max=w - 1)
            oky, okx = (sy >= 0) & (sy < Hs), (sx >= 0) & (sx < Ws)
            d = depth[b].to(torch.int64)[sy.clamp(0, Hs - 1)][:, sx.clamp(0, Ws - 1)]
            d = d * (oky.view(-1, 1) & okx.view(1, -1))
            sums[b, 0], sums[b, 1] = int(d.sum()), int((d > 0).sum())         # (sums are flip-invariant)
            cnt = int(sums[b, 1])
            mean = torch.tensor(float(int(sums[b, 0])) / cnt / 1000.0 if cnt else 0.0, dtype=torch.float64).to(torch.float32)
            if not hd:
                d = torch.zeros_like(d)
            m = d > 0
            dn = torch.where(m, d.to(torch.float32) / 1000.0 - mean, torch.zeros((), device=dev))
            x[b, 3], x[b, 4], x[b, 5] = dn, dn, dn
            mask[b] = m.to(mask.dtype)
            # bilinear RGB, fp32, the same operation order as the kernel
            f32 = torch.float32
            fy = torch.clamp((yy.to(f32) + 0.5) * (torch.tensor(h, dtype=f32) / torch.tensor(R, dtype=f32)) - 0.5, 0.0, float(h - 1))
            fx = torch.clamp((xs.to(f32) + 0.5) * (torch.tensor(w, dtype=f32) / torch.tensor(R, dtype=f32)) - 0.5, 0.0, float(w - 1))
            y0, x0 = fy.to(torch.int64), fx.to(torch.int64)
            y1, x1 = torch.clamp(y0 + 1, max=h - 1), torch.clamp(x0 + 1, max=w - 1)
            wy, wx = fy - y0.to(f32), fx - x0.to(f32)
            img = rgb[b].to(f32)
            acc = torch.zeros(R, R, 3, device=dev, dtype=f32)
            for t in range(4):
                ys, xs_ = i + (y1 if t & 2 else y0), j + (x1 if t & 1 else x0)
                wt = ((wy if t & 2 else 1.0 - wy).view(-1, 1) * (wx if t & 1 else 1.0 - wx).view(1, -1)).to(f32)
                ok = ((ys >= 0) & (ys < Hs)).view(-1, 1) & ((xs_ >= 0) & (xs_ < Ws)).view(1, -1)
                px = img[ys.clamp(0, Hs - 1)][:, xs_.clamp(0, Ws - 1)]
                acc = torch.where(ok.unsqueeze(-1), torch.addcmul(acc, wt.unsqueeze(-1), px), acc)
            x[b, 0:3] = ((acc / 255.0 - mean_c) / std_c).permute(2, 0, 1).to(x.dtype)
        return 0

    # ---------------------------------------------------------------- seg_head.cu
    def l2norm_max_fwd(self, m1, m2, P, C, out, inv1, inv2):
        a = m1.reshape(P, C)
        i1 = 1.0 / a.norm(dim=1).clamp(min=1e-12)
        o = a * i1[:, None]
        inv1.reshape(-1)[:P] = i1
        if m2 is not None:
            b = m2.reshape(P, C)
            i2 = 1.0 / b.norm(dim=1).clamp(min=1e-12)
            o = torch.maximum(o, b * i2[:, None])
            inv2.reshape(-1)[:P] = i2
        out.reshape(P, C).copy_(o)
        return 0

    def l2norm_max_bwd(self, dout, m1, m2, inv1, inv2, P, C, gscale, d1, d2, accumulate):
        g = dout.reshape(P, C) * gscale
        i1 = inv1.reshape(-1)[:P, None]
        n1 = m1.reshape(P, C) * i1
        g1, g2 = g, None
        if m2 is not None:
            i2 = inv2.reshape(-1)[:P, None]
            n2 = m2.reshape(P, C) * i2
            second = n2 > n1
            g1, g2 = torch.where(second, torch.zeros_like(g), g), torch.where(second, g, torch.zeros_like(g))
        r1 = i1 * (g1 - n1 * (g1 * n1).sum(1, keepdim=True))
        d1.reshape(P, C).copy_(d1.reshape(P, C) + r1 if accumulate else r1)
        if m2 is not None:
            r2 = i2 * (g2 - n2 * (g2 * n2).sum(1, keepdim=True))
            d2.reshape(P, C).copy_(d2.reshape(P, C) + r2 if accumulate else r2)
        return 0

    def seg_ce_fwd(self, logits, label, cw, P, Cn, ignore_index, acc, out2):
        l = logits.reshape(P, Cn)
        y = label.reshape(P)
        valid = (y != ignore_index) & (y >= 0) & (y < Cn)
        yc = torch.where(valid, y, torch.zeros_like(y))
        w = (cw[yc] if cw is not None else torch.ones(P, dtype=l.dtype, device=l.device)) * valid
        nll = torch.logsumexp(l, 1) - l.gather(1, yc[:, None])[:, 0]
        acc.zero_()
        acc[0] = (w * nll).double().sum()
        acc[1] = w.double().sum()
        acc[2] = (l.argmax(1) == y).double().sum()
        out2[0] = (acc[0] / acc[1]) if float(acc[1]) > 0 else 0.0
        out2[1] = acc[2] / P
        return 0

    def seg_ce_bwd(self, logits, label, cw, P, Cn, ignore_index, acc, gscale, dlogits):
        l = logits.reshape(P, Cn)
        y = label.reshape(P)
        valid = (y != ignore_index) & (y >= 0) & (y < Cn)
        yc = torch.where(valid, y, torch.zeros_like(y))
        w = (cw[yc] if cw is not None else torch.ones(P, dtype=l.dtype, device=l.device)) * valid
        k = gscale / float(acc[1]) if float(acc[1]) > 0 else 0.0
        d = torch.softmax(l, 1) - F.one_hot(yc, Cn).to(l.dtype)
        dlogits.reshape(P, Cn).copy_(d * (w * k)[:, None])
        return 0

    # ---------------------------------------------------------------- pointnet2.cu (statement: oracle/pn2_oracle.py, numpy)
    def pn2_furthest_point_sampling(self, xyz, B, N, M, idx):
        from oracle import pn2_oracle as PO
        idx.copy_(torch.from_numpy(PO.furthest_point_sampling(xyz.detach().cpu().numpy().reshape(B, N, 3), M)))
        return 0

    def pn2_ball_query(self, new_xyz, xyz, B, N, M, radius, nsample, idx):
        from oracle import pn2_oracle as PO
        idx.copy_(torch.from_numpy(PO.ball_query(radius, nsample, xyz.detach().cpu().numpy(), new_xyz.detach().cpu().numpy())))
        return 0

    def pn2_three_nn(self, unknown, known, B, n, m, dist2, idx):
        from oracle import pn2_oracle as PO
        d, i = PO.three_nn(unknown.detach().cpu().numpy(), known.detach().cpu().numpy())
        dist2.copy_(torch.from_numpy(d))
        idx.copy_(torch.from_numpy(i))
        return 0

    def pn2_three_interpolate(self, points, idx, weight, B, C, m, n, out):
        from oracle import pn2_oracle as PO
        out.copy_(torch.from_numpy(PO.three_interpolate(points.detach().cpu().numpy(), idx.cpu().numpy(), weight.detach().cpu().numpy())))
        return 0

    def pn2_three_interpolate_grad(self, grad_out, idx, weight, B, C, n, m, grad_points):
        from oracle import pn2_oracle as PO
        g = grad_out.detach().cpu().numpy()[:, :, :, None] * weight.detach().cpu().numpy()[:, None, :, :]
        grad_points.add_(torch.from_numpy(PO.scatter_add(g.reshape(B, C, n * 3), idx.cpu().numpy().reshape(B, n * 3), m)).to(grad_points.dtype))
        return 0

    def pn2_group_points(self, points, idx, B, C, N, npoint, nsample, out):
        from oracle import pn2_oracle as PO
        out.copy_(torch.from_numpy(PO.group_points(points.detach().cpu().numpy(), idx.cpu().numpy())))
        return 0

    def pn2_group_points_grad(self, grad_out, idx, B, C, N, npoint, nsample, grad_points):
        from oracle import pn2_oracle as PO
        g = PO.scatter_add(grad_out.detach().cpu().numpy().reshape(B, C, npoint * nsample), idx.cpu().numpy().reshape(B, -1), N)
        grad_points.add_(torch.from_numpy(g).to(grad_points.dtype))
        return 0

    def pn2_gather_points(self, points, idx, B, C, N, npoint, out):
        from oracle import pn2_oracle as PO
        out.copy_(torch.from_numpy(PO.gather_points(points.detach().cpu().numpy(), idx.cpu().numpy())))
        return 0

    def pn2_gather_points_grad(self, grad_out, idx, B, C, N, npoint, grad_points):
        from oracle import pn2_oracle as PO
        g = PO.scatter_add(grad_out.detach().cpu().numpy(), idx.cpu().numpy(), N)
        grad_points.add_(torch.from_numpy(g).to(grad_points.dtype))
        return 0

    def joint_stats(self, Lr, Ld, vis, use_depth, B, J, rs, lse, fin):
        r = rs.reshape(B, 2, 3)
        ls = lse.reshape(B, 2, J)
        for which, Lx in ((0, Lr.reshape(B, J, J)), (1, Ld.reshape(B, J, J))):
            l = torch.logsumexp(Lx, 1)                       # over k, per pixel-joint j
            ls[:, which] = l
            valid = vis.reshape(B, J) != 0
            if which == 1 and use_depth is not None:
                valid = valid & (use_depth != 0).view(B, 1)
            diag = torch.diagonal(Lx, dim1=1, dim2=2)
            r[:, which, 0] = ((l - diag) * valid).sum(1)
            r[:, which, 1] = valid.sum(1).to(Lx.dtype)
            r[:, which, 2] = ((Lx.argmax(1) == torch.arange(J, device=Lx.device)) & valid).sum(1).to(Lx.dtype)
            cn = r[:, which, 1].sum()
            fin[which] = r[:, which, 0].sum() / cn if cn > 0 else 0.0
            has = r[:, which, 1] > 0
            fin[2 + which] = (r[:, which, 2][has] / r[:, which, 1][has]).mean() if has.any() else 0.0
            fin[4 + which] = cn
        return 0

    def joint_grad(self, Lr, Ld, vis, use_depth, lse, fin, B, J, gscale):
        ls = lse.reshape(B, 2, J)
        eye = torch.eye(J, device=Lr.device).unsqueeze(0)
        for which, Lx in ((0, Lr.reshape(B, J, J)), (1, Ld.reshape(B, J, J))):
            valid = vis.reshape(B, J) != 0
            if which == 1 and use_depth is not None:
                valid = valid & (use_depth != 0).view(B, 1)
            cn = float(fin[4 + which])
            if cn > 0:
                g = gscale / cn * (torch.exp(Lx - ls[:, which].unsqueeze(1)) - eye) * valid.unsqueeze(1)
            else:
                g = torch.zeros_like(Lx)
            Lx.copy_(g)
        return 0

    @staticmethod
    def _scl_masks(B, J, use_rgb, use_depth, dev):
        N = 2 * B * J
        r = torch.arange(N, device=dev)
        ur = use_rgb if use_rgb is not None else torch.ones(B, dtype=torch.long, device=dev)
        ud = use_depth if use_depth is not None else torch.ones(B, dtype=torch.long, device=dev)
        off = torch.cat([(ur == 0).view(B, 1).expand(B, J).reshape(-1), (ud == 0).view(B, 1).expand(B, J).reshape(-1)])
        pos = ((r.view(-1, 1) % J) == (r.view(1, -1) % J)) & (r.view(-1, 1) != r.view(1, -1))
        pos = pos & ~off.view(-1, 1) & ~off.view(1, -1)
        return pos

    def scl_stats(self, Z, B, J, use_rgb, use_depth, rowstat, fin):
        N = 2 * B * J
        Zm = Z.reshape(N, N)
        pos = self._scl_masks(B, J, use_rgb, use_depth, Z.device).to(Z.dtype)
        rs = rowstat.reshape(N, 3)
        rs[:, 0] = torch.logsumexp(Zm, 1)
        rs[:, 1] = pos.sum(1)
        rs[:, 2] = (pos * Zm).sum(1)
        nd = (use_depth != 0).sum() if use_depth is not None else torch.tensor(B)
        on = bool(nd > 0)
        per = -(rs[:, 2] - rs[:, 1] * rs[:, 0]) / rs[:, 1].clamp(min=1)
        fin[0] = per.sum() / N if on else 0.0
        fin[1] = 1.0 if on else 0.0
        return 0

    def scl_grad(self, Z, B, J, use_rgb, use_depth, rowstat, fin, gscale):
        N = 2 * B * J
        Zm = Z.reshape(N, N)
        pos = self._scl_masks(B, J, use_rgb, use_depth, Z.device).to(Z.dtype)
        rs = rowstat.reshape(N, 3)
        np_ = rs[:, 1].unsqueeze(1)
        g = gscale / N * (np_ * torch.exp(Zm - rs[:, 0].unsqueeze(1)) - pos) / np_.clamp(min=1)
        g = g * (np_ > 0) * float(fin[1])
        Zm.copy_(g)
        return 0

    def colsum_finalize(self, part, nparts, C, out, accumulate):
        s = part.reshape(-1)[:nparts * 2 * C].reshape(nparts, 2, C)[:, 0].double().sum(0).to(out.dtype)
        if accumulate:
            out.reshape(-1)[:C].add_(s)
        else:
            out.reshape(-1)[:C].copy_(s)
        return 0

    def colsum_small(self, x, R, C, ld, out, accumulate):
        s = torch.as_strided(x, (R, C), (ld, 1), x.storage_offset()).sum(0)
        if accumulate:
            out.reshape(-1)[:C].add_(s)
        else:
            out.reshape(-1)[:C].copy_(s)
        return 0

    # ---------------------------------------------------------------- sgcn.cu
    def sgcn_adj(self, e, rows, cols, nnz, J, A):
        M = torch.full((J, J), -9e15, device=e.device, dtype=e.dtype)
        M[rows.long(), cols.long()] = e.reshape(-1)
        A.reshape(J, J).copy_(torch.softmax(M, 1))
        return 0

    def sgcn_adj_bwd(self, A, dA, rows, cols, nnz, J, de, accumulate):
        Am, dAm = A.reshape(J, J), dA.reshape(J, J)
        dot = (Am * dAm).sum(1, keepdim=True)
        dM = Am * (dAm - dot)
        v = dM[rows.long(), cols.long()]
        if accumulate:
            de.reshape(-1).add_(v)
        else:
            de.reshape(-1).copy_(v)
        return 0

    def sgcn_aggregate(self, x, A, B, J, Cin, xa):
        Am = A.reshape(J, J)
        xv = x.reshape(B, J, Cin)
        eye = torch.eye(J, device=x.device)
        o = xa.reshape(B, J, 2 * Cin)
        o[:, :, :Cin] = torch.diagonal(Am).view(1, J, 1) * xv
        o[:, :, Cin:] = torch.matmul(Am * (1 - eye), xv)
        return 0

    def sgcn_aggregate_bwd(self, dxa, x, A, B, J, Cin, dx, accumulate, dA):
        Am = A.reshape(J, J)
        g = dxa.reshape(B, J, 2 * Cin)
        eye = torch.eye(J, device=dxa.device)
        if dx is not None:
            v = torch.diagonal(Am).view(1, J, 1) * g[:, :, :Cin] + torch.matmul((Am * (1 - eye)).t(), g[:, :, Cin:])
            if accumulate:
                dx.reshape(B, J, Cin).add_(v)
            else:
                dx.reshape(B, J, Cin).copy_(v)
        if dA is not None:
            xv = x.reshape(B, J, Cin)
            d_off = torch.einsum("bjc,bkc->jk", g[:, :, Cin:], xv)
            d_diag = torch.einsum("bjc,bjc->j", g[:, :, :Cin], xv)
            dA.reshape(J, J).copy_(d_off * (1 - eye) + torch.diag(d_diag))
        return 0

    def joint_mean(self, x, B, J, C, out):
        out.reshape(B, C).copy_(x.reshape(B, J, C).mean(1))
        return 0

    def joint_mean_bwd(self, dout, B, J, C, dx, accumulate):
        v = (dout.reshape(B, 1, C) / J).expand(B, J, C)
        if accumulate:
            dx.reshape(B, J, C).add_(v)
        else:
            dx.reshape(B, J, C).copy_(v)
        return 0

    def sgd_step(self, p, g, buf, n, lr, momentum, wd, first, gscale):
        pv, gv, bv = p.reshape(-1)[:n], g.reshape(-1)[:n], buf.reshape(-1)[:n]
        gg = gv * gscale + wd * pv
        if first:
            bv.copy_(gg)
        else:
            bv.mul_(momentum).add_(gg)
        pv.sub_(lr * bv)
        return 0

    def zero(self, p, nbytes):
        p.reshape(-1).view(torch.uint8)[:nbytes].zero_()
        return 0
