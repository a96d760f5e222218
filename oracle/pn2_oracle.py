"""CPU restatement (numpy, fp32) of the reference's PointNet++ primitives.  TEST INFRASTRUCTURE ONLY.

Each function follows the kernel of the same purpose in /root/reference/pycontrast/networks/pointnet2/src (cited per function) and
the Python wrapper that calls it (networks/pointnet2/pointnet2_utils.py).  The reference ships no tests for these ops and its kernels
only run on a GPU, so this restatement is pinned on the GPU box: tests/test_pointnet2_gpu.py runs the reference kernels themselves
(oracle/_ref/libpn2_ref.so, built from the reference sources by oracle/build_ref.py) against it and against hcm_pn2_*.
fp32 arithmetic is spelled out with the contraction nvcc applies to `a*b + c*d + e*f`: fma(e, f, fma(a, b, c*d)) — the SECOND product
is the one rounded on its own (checked in the SASS of both builds: FMUL of the middle term, then two FFMAs) — so that the
comparison is bit-exact, not approximate."""
import numpy as np

f32 = np.float32


def _fma(a, b, c):
    """fp32 fused multiply-add: exact product and sum in fp64 (24-bit x 24-bit products fit), one rounding."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def sqdist(a, b):
    """(ax-bx)*(ax-bx) + (ay-by)*(ay-by) + (az-bz)*(az-bz) as nvcc compiles it; a, b [...,3] float32."""
    d = (a.astype(f32) - b.astype(f32)).astype(f32)
    r = (d[..., 1] * d[..., 1]).astype(f32)
    r = _fma(d[..., 0], d[..., 0], r)
    return _fma(d[..., 2], d[..., 2], r)


def _bitrev(x, bits):
    r = np.zeros_like(x)
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def furthest_point_sampling(xyz, M):
    """sampling_gpu.cu:97-205 + pointnet2_utils.py:12-29.  xyz [B,N,3] -> idx [B,M] int32.  Starts at point 0, distances start at
    1e10; the winner among equal distances is the one the reference's thread layout picks: T = largest power of two <= min(N, 1024)
    threads, thread t scans k = t, t+T, .. keeping its first maximum, the shared-memory tree keeps the lower position on ties,
    i.e. the smaller bit-reversed thread index."""
    xyz = np.asarray(xyz, dtype=f32)
    B, N, _ = xyz.shape
    T = 1
    while 2 * T <= N and 2 * T <= 1024:
        T *= 2
    bits = T.bit_length() - 1
    k = np.arange(N)
    rank = _bitrev(k % T, bits).astype(np.int64) * (N + 1) + k          # total order among equal values
    out = np.zeros((B, M), dtype=np.int32)
    for b in range(B):
        md = np.full(N, 1e10, dtype=f32)
        old = 0
        for j in range(1, M):
            md = np.minimum(sqdist(xyz[b], xyz[b, old]), md)
            cand = np.flatnonzero(md == md.max())
            old = int(cand[np.argmin(rank[cand])])
            out[b, j] = old
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    """ball_query_gpu.cu:9-46 + pointnet2_utils.py:203-221: first nsample in-radius points in index order, padded with the first;
    all zeros when none."""
    xyz, new_xyz = np.asarray(xyz, dtype=f32), np.asarray(new_xyz, dtype=f32)
    B, M, _ = new_xyz.shape
    r2 = f32(radius) * f32(radius)
    out = np.zeros((B, M, nsample), dtype=np.int32)
    for b in range(B):
        for q in range(M):
            hit = np.flatnonzero(sqdist(new_xyz[b, q][None], xyz[b]) < r2)[:nsample]
            if len(hit):
                out[b, q, :] = hit[0]
                out[b, q, :len(hit)] = hit
    return out


def three_nn(unknown, known):
    """interpolate_gpu.cu:9-49: squared distance and index of the three nearest known points, strict `<` in index order."""
    unknown, known = np.asarray(unknown, dtype=f32), np.asarray(known, dtype=f32)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.full((B, n, 3), np.inf, dtype=f32)
    idx = np.zeros((B, n, 3), dtype=np.int32)
    for b in range(B):
        d = sqdist(unknown[b][:, None, :], known[b][None, :, :])          # [n, m]
        order = np.argsort(d, axis=1, kind="stable")[:, :3]               # stable: equal distances keep index order, as strict `<`
        k = order.shape[1]
        idx[b, :, :k] = order
        d2[b, :, :k] = np.take_along_axis(d, order, 1)
    return d2, idx


def three_interpolate(points, idx, weight):
    """interpolate_gpu.cu:70-86: out[b,c,p] = w0*f[i0] + w1*f[i1] + w2*f[i2]."""
    points, weight = np.asarray(points, dtype=f32), np.asarray(weight, dtype=f32)
    B, C, m = points.shape
    g = [np.take_along_axis(points, np.broadcast_to(idx[:, None, :, j], (B, C, idx.shape[1])).astype(np.int64), 2) for j in range(3)]
    w = [np.broadcast_to(weight[:, None, :, j], g[0].shape) for j in range(3)]
    r = (w[1] * g[1]).astype(f32)
    r = _fma(w[0], g[0], r)
    return _fma(w[2], g[2], r)


def group_points(points, idx):
    """group_points_gpu.cu:44-62: out[b,c,p,s] = points[b,c,idx[b,p,s]]."""
    B, C, N = points.shape
    P, S = idx.shape[1], idx.shape[2]
    flat = np.broadcast_to(idx.reshape(B, 1, P * S), (B, C, P * S)).astype(np.int64)
    return np.take_along_axis(np.asarray(points), flat, 2).reshape(B, C, P, S)


def gather_points(points, idx):
    """sampling_gpu.cu:9-26: out[b,c,p] = points[b,c,idx[b,p]]."""
    B, C, N = points.shape
    return np.take_along_axis(np.asarray(points), np.broadcast_to(idx[:, None, :], (B, C, idx.shape[1])).astype(np.int64), 2)


def scatter_add(grad_out, idx, N):
    """The three gradient kernels (atomicAdd scatters): grad_points[b,c,idx[b,e]] += grad_out[b,c,e]; summed in fp64 here."""
    B, C, E = grad_out.shape
    out = np.zeros((B, C, N), dtype=np.float64)
    for b in range(B):
        np.add.at(out[b], (slice(None), idx[b].reshape(-1).astype(np.int64)), grad_out[b].astype(np.float64))
    return out
