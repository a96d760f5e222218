"""Synthetic (RGB, depth, 2D-keypoint) triplets in the reference's batch-tuple layout.

The metric is quoted on synthetic data; there are no datasets in the image.
Layout follows what the reference's DataLoader yields and the step loops index
(datasets/dataset.py:614-617, learning/contrast_trainer.py:553-556, 925-931):

  data[0] rgbd   [B,6,R,R] f32  ch0-2 RGB, ch3-5 mean-centred depth (replicated), 0 outside mask
  data[1] index  [B] i64        row of the memory bank
  data[2] joints [B,J,2] f32    normalised 2D key-points, zero where invisible
  data[3] joints3d [B,25,3]     unused on the path
  data[4] pixel joints (y,x) [B,J,2] f32
  data[5] joints_vis [B,J] i32
  data[6] true_depth [B] i64    1 = sample carries real depth  (the 50/50 NTU / MPII mix, F11)
  data[7] depth_mask [B,R,R] f32
  data[8] scale [B] f32         unused on the path
"""
import torch


def make_batch(B, R, J=16, n_data=165894, seed=1234, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.randn(B, 3, R, R, generator=g)
    use_depth = (torch.rand(B, generator=g) < 0.5).long()
    if use_depth.sum() == 0:
        use_depth[0] = 1
    yy, xx = torch.meshgrid(torch.arange(R, dtype=torch.float32), torch.arange(R, dtype=torch.float32),
                            indexing="ij")
    ell = (((yy - R / 2) / (0.35 * R)) ** 2 + ((xx - R / 2) / (0.2 * R)) ** 2 <= 1.0).float()
    mask = ell.unsqueeze(0) * use_depth.view(B, 1, 1).float()
    depth = mask * (0.3 * torch.randn(B, R, R, generator=g))
    rgbd = torch.cat([rgb, depth.unsqueeze(1).expand(-1, 3, -1, -1)], 1).contiguous()
    index = torch.randperm(n_data, generator=g)[:B].clone()
    vis = (torch.rand(B, J, generator=g) < 0.85).int()
    first_depth = int(use_depth.argmax())
    if vis[first_depth].sum() == 0:
        vis[first_depth, 0] = 1
    joints = (torch.rand(B, J, 2, generator=g) * 2 - 1) * vis.unsqueeze(-1).float()
    py = R / 2 + (torch.rand(B, J, generator=g) * 2 - 1) * 0.35 * R
    px = R / 2 + (torch.rand(B, J, generator=g) * 2 - 1) * 0.2 * R
    pix = torch.stack([py, px], -1) * vis.unsqueeze(-1).float()
    joints3d = torch.zeros(B, 25, 3)
    scale = torch.ones(B)
    data = [rgbd, index, joints, joints3d, pix, vis, use_depth, mask, scale]
    if pin:
        data = [t.pin_memory() for t in data]
    if device != "cpu":
        data = [t.to(device, non_blocking=True) for t in data]
    return data


def make_nce_idx(B, K, n_data, index, seed=99, device="cpu"):
    """idx [B,K+1] uniform over the bank with column 0 = the sample's own row
    (memory/mem_bank.py:176-177; AliasMethod with uniform probs is a uniform draw)."""
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, n_data, (B, K + 1), generator=g)
    idx[:, 0] = index.cpu()
    return idx.to(device)


def make_dense_idx(depth_mask, h, S, seed=7, device="cpu"):
    """[B,S] pixel ids drawn with replacement from each sample's nearest-resized mask
    (learning/contrast_trainer.py:671-685).  Rows of samples with an empty mask are zeros."""
    g = torch.Generator().manual_seed(seed)
    B, R = depth_mask.shape[0], depth_mask.shape[-1]
    step = R // h
    m = depth_mask.cpu()[:, ::step, ::step][:, :h, :h].reshape(B, -1)
    out = torch.zeros(B, S, dtype=torch.long)
    for b in range(B):
        if m[b].sum() > 0:
            out[b] = torch.multinomial(m[b], S, replacement=True, generator=g)
    return out.to(device)
