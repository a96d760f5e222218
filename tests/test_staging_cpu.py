"""Input staging for real data, CPU side.

The GPU kernel `hcm_stage_input` is compared with its torch restatement (`kernel_ref.TorchKernels.stage_input`) in
tests/test_kernels_gpu.py::test_stage_input.  HERE the restatement itself is pinned to the calls the reference's loader makes
(pycontrast/datasets/dataset.py:98-160: `TF.resized_crop(..., NEAREST)` of the 16-bit depth frame, `Image.transpose(FLIP_LEFT_RIGHT)`,
mm -> m; :594-602: depth_mask and mean-centring) by executing those torchvision / PIL calls on random crop windows — including
windows that leave the frame — and requiring the mask to be bit-identical and the depth planes equal to fp32 rounding.  The RGB planes
are compared where PIL's bilinear filter is the plain 2x2 one (magnifying crops); PIL rounds the resized image to uint8, the kernel
keeps fp32, hence the one-grey-level tolerance.  Also covers the host logic of `staging.RawFrameStager` (slot rotation)."""
import numpy as np
import PIL
import pytest
import torch
from PIL import Image

from kernel_ref import TorchKernels

TF = pytest.importorskip("torchvision.transforms.functional")


def _reference_sample(rgb, dmm, i, j, h, w, flip, R):
    """dataset.py:98-160 + :594-602 on one decoded sample (numpy uint8 [Hs,Ws,3], uint16 [Hs,Ws])."""
    img = TF.resized_crop(Image.fromarray(rgb), i, j, h, w, (R, R))
    depth = TF.resized_crop(Image.fromarray(dmm), i, j, h, w, (R, R), interpolation=PIL.Image.NEAREST)
    if flip:
        img = img.transpose(Image.FLIP_LEFT_RIGHT)
        depth = depth.transpose(Image.FLIP_LEFT_RIGHT)
    img = torch.from_numpy(np.array(img, dtype=np.float32))
    img /= 255.0
    img -= torch.from_numpy(np.array([0.485, 0.456, 0.406]))
    img /= torch.from_numpy(np.array([0.229, 0.224, 0.225]))
    depth = torch.from_numpy(np.array(depth).astype(np.float32) / 1000.0)
    mask = depth > 0
    mean = depth.sum() / mask.sum()
    nd = depth - mean
    nd[~mask] = 0
    return img.permute(2, 0, 1).float(), nd, mask.float()


def test_restatement_matches_the_reference_loader_calls():
    K = TorchKernels()
    g = torch.Generator().manual_seed(0)
    Hs, Ws, R, B = 424, 512, 128, 6
    for trial in range(4):
        # smooth RGB (bilinear comparable), noisy depth with holes
        yy, xx = torch.meshgrid(torch.arange(Hs), torch.arange(Ws), indexing="ij")
        ph = torch.rand(B, 3, 1, 1, generator=g) * 6.28
        rgb = (127.5 + 120 * torch.sin(yy / 37.0 + xx / 53.0 + ph)).permute(0, 2, 3, 1).round().to(torch.uint8).contiguous()
        dmm = torch.randint(0, 4000, (B, Hs, Ws), generator=g, dtype=torch.int32)
        dmm = (dmm * (torch.rand(B, Hs, Ws, generator=g) > 0.3)).to(torch.uint16)
        crop = torch.zeros(B, 4, dtype=torch.int32)
        for b in range(B):
            if b < 3:      # magnifying windows inside / partly outside the frame
                h, w = int(torch.randint(20, R, (1,), generator=g)), int(torch.randint(20, R, (1,), generator=g))
            else:          # minifying windows (the depth path only is comparable)
                h, w = int(torch.randint(R, Hs + 40, (1,), generator=g)), int(torch.randint(R, Ws + 40, (1,), generator=g))
            i, j = int(torch.randint(-30, Hs - 20, (1,), generator=g)), int(torch.randint(-30, Ws - 20, (1,), generator=g))
            crop[b] = torch.tensor([i, j, h, w])
        flip = torch.randint(0, 2, (B,), generator=g, dtype=torch.int32)
        sums = torch.zeros(B, 2, dtype=torch.int64)
        x = torch.full((B, 6, R, R), float("nan"))
        mask = torch.full((B, R, R), float("nan"))
        K.stage_input(rgb, dmm, crop, flip, None, B, Hs, Ws, R, sums, x, mask)
        for b in range(B):
            i, j, h, w = [int(v) for v in crop[b]]
            img, nd, m = _reference_sample(rgb[b].numpy(), dmm[b].numpy(), i, j, h, w, bool(flip[b]), R)
            assert torch.equal(m, mask[b]), (trial, b, crop[b].tolist())
            assert int(sums[b, 1]) == int(m.sum())
            assert float((nd - x[b, 3]).abs().max()) < 1e-5
            assert torch.equal(x[b, 3], x[b, 4]) and torch.equal(x[b, 3], x[b, 5])
            if b < 3:
                inside = i >= 0 and j >= 0 and i + h <= Hs and j + w <= Ws
                if inside:      # PIL pads an out-of-frame crop BEFORE filtering; the kernel zeroes single taps: compare inside only
                    assert float((img - x[b, :3]).abs().max()) < 1.01 / 255.0 / 0.224, (trial, b)


def test_raw_frame_stager_rotates_slots_on_cpu():
    from hcmoco_b200.staging import RawFrameStager
    K = TorchKernels()
    K.device = "cpu"
    B, Hs, Ws, R = 2, 40, 48, 16
    st = RawFrameStager(K, B, Hs, Ws)
    g = torch.Generator().manual_seed(1)
    outs = []
    for step in range(3):
        rgb = torch.randint(0, 256, (B, Hs, Ws, 3), generator=g, dtype=torch.uint8)
        dmm = torch.randint(1, 3000, (B, Hs, Ws), generator=g, dtype=torch.int32).to(torch.uint16)
        crop = torch.tensor([[2, 3, 30, 30], [0, 0, Hs, Ws]], dtype=torch.int32)
        flip = torch.tensor([step & 1, 0], dtype=torch.int32)
        hd = torch.tensor([1, step != 1], dtype=torch.int64)
        k = st.stage(rgb, dmm, crop, flip, hd)
        assert k == step % 2
        x, mask = torch.zeros(B, 6, R, R), torch.zeros(B, R, R)
        st.consume(k, x, mask)
        assert torch.isfinite(x).all() and float(mask[0].sum()) == R * R
        assert float(mask[1].sum()) == (0 if step == 1 else R * R)      # has_depth = 0 -> empty mask, zero depth planes
        if step == 1:
            assert float(x[1, 3:].abs().max()) == 0.0
        outs.append(x)
    assert not torch.equal(outs[0], outs[2])
