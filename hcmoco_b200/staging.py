"""Input staging for real data: decoded frames in, the step's static input buffers out.

The reference's loader (datasets/dataset.py:104-160, 570-617; 40 CPU workers) resizes, crops, flips and normalises every sample on
the host and ships 6 fp32 planes per triplet.  `RawFrameStager` takes what a decoder produces — uint8 RGB frames, uint16 depth
frames in millimetres (pinned host tensors or device tensors) and the crop / flip parameters the sampler drew — copies the small
integer frames on a copy stream (double-buffered, as `pretrain.InputStager`) and lets `hcm_stage_input` write `engine.x` and
`engine.depth_mask` in place.  Scope: the NTU RGB-D branch (resized crop + flip + normalisation + depth mean-centring + mask);
the MPII / COCO branch (affine warp with rotation) still arrives as host-side tensors with has_depth = 0."""
import torch


class RawFrameStager:
    def __init__(self, K, B, Hs, Ws):
        self.K, self.B, self.Hs, self.Ws = K, B, Hs, Ws
        dev = K.device
        self.rgb = [torch.empty(B, Hs, Ws, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.depth = [torch.empty(B, Hs, Ws, dtype=torch.uint16, device=dev) for _ in range(2)]
        self.crop = [torch.empty(B, 4, dtype=torch.int32, device=dev) for _ in range(2)]
        self.flip = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(2)]
        self.has_depth = [torch.empty(B, dtype=torch.int64, device=dev) for _ in range(2)]
        self.sums = torch.zeros(B, 2, dtype=torch.int64, device=dev)
        self.slot = 0
        cuda = str(dev).startswith("cuda")
        self.stream = torch.cuda.Stream() if cuda else None
        self.ready = [torch.cuda.Event() for _ in range(2)] if cuda else None
        self.free = [torch.cuda.Event() for _ in range(2)] if cuda else None
        if cuda:
            for ev in self.free:
                ev.record()

    def stage(self, rgb, depth, crop, flip, has_depth):
        """Start the copies of one batch of decoded frames (returns the slot to pass to `consume`)."""
        k = self.slot
        self.slot ^= 1
        dst = (self.rgb[k], self.depth[k], self.crop[k], self.flip[k], self.has_depth[k])
        if self.stream is None:
            for d, s in zip(dst, (rgb, depth, crop, flip, has_depth)):
                d.copy_(s)
            return k
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.free[k])
            for d, s in zip(dst, (rgb, depth, crop, flip, has_depth)):
                d.copy_(s, non_blocking=True)
            self.ready[k].record(self.stream)
        return k

    def consume(self, k, x, depth_mask):
        """Enqueue the staging kernels of slot k on the current stream: x [B,6,R,R], depth_mask [B,R,R] (the engine's buffers)."""
        R = x.shape[-1]
        if self.stream is not None:
            cur = torch.cuda.current_stream()
            cur.wait_event(self.ready[k])
        self.K.stage_input(self.rgb[k], self.depth[k], self.crop[k], self.flip[k], self.has_depth[k], self.B, self.Hs, self.Ws, R,
                           self.sums, x, depth_mask)
        if self.stream is not None:
            self.free[k].record(cur)
