#!/usr/bin/env python
"""Benchmark of the HCMoCo pre-train step (BASELINE.json metric: pre-train triplets/sec).

    python bench.py --gpus N --steps K --warmup W            # this engine (CUDA, one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (oracle)

Workload at N=1: BASELINE.json configs[1] — first-stage sample-level NCE, HRNet-w18 x2 + SemGCN, 256x256,
per-GPU batch 64, K=16384 negatives, bank of 165 894 rows (weak scaling: every rank runs that batch).
A "step" = forward + six NCE losses + backward + memory-bank update + SGD, on synthetic triplets.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "pretrain_triplets_per_sec"
UNIT = "triplets/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (configs[1]: 64)")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--width", type=int, default=18)
    ap.add_argument("--stage", type=int, default=1)
    ap.add_argument("--n-data", type=int, default=165894)
    ap.add_argument("--nce-k", type=int, default=16384)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling aid: after the warm-up run ONE step between cudaProfilerStart/Stop and exit "
                         "(use with `ncu --profile-from-start off`); prints no bench line")
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--detail", default=None, help="write the per-shape conv timing table to this file")
    return ap.parse_args()


def workload_name(a):
    return "stage%d HRNet-w%d+SemGCN, %dx%d, per-GPU batch %d, K=%d, bank n=%d" % (
        a.stage, a.width, a.res, a.res, a.batch, a.nce_k, a.n_data)


# ------------------------------------------------------------------------------------------ FLOP / byte model
def conv_flops_per_image(width, R):
    """Useful conv FLOPs (2*MAC) of ONE HRNet forward on one R x R image, from the layer list itself."""
    from hcmoco_b200 import layout as L
    keys = L.model_keys(width, 1, "mpii")
    ch = L.WIDTHS[width]
    total = 0

    def res_of(k):
        # spatial size of the conv OUTPUT, from its position in the network
        parts = k.split(".")
        if parts[1] == "conv1":
            return R // 2
        if parts[1] in ("conv2", "layer1"):
            return R // 4
        if parts[1].startswith("transition"):
            return (R // 4) >> int(parts[2])
        if parts[1].startswith("stage"):
            if parts[3] == "branches":
                return (R // 4) >> int(parts[4])
            i, j = int(parts[4]), int(parts[5])
            if j > i:
                return (R // 4) >> j                      # 1x1 at the source (low) resolution
            hop = int(parts[6])
            return (R // 4) >> (j + hop + 1)
        raise KeyError(k)

    for k, shp in keys.items():
        if k.startswith("encoder1.") and len(shp) == 4:
            cout, cin, ks, _ = shp
            r = res_of(k)
            total += 2 * r * r * cout * cin * ks * ks
    return total


# ------------------------------------------------------------------------------------------ clocks sampler
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_steps(a, steps, warmup, batch):
    """The reference algorithm (oracle restatement, fp32, all host threads) on a bounded sample of the
    workload: same model, resolution, K and bank, `batch` triplets per step."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from oracle import hcmoco_oracle as O
    from synth import synthetic_banks, synthetic_state
    from hcmoco_b200.synthetic import make_batch, make_dense_idx, make_nce_idx
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    layout = O.model_layout(a.width, a.stage, "mpii")
    P = synthetic_state(layout, 0)
    mom = O.make_momentum(P)
    banks = synthetic_banks(a.n_data, 128, 0)
    d = make_batch(batch, a.res, 16, a.n_data, seed=1234)
    bt = dict(x=d[0], index=d[1], skeleton=d[2], joints_yx=d[4], joints_vis=d[5], use_depth=d[6], depth_mask=d[7])
    nce = make_nce_idx(batch, a.nce_k, a.n_data, d[1])
    dense = make_dense_idx(d[7], a.res // 4, 400)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(P, mom, banks, bt, nce, dense, width=a.width, stage=a.stage, first=(s == 0))
        times.append(time.perf_counter() - t0)
    t = sum(times[warmup:]) / max(1, steps)
    return batch / t, t, cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = min(a.steps, 5), min(a.warmup, 1)
    val, t, cores = cpu_reference_steps(a, steps, warmup, a.cpu_batch)
    sample = "%d timed steps of %d triplets (same model/resolution/K/bank as the workload), %d warm-up" % (
        steps, a.cpu_batch, warmup)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
           "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(a), "sample": sample},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def dense_affinity_roofline(K, hbm_peak, B=32, h=64, S=400, reps=20):
    """Device time of hcm_dense_affinity_{fwd,bwd} alone on configs[2]'s per-GPU shape; algorithmic bytes =
    2*S*128*4 gathered features per depth-bearing triplet."""
    g = torch.Generator().manual_seed(7)
    G1 = torch.randn(B, h * h, 128, generator=g).cuda()
    G2 = torch.randn(B, h * h, 128, generator=g).cuda()
    pix = torch.randint(0, h * h, (B, S), generator=g).cuda()
    kept = torch.ones(B, device="cuda")
    use_depth = torch.ones(B, dtype=torch.int64, device="cuda")
    stat, fin = torch.zeros(B, 2, S, 4, device="cuda"), torch.zeros(8, device="cuda")
    d1, d2 = torch.zeros_like(G1), torch.zeros_like(G2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2: evict the maps between launches
    out = {}
    for name, fn in (("dense_affinity_fwd", lambda: K.dense_affinity_fwd(G1, G2, pix, kept, use_depth, B, S, h, 128, 1 / 0.07, stat, fin)),
                     ("dense_affinity_bwd", lambda: K.dense_affinity_bwd(G1, G2, pix, stat, kept, fin, B, S, h, 128, 1 / 0.07, 1.0, d1, d2))):
        fn()
        ms = 0.0
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        ms /= reps
        nbytes = B * 2 * S * 128 * 4
        flops = 2.0 * B * 2 * S * S * 128 * (1 if name.endswith("fwd") else 2)      # both L and L^T strips (+ dX = G*Y)
        out[name] = {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": nbytes, "us_per_launch": ms * 1e3,
                     "useful_tflops": flops / (ms * 1e-3) / 1e12,
                     "shape": "B=%d depth-bearing triplets, %dx%d maps, S=%d, L2 flushed between launches" % (B, h, h, S),
                     "note": "AI = S/4 = 100 FLOP/B (x2 for both softmax directions): the 3-pass split-bf16 contraction and the "
                             "exp-heavy epilogue bound it, not HBM (DESIGN.md 3.2)"}
    return out


# ------------------------------------------------------------------------------------------ engine arm
def run_engine(a):
    import torch.distributed as dist
    from hcmoco_b200.kernels import CudaKernels
    from hcmoco_b200.pretrain import PretrainStep
    from hcmoco_b200.synthetic import make_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout for the ONE JSON line: NCCL prints its version (and, with NCCL_DEBUG=INFO, its topology) to stdout while
        # the communicator is created — point fd 1 at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    K = CudaKernels()
    step = PretrainStep(K, width=a.width, stage=a.stage, B=a.batch, R=a.res, n_data=a.n_data, nce_k=a.nce_k,
                        world_size=world, rank=rank, use_graph=not a.no_graph, seed=0)
    # synthetic triplets: 2 distinct host batches in pinned memory (e2e) and their device copies (device arm)
    host = [make_batch(a.batch, a.res, 16, a.n_data, seed=1234 + rank + 17 * i, pin=True) for i in range(2)]
    dev = [[t.cuda(non_blocking=True) for t in b] for b in host]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, n, read_back):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pending = None
        for i in range(n):
            # host batches: the next step's H2D copy is started on a copy stream before this step is enqueued, and every
            # step's losses / accuracies are read back (D2H into pinned memory) one step later, so the host never idles the GPU
            step.run(batches[i % len(batches)], next_batch=batches[(i + 1) % len(batches)] if read_back else None)
            if read_back:
                if pending is not None:
                    pending()
                pending = step.results_async()
        if pending is not None:
            pending()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    for i in range(a.warmup):
        step.run(dev[i % 2])
    if a.ncu_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step.run(dev[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    calls0 = K.launches
    ms_dev = timed(dev, a.steps, False)
    launches = step.launches_per_step * a.steps
    ms_e2e = timed(host, a.steps, True)
    clk = clocks.stop() if rank == 0 else None
    h2d = sum(t.numel() * t.element_size() for i, t in enumerate(host[0]) if i in (0, 1, 2, 4, 5, 6, 7))
    d2h = 32 * 4
    # per-kernel-family device time inside one real step (events around every C-ABI launch, no graph)
    fam = step.profile_families(dev[0])
    if a.detail and rank == 0:
        with open(a.detail, "w") as f:
            for k, v in sorted(step.detail.items(), key=lambda kv: -kv[1]["ms"]):
                f.write("%9.3f ms %4d calls %8.1f us/call  %s\n" % (v["ms"], v["calls"], 1e3 * v["ms"] / v["calls"], k))
    step_ms = ms_dev / a.steps
    value = a.batch * world * a.steps / (ms_dev / 1e3)
    e2e = a.batch * world * a.steps / (ms_e2e / 1e3)
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    enc_flops = conv_flops_per_image(a.width, a.res)
    conv_useful = 3 * 2 * enc_flops * a.batch           # fwd + dgrad + wgrad, two encoders, per rank per step
    conv_names = ("conv2d", "tc_conv", "tc_wgrad", "tc_dgrad", "_run_packs", "tc_pack")   # conv kernels + their weight packing
    conv_ms = sum(v["ms"] for k, v in fam.items() if k.startswith(conv_names))
    tot_ms = sum(v["ms"] for v in fam.values())
    # ---- roofline of the DOMINANT KERNEL: the (kernel, shape) with the largest summed device time inside one real step
    dom_shape, dom = max(step.detail.items(), key=lambda kv: kv[1]["ms"])
    parts = dom_shape.split()                            # e.g. "tc_conv 64x64 18->18 k3 s1"
    hh, ww = [int(v) for v in parts[1].split("x")]
    cin, cout = [int(v) for v in parts[2].split("->")]
    ks = int(parts[3][1:]) or 2
    dom_us = 1e3 * dom["ms"] / dom["calls"]
    dom_flops = 2.0 * a.batch * hh * ww * cin * cout * ks * ks
    dom_bytes = 4.0 * a.batch * hh * ww * (cin + cout)                      # fp32 activations in + out (weights are KBs)
    ridge = tc_peak * 1e3 / hbm_peak                                          # FLOP/B
    hbm_bound = dom_flops / dom_bytes < ridge
    ncu_traffic = {}                                                         # per-launch DRAM bytes from the committed ncu captures
    try:
        ncu_traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    tr = ncu_traffic.get("%s B=%d" % (dom_shape, a.batch))
    roof = {"kernel": dom_shape + " (B=%d; %d launches per step)" % (a.batch, dom["calls"]),
            "bound": "hbm" if hbm_bound else "tensor",
            "achieved": (dom_bytes / (dom_us * 1e-6) / 1e9) if hbm_bound else dom_flops / (dom_us * 1e-6) / 1e12,
            "peak": hbm_peak if hbm_bound else tc_peak, "unit": "GB/s" if hbm_bound else "TFLOP/s",
            "traffic": None if tr is None else tr["dram_bytes"],
            "traffic_source": None if tr is None else tr["source"],
            "peak_source": peak_src + (", copy bandwidth" if hbm_bound else ", bf16 dense sustained"),
            "arithmetic_intensity_flop_per_byte": dom_flops / dom_bytes, "ridge_flop_per_byte": ridge,
            "algorithmic_bytes_per_launch": dom_bytes, "algorithmic_flops_per_launch": dom_flops,
            "us_per_launch": dom_us, "share_of_step": dom["ms"] / tot_ms,
            "useful_tflops": dom_flops / (dom_us * 1e-6) / 1e12}
    roof["frac"] = roof["achieved"] / roof["peak"]
    other = {"conv_family": {"kernel": "tc_conv / tc_wgrad / tc_dgrad_s2 (tcgen05) + SIMT stem + weight packing", "bound": "tensor",
                             "achieved": conv_useful / (conv_ms * 1e-3) / 1e12, "peak": tc_peak, "unit": "TFLOP/s",
                             "frac": conv_useful / (conv_ms * 1e-3) / 1e12 / tc_peak, "share_of_step": conv_ms / tot_ms,
                             "algorithmic_flops_per_step": conv_useful, "kernel_ms_per_step": conv_ms,
                             "note": "useful FLOPs only: the bf16 hi/lo split passes are not counted"}}
    # north-star KPI kernel 1: the widest stage-4 conv (3x3, C4 -> C4 at R/32)
    from hcmoco_b200 import layout as L
    c4, r32 = L.WIDTHS[a.width][-1], a.res // 32
    kpi = step.detail.get("tc_conv %dx%d %d->%d k3 s1" % (r32, r32, c4, c4))
    if kpi:
        us = 1e3 * kpi["ms"] / kpi["calls"]
        fl = 2.0 * a.batch * r32 * r32 * c4 * c4 * 9
        other["stage4_conv"] = {"kernel": "tc_conv %dx%d %d->%d k3 s1 (forward + data gradient launches)" % (r32, r32, c4, c4),
                                "bound": "tensor", "achieved": fl / (us * 1e-6) / 1e12, "peak": tc_peak, "unit": "TFLOP/s",
                                "frac": fl / (us * 1e-6) / 1e12 / tc_peak, "us_per_launch": us,
                                "algorithmic_flops_per_launch": fl}
    nce_bytes = 3 * (a.nce_k + 1) * 128 * 4 * a.batch
    for name in ("nce_logits", "nce_bwd"):
        if name in fam:
            g = nce_bytes / (fam[name]["ms"] * 1e-3) / 1e9
            tr = ncu_traffic.get("%s B=%d" % (name, a.batch))
            other[name] = {"bound": "hbm", "achieved": g, "peak": hbm_peak, "unit": "GB/s", "frac": g / hbm_peak,
                           "algorithmic_bytes": nce_bytes, "ms": fam[name]["ms"],
                           "traffic": None if tr is None else tr["dram_bytes"]}
    # north-star KPI kernel 2: the fused dense-affinity kernels on the second-stage per-GPU shape (B=32, 64x64 maps, S=400)
    other.update(dense_affinity_roofline(K, hbm_peak))
    nce = other
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": workload_name(a), "global_batch": a.batch * world, "parallelism": "dp%d" % world,
                      "l2": "working set per step (tens of GB of activations) exceeds the 126 MB L2; no flush needed",
                      "cuda_graph": not a.no_graph},
           "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": ms_e2e / a.steps},
           "gpu_launches": launches, "clocks": clk, "roofline": roof, "roofline_other": nce,
           "kernel_families_ms": {k: round(v["ms"], 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
           "kernel_families_calls": {k: v["calls"] for k, v in fam.items()}}
    if not a.no_cpu_baseline and world == 1:
        val, t, cores = cpu_reference_steps(a, 3, 1, a.cpu_batch)
        out["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "3 timed steps of %d triplets, same model/resolution/K/bank, 1 warm-up"
                                         % a.cpu_batch}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)
