#!/usr/bin/env python
"""Benchmark of the HCMoCo pre-train step (BASELINE.json metric: pre-train triplets/sec).

    python bench.py --gpus N --steps K --warmup W            # this engine (CUDA, one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (oracle)

Workload at N=1: BASELINE.json configs[1] — first-stage sample-level NCE, HRNet-w18 x2 + SemGCN, 256x256,
per-GPU batch 64, K=16384 negatives, bank of 165 894 rows (weak scaling: every rank runs that batch).
A "step" = forward + six NCE losses + backward + memory-bank update + SGD, on synthetic triplets.
Prints ONE JSON line (rank 0).  Besides the contract keys the line carries
  roofline            the dominant kernel (median of 5 warmed, event-timed passes inside a real step)
  roofline_other      the conv family, the north-star KPI kernels (stage-4 conv, dense affinity), the NCE kernels
  extra_configs       the second-stage / w32 configurations of BASELINE.json (configs[2..4], per-GPU shapes) timed the same way
  gpu_eager_baseline  the reference ALGORITHM (oracle statement) under PyTorch eager + cuDNN on this GPU, TF32 on / off
                      ("the existing Blackwell library path")
  cpu_baseline        the same algorithm on the host cores
"""
import argparse
import gc
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
DOMINANT = "tc_conv 64x64 18->18 k3 s1"       # fallback name only; the kernel is picked by median time (see run_engine)


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
    ap.add_argument("--skeleton", default="mpii")
    ap.add_argument("--n-data", type=int, default=165894)
    ap.add_argument("--nce-k", type=int, default=16384)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs legs (configs[2..4])")
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling aid: after the warm-up run ONE step between cudaProfilerStart/Stop and exit "
                         "(use with `ncu --profile-from-start off`); prints no bench line")
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--detail", default=None, help="write the per-shape conv timing table to this file")
    return ap.parse_args()


def workload_name(a):
    return "stage%d HRNet-w%d+SemGCN, %dx%d, per-GPU batch %d, K=%d, bank n=%d" % (
        a.stage, a.width, a.res, a.res, a.batch, a.nce_k, a.n_data)


def config_of(a, world):
    """The `config` object: identical for the engine arm and the reference arm (the driver compares them)."""
    return {"workload": workload_name(a), "global_batch": a.batch * world, "parallelism": "dp%d" % world,
            "l2": "working set per step (tens of GB of activations) exceeds the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------ FLOP / byte model
def conv_flops_per_image(width, R):
    """Useful conv FLOPs (2*MAC) of ONE HRNet forward on one R x R image, from the layer list itself."""
    from hcmoco_b200 import layout as L
    keys = L.model_keys(width, 1, "mpii")
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


# ------------------------------------------------------------------------------------------ the reference algorithm (oracle)
def _oracle_problem(a, batch, device="cpu"):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from oracle import hcmoco_oracle as O
    from synth import synthetic_banks, synthetic_state
    from hcmoco_b200.synthetic import make_batch, make_dense_idx, make_nce_idx
    J = 16 if a.skeleton == "mpii" else 13
    layout = O.model_layout(a.width, a.stage, a.skeleton)
    P = synthetic_state(layout, 0)
    P = type(P)((k, v.to(device)) for k, v in P.items())
    mom = O.make_momentum(P)
    banks = [b.to(device) for b in synthetic_banks(a.n_data, 128, 0)]
    d = make_batch(batch, a.res, J, a.n_data, seed=1234)
    bt = dict(x=d[0], index=d[1], skeleton=d[2], joints_yx=d[4], joints_vis=d[5], use_depth=d[6], depth_mask=d[7])
    nce = make_nce_idx(batch, a.nce_k, a.n_data, d[1]).to(device)
    dense = make_dense_idx(d[7], a.res // 4, 400).to(device)
    bt = {k: v.to(device) for k, v in bt.items()}

    def step(first):
        return O.train_step(P, mom, banks, bt, nce, dense, width=a.width, skeleton=a.skeleton, stage=a.stage, first=first)

    return step


def cpu_reference_steps(a, steps, warmup, batch):
    """The reference algorithm (oracle restatement, fp32, all host threads) on a bounded sample of the
    workload: same model, resolution, K and bank, `batch` triplets per step."""
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    step = _oracle_problem(a, batch)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        step(s == 0)
        times.append(time.perf_counter() - t0)
    t = sum(times[warmup:]) / max(1, steps)
    return batch / t, t, cores


def gpu_eager_steps(a, batch, tf32, steps=5, warmup=3):
    """The same algorithm statement under PyTorch eager on THIS GPU: cuDNN convolutions / batch-norm, cuBLAS, ATen, fp32 tensors,
    `cudnn.benchmark=True` as the reference sets it (learning/base_trainer.py:28); tf32=True is PyTorch's shipped cuDNN default
    (the reference as a user runs it), tf32=False is the fp32-exact mode the parity bar is stated in."""
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    step = _oracle_problem(a, batch, "cuda")
    for s in range(warmup):
        step(s == 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        step(False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del step
    gc.collect()
    torch.cuda.empty_cache()
    return {"value": batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": batch, "steps": steps, "warmup": warmup}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = a.steps, a.warmup
    val, t, cores = cpu_reference_steps(a, steps, warmup, a.cpu_batch)
    sample = "each step = %d triplets of the workload (same model/resolution/K/bank), %d timed steps, %d warm-up" % (
        a.cpu_batch, steps, warmup)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
           "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": config_of(a, world),
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
    work = torch.empty((K.dense_affinity_work_bytes(B, S) + 3) // 4, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2: evict the maps between launches
    out = {}
    for name, fn in (("dense_affinity_fwd", lambda: K.dense_affinity_fwd(G1, G2, pix, kept, use_depth, B, S, h, 128, 1 / 0.07, stat, fin, work)),
                     ("dense_affinity_bwd", lambda: K.dense_affinity_bwd(G1, G2, pix, stat, kept, fin, B, S, h, 128, 1 / 0.07, 1.0, 1.0, d1, d2, work, 0))):
        fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        nbytes = B * 2 * S * 128 * 4
        flops = 2.0 * B * 2 * S * S * 128 * (1 if name.endswith("fwd") else 2)      # both L and L^T strips (+ dX = G*Y)
        out[name] = {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": nbytes, "us_per_launch": ms * 1e3,
                     "useful_tflops": flops / (ms * 1e-3) / 1e12,
                     "shape": "B=%d depth-bearing triplets, %dx%d maps, S=%d, L2 flushed between launches" % (B, h, h, S),
                     "note": "AI = S/4 = 100 FLOP/B (x2 for both softmax directions): the split-bf16 contraction and the "
                             "exp-heavy epilogue bound it, not HBM (DESIGN.md 3.2)"}
    del flush
    return out


# ------------------------------------------------------------------------------------------ engine arm
class Runner:
    """One PretrainStep + its synthetic batches; `timed()` is the contract's timed region."""

    def __init__(self, a, K, world, rank, dist):
        from hcmoco_b200.pretrain import PretrainStep
        from hcmoco_b200.synthetic import make_batch
        self.a, self.world, self.dist = a, world, dist
        J = 16 if a.skeleton == "mpii" else 13
        self.step = PretrainStep(K, width=a.width, stage=a.stage, skeleton=a.skeleton, B=a.batch, R=a.res, n_data=a.n_data,
                                 nce_k=a.nce_k, world_size=world, rank=rank, use_graph=not a.no_graph, seed=0)
        # synthetic triplets: 2 distinct host batches in pinned memory (e2e) and their device copies (device arm)
        self.host = [make_batch(a.batch, a.res, J, a.n_data, seed=1234 + rank + 17 * i, pin=True) for i in range(2)]
        self.dev = [[t.cuda(non_blocking=True) for t in b] for b in self.host]
        torch.cuda.synchronize()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, batches, n, read_back):
        step = self.step
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pending = None
        for i in range(n):
            # host batches: the next step's H2D copies are started on a copy stream before this step is enqueued, and every
            # step's losses / accuracies are read back (D2H into pinned memory) one step later, so the host never idles the GPU
            step.run(batches[i % len(batches)], next_batch=batches[(i + 1) % len(batches)] if read_back else None)
            if read_back:
                if pending is not None:
                    pending()
                pending = step.results_async()
        if pending is not None:
            pending()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def _diag(self, what):
        if self.world > 1 and os.environ.get("HCM_BENCH_DIAG"):
            ok = replicas_identical(self.step, self.dist, self.world)
            if int(os.environ.get("RANK", "0")) == 0:
                sys.stderr.write("[diag] %-28s identical=%s %s\n" % (what, ok, replicas_identical.detail))

    def measure(self, steps, warmup):
        a = self.a
        self._diag("after construction")
        for i in range(warmup):
            self.step.run(self.dev[i % 2])
        self._diag("after warm-up")
        ms_dev = self.timed(self.dev, steps, False)
        self._diag("after device-timed steps")
        self.step.run(self.host[0], next_batch=self.host[0])       # e2e warm-up: stager buffers, pinned result slots
        self._diag("after e2e warm-up step")
        ms_e2e = self.timed(self.host, steps, True)
        self._diag("after e2e steps")
        h2d = self.step.h2d_bytes(self.host[0])
        return {"value": a.batch * self.world * steps / (ms_dev / 1e3), "ms_per_step": ms_dev / steps,
                "e2e": {"value": a.batch * self.world * steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 32 * 4, "ms_per_step": ms_e2e / steps},
                "gpu_launches": self.step.launches_per_step * steps}

    def close(self):
        self.step = self.host = self.dev = None
        gc.collect()
        torch.cuda.empty_cache()


def replicas_identical(step, dist, world):
    """Every rank must hold bit-identical parameters, momentum and memory banks after the timed steps (SURVEY.md 8(e);
    contrast_trainer.py:578-579, mem_bank.py:195-199): all-gather an integer checksum of each and compare."""
    e = step.eng
    sums = []
    for t in (e.store.p, e.store.m, e.banks[0], e.banks[1], e.banks[2]):
        v = t.reshape(-1).view(torch.int32).to(torch.int64)
        sums += [v.sum(), (v * (torch.arange(v.numel(), device=v.device) % 8191 + 1)).sum()]
    mine = torch.stack(sums)
    allv = torch.empty(world, mine.numel(), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(allv, mine)
    same = (allv == allv[0:1]).all(0).reshape(5, 2).all(1).tolist()
    detail = dict(zip(("params", "momentum", "memory_1", "memory_2", "memory_3"), same))
    # how far apart (and whether the training state is still finite): max |p - p(rank 0)| over the ranks
    ref = e.store.p.clone()
    dist.broadcast(ref, 0)
    d = torch.stack([(e.store.p - ref).abs().max().nan_to_num(nan=-1.0), (~torch.isfinite(e.store.p)).sum().float(),
                     e.store.p.abs().max().nan_to_num(nan=-1.0)])
    alld = torch.empty(world, 3, device="cuda")
    dist.all_gather_into_tensor(alld, d)
    detail["params_max_abs_diff_vs_rank0"] = float(alld[:, 0].max())
    detail["params_nonfinite"] = int(alld[:, 1].max())
    detail["params_max_abs"] = float(alld[:, 2].max())
    replicas_identical.detail = detail
    return bool(all(same))


def run_engine(a):
    import torch.distributed as dist
    from hcmoco_b200.kernels import CudaKernels

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
    run = Runner(a, K, world, rank, dist)
    step = run.step
    if a.ncu_step:
        for i in range(a.warmup):
            step.run(run.dev[i % 2])
        torch.cuda.synchronize()
        log = os.environ.get("HCM_LAUNCH_LOG")      # conv launches of the profiled step in host order (zip with the ncu list)
        if log:
            names = []
            for nm in ("tc_conv", "tc_wgrad", "tc_dgrad_s2", "conv2d_fwd", "conv2d_wgrad", "conv2d_dgrad"):
                fn = getattr(K, nm)
                o = {"tc_conv": 4, "conv2d_fwd": 4, "conv2d_dgrad": 3, "conv2d_wgrad": 3, "tc_wgrad": 4, "tc_dgrad_s2": 3}[nm]

                def wrap(*args, _fn=fn, _nm=nm, _o=o):
                    nl = 4 // K.tc_dgrad_s2_nqs(*args[3:8]) if _nm == "tc_dgrad_s2" else 1
                    names.append("%s %dx%d %d->%d k%d%s|%d" % (_nm, args[_o + 1], args[_o + 2], args[_o + 3], args[_o + 4], args[_o + 5],
                                                               "" if _nm == "tc_dgrad_s2" else " s%d" % args[_o + 6], nl))
                    return _fn(*args)
                setattr(K, nm, wrap)
            step = run.step = type(step)(K, width=a.width, stage=a.stage, skeleton=a.skeleton, B=a.batch, R=a.res, n_data=a.n_data,
                                         nce_k=a.nce_k, world_size=world, rank=rank, use_graph=False, seed=0)
            step.run(run.dev[0])
            torch.cuda.synchronize()
            names.clear()
        torch.cuda.profiler.start()
        step.run(run.dev[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if log:
            with open(log, "w") as f:
                f.write("\n".join(names) + "\n")
        return
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    main = run.measure(a.steps, a.warmup)
    clk = clocks.stop() if rank == 0 else None
    identical = replicas_identical(step, dist, world) if world > 1 else None
    # per-kernel-family device time inside one real step (events around every C-ABI launch, no graph): 5 warmed passes,
    # per-shape MEDIAN (one pass is noisy: a host hiccup between two event records lands in whatever kernel it hits)
    fams, details = [], []
    for _ in range(5):
        fams.append(step.profile_families(run.dev[0]))
        details.append(step.detail)

    def med(rows, key):
        v = sorted(r[key]["ms"] for r in rows if key in r)
        return v[len(v) // 2]

    fam = {k: {"ms": med(fams, k), "calls": fams[0][k]["calls"]} for k in fams[0]}
    detail = {k: {"ms": med(details, k), "calls": details[0][k]["calls"]} for k in details[0]}
    if a.detail and rank == 0:
        with open(a.detail, "w") as f:
            for k, v in sorted(detail.items(), key=lambda kv: -kv[1]["ms"]):
                f.write("%9.3f ms %4d calls %8.1f us/call  %s\n" % (v["ms"], v["calls"], 1e3 * v["ms"] / v["calls"], k))
    step_ms = main["ms_per_step"]
    if rank != 0:
        run.close()
        extra_configs(a, K, world, rank, dist)        # every rank takes part in the collective legs
        if world > 1:
            dist.destroy_process_group()
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
    # ---- roofline of the DOMINANT KERNEL: the (kernel, shape) with the largest summed median device time inside one real step
    dom_shape, dom = max(detail.items(), key=lambda kv: kv[1]["ms"])
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
            "timing": "median over 5 warmed passes of CUDA events around every launch of this shape inside a real step",
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
    kpi = detail.get("tc_conv %dx%d %d->%d k3 s1" % (r32, r32, c4, c4))
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
    cfg = config_of(a, world)
    cfg["cuda_graph"] = not a.no_graph
    out = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": cfg, "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "clocks": clk,
           "roofline": roof, "roofline_other": other,
           "kernel_families_ms": {k: round(v["ms"], 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
           "kernel_families_calls": {k: v["calls"] for k, v in fam.items()}}
    if identical is not None:
        out["replicas_identical"] = identical
        out["replicas_identical_detail"] = getattr(replicas_identical, "detail", None)
    run.close()
    step = run = None
    ex = extra_configs(a, K, world, rank, dist)
    if ex:
        out["extra_configs"] = ex
    if not a.no_extra and world == 1 and (a.stage, a.width, a.res, a.batch) == (1, 18, 256, 64):
        try:
            out["widened_rows"] = widened_rows(a, K, hbm_peak)
        except Exception as ex_:
            out["widened_rows"] = {"error": repr(ex_)[:300]}
            gc.collect()
            torch.cuda.empty_cache()
    if not a.no_gpu_eager and world == 1:
        eager = {}
        for name, tf32 in (("tf32_on", True), ("tf32_off", False)):
            try:
                eager[name] = gpu_eager_steps(a, a.batch, tf32)
            except Exception as ex_:          # never lose the bench line to the baseline leg
                eager[name] = {"error": repr(ex_)[:200]}
                gc.collect()
                torch.cuda.empty_cache()
        eager["what"] = ("the reference algorithm (oracle statement: F.conv2d / F.batch_norm / autograd / index_select + bmm) "
                         "under PyTorch %s eager on this GPU, fp32 tensors, cudnn.benchmark=True, same batch and shapes as the "
                         "workload; tf32_on = PyTorch's default cuDNN TF32 convolutions (misses the 1e-3 parity bar, SURVEY F8), "
                         "tf32_off = fp32-exact" % torch.__version__)
        out["gpu_eager_baseline"] = eager
    if not a.no_cpu_baseline and world == 1:
        val, t, cores = cpu_reference_steps(a, 3, 1, a.cpu_batch)
        out["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "3 timed steps of %d triplets, same model/resolution/K/bank, 1 warm-up"
                                         % a.cpu_batch}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _median_ms(fn, reps=7):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def widened_rows(a, K, hbm_peak):
    """Measured numbers for the SURVEY.md section 8(f) rows built beside the pre-train path: input staging of
    decoded frames (f1), the segmentation fine-tuning step (f3), farthest point sampling of the PointNet++ primitives (f4)."""
    from types import SimpleNamespace
    from hcmoco_b200.synthetic import make_batch
    out = {}
    g = torch.Generator().manual_seed(0)
    # f1: B=64 NTU-sized frames (424 x 512 uint8 RGB + uint16 depth) -> x [64,6,256,256] + depth_mask
    B, Hs, Ws, R = 64, 424, 512, 256
    rgb = torch.randint(0, 256, (B, Hs, Ws, 3), generator=g, dtype=torch.uint8).cuda()
    depth = torch.randint(500, 4000, (B, Hs, Ws), generator=g, dtype=torch.int32).to(torch.uint16).cuda()
    crop = torch.tensor([[20, 40, 380, 380]] * B, dtype=torch.int32).cuda()
    flip = torch.zeros(B, dtype=torch.int32).cuda()
    sums, x, mask = torch.zeros(B, 2, dtype=torch.int64).cuda(), torch.empty(B, 6, R, R).cuda(), torch.empty(B, R, R).cuda()
    K.stage_input(rgb, depth, crop, flip, None, B, Hs, Ws, R, sums, x, mask)
    ms = _median_ms(lambda: K.stage_input(rgb, depth, crop, flip, None, B, Hs, Ws, R, sums, x, mask))
    nbytes = B * (380 * 380 * 5 + 7 * R * R * 4)             # the crop window's uint8 x3 + uint16, the 6 + 1 fp32 output planes
    out["f1_stage_input"] = {"ms": ms, "triplets_per_s": B / (ms * 1e-3), "achieved_GBps": nbytes / (ms * 1e-3) / 1e9,
                             "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / hbm_peak,
                             "shape": "B=64, 424x512 frames, 380x380 crop -> 256x256 (2 launches)"}
    del rgb, depth, x, mask
    # f3: the fused segmentation fine-tuning step, second-stage model, B=32, every sample labelled, 25 classes
    from hcmoco_b200.api import HCMoCoMem, HCMoCoModel
    from hcmoco_b200.segment import FCNHead, NTU_CLASS_WEIGHTS, SegTrainer
    Bs = 32
    opt = SimpleNamespace(modal="RGBD2S", arch="HRNet", jigsaw=False, head="linear", pool_method="mean", width=18, linear_feat_map=1,
                          skeleton_meta_name="mpii", in_channel_list=[3, 3], feat_dim=128, mem="bank+jointspri3d", nce_k=a.nce_k,
                          nce_t=0.07, nce_m=0.5, temperature=0.07, pri3d_num_samples_per_image=400, modality_missing=1,
                          supervise_type=0, cmc_loss_weights=1, other_loss_weights=1, print_freq=1000, n_class=25, cuda_graph=True)
    model = HCMoCoModel(opt, K)
    mem = HCMoCoMem(128, a.n_data, a.nce_k, 0.07, 0.5, K)
    clf = FCNHead(128, 128, 25, 1, 1, K, NTU_CLASS_WEIGHTS)
    tr = SegTrainer(opt)
    model.attach_memory(mem)
    d = [t.cuda() for t in make_batch(Bs, R, 16, a.n_data, seed=5)]
    label = torch.randint(0, 25, (Bs, R, R), generator=g).cuda()
    data = list(d) + [None] * (11 - len(d))
    data[9], data[10] = label, torch.ones(Bs, dtype=torch.int64).cuda()
    for _ in range(3):
        r = tr.seg_step(model, clf, mem, data, 0.03, 0.9, 1e-4)
    torch.cuda.synchronize()
    n = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = tr.seg_step(model, clf, mem, data, 0.03, 0.9, 1e-4)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res = r()
    out["f3_seg_step"] = {"ms_per_step": ms, "triplets_per_s": Bs / (ms * 1e-3), "seg_loss": float(res["seg_loss"]),
                          "workload": "SegTrainer.seg_step: stage2 HRNet-w18+SemGCN, 256x256, batch 32 (all labelled), 25 classes; "
                                      "engine programs as three CUDA graphs + the FCN head launched eagerly"}
    del model, mem, clf, tr
    gc.collect()
    torch.cuda.empty_cache()
    # f4: farthest point sampling at the Pointnet2MSG sizes (first two set-abstraction levels)
    xyz = torch.randn(Bs, 4096, 3, generator=g).cuda()
    fps = {}
    for M in (4096, 1024):
        idx = torch.zeros(Bs, M, dtype=torch.int32).cuda()
        K.pn2_furthest_point_sampling(xyz, Bs, 4096, M, idx)
        fps["M=%d" % M] = _median_ms(lambda: K.pn2_furthest_point_sampling(xyz, Bs, 4096, M, idx), 3)
    out["f4_fps_ms"] = dict(fps, shape="B=32 clouds of 4096 points")
    return out


def extra_configs(a, K, world, rank, dist):
    """BASELINE.json configs[2..4] at their per-GPU shapes (8-GPU global batches 256 / 256 / 128): the second-stage objectives
    (dense + sparse + SCL, the 1x1 projections) and every HRNet-w32 / 384x384 shape, timed like the main workload."""
    if a.no_extra or (a.stage, a.width, a.res, a.batch) != (1, 18, 256, 64):
        return None
    out = {}
    for name, kw in (("C3_configs[2]", dict(stage=2, width=18, res=256, batch=32, skeleton="mpii")),
                     ("C4_configs[3]", dict(stage=2, width=18, res=256, batch=32, skeleton="coco_reduce")),
                     ("C5_configs[4]", dict(stage=2, width=32, res=384, batch=16, skeleton="mpii"))):
        b = argparse.Namespace(**vars(a))
        for k, v in kw.items():
            setattr(b, k, v)
        steps = max(3, min(a.steps, 10))
        try:
            r = Runner(b, K, world, rank, dist)
            m = r.measure(steps, 3)
            r.close()
            r = None
            m.update(workload=workload_name(b) + ", J=%d" % (16 if b.skeleton == "mpii" else 13), steps=steps, warmup=3,
                     unit=UNIT, n_gpus=world)
            out[name] = m
        except Exception as ex_:
            out[name] = {"error": repr(ex_)[:300]}
            gc.collect()
            torch.cuda.empty_cache()
    return out if rank == 0 else None


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)
