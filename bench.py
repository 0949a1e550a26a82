#!/usr/bin/env python
"""bench.py -- RVQ frames/sec on B200 (BASELINE.json metric), roofline, CPU baseline, e2e.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4s|c5q] [--impl reference]

A "step" is one pass of the hot path (rvq_encode [+ EMA statistics, all-reduce, finalize for c3]) over one
batch of synthetic latent frames resident in HBM.  Under torchrun every rank encodes its own shard
(weak scaling, codebooks replicated); the only collective is the all-reduce of the EMA statistics (c3).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nq, K, d, frames per GPU, update_codebook)
    "c2": dict(nq=8, K=1024, d=128, frames=1 << 20, update=False,
               desc="BASELINE configs[1]: RVQ encode only, 8 quantizers x 1024 codes, d=128, 1M synthetic frames"),
    "c3": dict(nq=12, K=1024, d=256, frames=1 << 20, update=True,
               desc="BASELINE configs[2]: RVQ encode + EMA update, 12 x 1024 codes, d=256, 1M frames per GPU"),
    "c3m": dict(nq=12, K=1024, d=256, frames=1 << 20, update=True, som=True, cutoff=1.0,
                desc="c3 plus codebook maintenance (SURVEY 8f rows 2-3): SOM neighbourhood spreading (hard kernel) and "
                     "stale-code re-seeding (vq_cutoff_freq=1) every step"),
    "c4s": dict(nq=32, K=4096, d=512, frames=1 << 17, update=False,
                desc="BASELINE configs[3] shape (32 x 4096 codes, d=512), 128K frames per GPU"),
    "c4": dict(nq=32, K=4096, d=512, frames=1 << 21, update=False,
               desc="BASELINE configs[3] at full size: 32 x 4096 codes, d=512, 2M frames per GPU (16M frames on 8 GPUs)"),
    "c5q": dict(nq=8, K=1024, d=512, frames=1 << 18, update=False,
                desc="reference model default quantizer (8 x 1024, d=512), 256K frames per GPU"),
}
SPEC_BF16_TFLOPS = 2250.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return dict(bf16=j.get("bf16_tflops", 1590.0), bf16_sustained=j.get("bf16_tflops_sustained", 1400.0),
                    hbm=j.get("hbm_gbs", 6650.0), source="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def synth_codebooks(nq, K, d, seed=4321):
    """Per-stage scale decay 0.7^q (SURVEY 8d fallback init) so later stages see realistic residual scales."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randn(K, d, generator=g) * (0.7 ** q) for q in range(nq)])


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def cpu_reference_rate(wl, frames, threads, repeats=1):
    """frames/s of the oracle restatement (the reference's own CPU implementation is the absent third-party
    som_quantizer package; see oracle/rvq_oracle.py) on `frames` frames of workload `wl`."""
    import torch
    from oracle import rvq_oracle as O
    torch.set_num_threads(threads)
    cbs = list(synth_codebooks(wl["nq"], wl["K"], wl["d"]))
    g = torch.Generator().manual_seed(99)
    x = torch.randn(frames, wl["d"], generator=g)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        idx, xq, r, commit = O.rvq_encode_ref(x, cbs, wl["nq"])
        if wl["update"]:
            res = O.stage_residuals_from_indices(x, cbs, idx)
            for q in range(wl["nq"]):
                cnt, sm = O.ema_stats_ref(res[q], idx[:, q], wl["K"])
                O.ema_finalize_ref(cbs[q], torch.ones(wl["K"]), cbs[q].clone(), cnt, sm)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return frames / best, best


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    sample = min(wl["frames"], 1 << 16)
    for _ in range(args.warmup):
        cpu_reference_rate(wl, min(sample, 1 << 13), threads)
    t0 = time.perf_counter()
    rates = []
    for _ in range(args.steps):
        r, _ = cpu_reference_rate(wl, sample, threads)
        rates.append(r)
    dt = time.perf_counter() - t0
    val = sample * args.steps / sum(sample / r for r in rates)
    out = dict(impl="reference", metric="rvq_frames_per_sec", value=val, unit="frames/s", n_gpus=args.gpus,
               steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * sample / val, higher_is_better=True,
               scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
               config=dict(workload=args.workload, desc=wl["desc"], nq=wl["nq"], K=wl["K"], d=wl["d"],
                           frames_per_step=sample),
               cpu_baseline=dict(value=val, unit="frames/s", cores=threads, kind="port",
                                 sample=f"{sample} frames per step of the same workload; oracle/rvq_oracle.py "
                                        f"(restated reference: upstream som_quantizer is not installable)"),
               e2e=dict(value=val, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               wall_s=dt)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="override frames per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--algo", default="tensor")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.frames:
        wl["frames"] = args.frames
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, wl, rank, world)

    import torch
    import torch.distributed as dist
    from audio_generation_b200 import ResidualQuantizer
    from audio_generation_b200.quantizer import HostEncoder

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the RVQ path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"         # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    nq, K, d, N = wl["nq"], wl["K"], wl["d"], wl["frames"]

    # c2/c3 time the north-star path (encode [+ EMA count/sum, all-reduce, refresh]); the SOM neighbourhood and
    # stale-code re-seeding of SURVEY 8f are switched on by the c3m workload only
    quant = ResidualQuantizer(nq, d, "ema", K, algo=args.algo, use_som=bool(wl.get("som", False)),
                              vq_cutoff_freq=float(wl.get("cutoff", 0.0)))
    with torch.no_grad():
        quant.codebooks.copy_(synth_codebooks(nq, K, d))
        quant.ema_sum.copy_(quant.codebooks)
    quant = quant.to(dev)
    quant.train(wl["update"])
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(N, d, device=dev, generator=g)

    def step():
        with torch.no_grad():
            return quant(x, None, update_codebook=wl["update"])

    # EMA workloads start from synthetic codebooks that the first updates pull towards the data (near-degenerate
    # codebooks, many exact re-ranks): time the steady state, not that transient
    n_warm = max(args.warmup, 40 if wl["update"] else 3)
    # the clock sampler (nvidia-smi -lms) starts BEFORE the warm-up: its start-up (NVML initialisation) was seen to
    # stall kernel launches for tens of milliseconds when it fell into the timed region; rows are filtered by time
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record()
    out = None
    for _ in range(n_warm):
        out = step()          # held like in the timed loop: the allocator reaches its steady state here
    w1.record()
    torch.cuda.synchronize()
    # ... and keep warming until the GPU has seen ~0.4 s of this kernel (a box fresh out of idle needs more than a
    # handful of millisecond-long launches to reach its steady clocks).  The number of extra steps is agreed
    # across ranks (EMA workloads all-reduce every step); the timed region below is exactly K steps.
    extra = torch.tensor([max(0.0, 400.0 - w0.elapsed_time(w1)) / max(w0.elapsed_time(w1) / n_warm, 1e-3)], device=dev)
    if world > 1:
        dist.all_reduce(extra, op=dist.ReduceOp.MAX)
    for _ in range(min(int(extra.item()), 2000)):
        out = step()
        n_warm += 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    quant.kernel_events = []          # (start, stop) CUDA events around every rvq_encode launch, on the launch stream
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in quant.kernel_events)
    quant.kernel_events = None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    frames_total = N * world * args.steps
    value = frames_total / (ms * 1e-3)
    ms_per_step = ms / args.steps

    # ---- e2e: host pinned frames -> codes on the host, through the public HostEncoder API
    e2e = None
    if not args.no_e2e and wl["update"]:
        # training-side call: pinned host latents -> device, forward with codebook update, commit loss back to the host
        n_e2e = min(N, 1 << 19)
        xh = torch.randn(n_e2e, d).pin_memory()
        xd = torch.empty(n_e2e, d, device=dev)
        loss_h = torch.empty((), dtype=torch.float32).pin_memory()

        def e2e_step():
            xd.copy_(xh, non_blocking=True)
            with torch.no_grad():
                _, _, commit = quant(xd, None, update_codebook=True)
            loss_h.copy_(commit, non_blocking=True)

        for _ in range(3):
            e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(args.steps, 10))
        e0.record()
        for _ in range(reps):
            e2e_step()
        e1.record()
        torch.cuda.synchronize()
        ems = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ems], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = dict(value=n_e2e * world * reps / (ems * 1e-3), unit="frames/s", h2d_bytes_per_step=n_e2e * d * 4,
                   d2h_bytes_per_step=4, frames_per_step=n_e2e,
                   api="som_quantizer.ResidualQuantizer.forward(update_codebook=True) on host-fed latents")
    if not args.no_e2e and not wl["update"]:
        n_e2e = min(N, 1 << 19)
        xh = torch.randn(n_e2e, d).pin_memory()
        ih = torch.empty((n_e2e, nq), dtype=torch.int64).pin_memory()
        he = HostEncoder(quant.eval())
        for _ in range(3):
            he.encode(xh, ih)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(args.steps, 10))
        e0.record()
        for _ in range(reps):
            he.encode(xh, ih)
        e1.record()
        torch.cuda.synchronize()
        ems = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ems], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = dict(value=n_e2e * world * reps / (ems * 1e-3), unit="frames/s",
                   h2d_bytes_per_step=n_e2e * d * 4, d2h_bytes_per_step=n_e2e * nq * 8,
                   frames_per_step=n_e2e, api="audio_generation_b200.quantizer.HostEncoder.encode")
        quant.train(wl["update"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops_per_frame = nq * 2 * K * d
    per_gpu_rate = value / world
    achieved = N * flops_per_frame / (kernel_ms * 1e-3) / 1e12      # the fused encode kernel alone (rank 0)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload)
        except Exception:
            traffic = None
    roofline = dict(bound="tensor", achieved=achieved, peak=peaks["bf16"], unit="TFLOP/s", frac=achieved / peaks["bf16"],
                    traffic=traffic, peak_source=peaks["source"] + " (burst bf16 cuBLAS 8192^3)",
                    frac_of_sustained=achieved / peaks["bf16_sustained"], frac_of_spec=achieved / SPEC_BF16_TFLOPS,
                    kernel="rvq_encode_tr_kernel" if d <= 128 else "rvq_encode_tc_kernel", flops_per_frame=flops_per_frame,
                    kernel_ms=kernel_ms, step_frac_of_peak=per_gpu_rate * flops_per_frame / 1e12 / peaks["bf16"],
                    note="algorithmic flops = nq*2*K*d per frame (distance GEMM only); achieved = frames per launch x "
                         "flops per frame / mean launch duration of the fused encode kernel (CUDA events on the launch "
                         "stream inside the timed region); step_frac_of_peak uses the whole step instead")
    cpu = None
    if not args.no_cpu and world == 1:
        threads = os.cpu_count() or 1
        sample = min(N, 1 << 17)
        cpu_reference_rate(wl, 1 << 13, threads)
        rate, dt = cpu_reference_rate(wl, sample, threads)
        cpu = dict(value=rate, unit="frames/s", cores=threads, kind="port",
                   sample=f"{sample} frames of the same workload in {dt:.2f}s; oracle/rvq_oracle.py restatement "
                          f"(upstream som_quantizer not installable)")
    # kernels of librvq_sm100a.so per step: the fused encode; with the update also k3_counts + k3_codes (EMA refresh)
    # and k0_stage_max + k0_convert (operands of the refreshed codebooks); c3m adds som_spread, reseed_gather/apply
    launches_per_step = 1 + (4 if wl["update"] else 0) + (1 if wl.get("som") else 0) + (2 if wl.get("cutoff") else 0)
    out = dict(metric="rvq_frames_per_sec", value=value, unit="frames/s", n_gpus=world, steps=args.steps,
               warmup=n_warm, ms_per_step=ms_per_step, higher_is_better=True, scaling="weak",
               vs_baseline=None, dtype="f16-filter/f32-exact", data="synthetic",
               config=dict(workload=args.workload, desc=wl["desc"], nq=nq, K=K, d=d, frames_per_gpu=N,
                           update_codebook=wl["update"], parallelism=f"frames sharded x{world}, codebooks replicated",
                           l2="inputs (%.0f MB) larger than L2; no explicit flush" % (N * d * 4 / 1e6), algo=args.algo),
               roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=args.steps * launches_per_step,
               clocks=clocks)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
