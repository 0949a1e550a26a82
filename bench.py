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
    "c5": dict(nq=8, K=1024, d=512, frames=8 * 500, update=False,
               desc="BASELINE configs[4]: causal conv encoder -> RVQ -> decoder on 8 x 10 s of 24 kHz audio per GPU "
                    "(64 x 10 s on 8 GPUs); conv stacks = scripts/c5_harness.py stand-in with the reference's channel / "
                    "stride plan (out of scope as kernels: torch / cuDNN)"),
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
    """ONE `nvidia-smi -lms` process (started by rank 0) that samples EVERY GPU of the box: per-rank clocks without
    one NVML client per rank (eight pollers were seen to perturb an 8-GPU step)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.rows, self.proc = [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout=10.0):
        """Block until the first sample has arrived (nvidia-smi needs 1-2 s to start on an 8-GPU box; a timed region
        that ends before that would have no clocks at all)."""
        t_end = time.time() + timeout
        while self.proc is not None and not self.rows and time.time() < t_end:
            time.sleep(0.05)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            self.proc = None

    def window(self, t0, t1, gpus):
        """Per-GPU clock summary (list indexed like `gpus`) of the samples taken between wall-clock times t0 and t1
        (the sampler keeps running)."""
        if self.proc is None:
            return [None] * len(gpus)
        time.sleep(0.15)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15]
        if not rows and self.rows:
            # no sample inside the region (shorter than the sampling period): the ones closest to it in time
            mid = 0.5 * (t0 + t1)
            tmin = min(abs(t - mid) for (t, _) in self.rows)
            rows = [r for (t, r) in self.rows if abs(t - mid) <= tmin + 0.03]   # (the rows of one poll arrive together)
        out = []
        for gidx in gpus:
            sm, mx, pw, reasons = [], [], [], set()
            for r in rows:
                f = [x.strip() for x in r.split(",")]
                try:
                    if int(f[0]) != gidx:
                        continue
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                    pw.append(float(f[3]))
                except Exception:
                    continue
                for nme, v in zip(names, f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            out.append(dict(gpu=gidx, sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w=statistics.median(pw),
                            reasons=sorted(reasons), samples=len(sm)) if sm else None)
        return out


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


def bind_to_gpu_numa(local):
    """Best effort: run this rank (and first-touch its pinned buffers) on the NUMA node its GPU hangs off, so that
    concurrent host<->device copies of several ranks do not all cross one memory controller / PCIe root."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        devs = [d for d in os.listdir("/sys/bus/pci/devices") if d.lower().startswith("%04x:%02x:" % (dom, bus))]
        node = int(open(f"/sys/bus/pci/devices/{devs[0]}/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def build_quantizer(wl, dev, args):
    import torch
    from audio_generation_b200 import ResidualQuantizer
    nq, K, d = wl["nq"], wl["K"], wl["d"]
    # c2/c3 time the north-star path (encode [+ EMA count/sum, all-reduce, refresh]); the SOM neighbourhood and
    # stale-code re-seeding of SURVEY 8f are switched on by the c3m workload only
    quant = ResidualQuantizer(nq, d, "ema", K, algo=args.algo, use_som=bool(wl.get("som", False)),
                              vq_cutoff_freq=float(wl.get("cutoff", 0.0)), kernel=args.kernel)
    with torch.no_grad():
        quant.codebooks.copy_(synth_codebooks(nq, K, d))
        quant.ema_sum.copy_(quant.codebooks)
    quant = quant.to(dev)
    quant.train(wl["update"])
    return quant


def timed_steps(quant, x, wl, steps, warmup, rank, world, dev, settle=True):
    """W (+ clock-settling) untimed steps, then exactly `steps` timed steps bracketed by barrier + synchronize;
    device time = max over ranks.  Returns a dict of the raw numbers."""
    import torch
    import torch.distributed as dist

    def step():
        with torch.no_grad():
            return quant(x, None, update_codebook=wl["update"])

    # EMA workloads start from synthetic codebooks that the updates pull towards the data with a time constant of
    # 1 / (1 - decay) = 100 steps, and the kernel's time depends on the codebooks (how many frames need the exact
    # re-rank): time the steady state, not that transient (profiles/r2b_c3_comm_probe_2gpu.log: the first 30 steps after
    # the statistics change scale run 5 % slower than the steady state)
    n_warm = max(warmup, 150 if wl["update"] else 3) if settle else warmup
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
    if not settle:
        extra.zero_()
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
    quant.comm_events = []            # ... around every all-reduce of the statistics
    quant.update_events = []          # ... and around the codebook maintenance behind the kernel (all-reduce included)
    t_wall0 = time.time()
    ev0.record()
    for _ in range(steps):
        out = step()
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    del out
    ms_local = ev0.elapsed_time(ev1)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in quant.kernel_events)
    comm_ms = statistics.mean(a.elapsed_time(b) for a, b in quant.comm_events) if quant.comm_events else 0.0
    upd_ms = statistics.mean(a.elapsed_time(b) for a, b in quant.update_events) if quant.update_events else 0.0
    # device time between the end of the encode kernel and the start of the all-reduce (reseed gather, host launch gap)
    pre_ms = (statistics.mean(u[0].elapsed_time(c[0]) for u, c in zip(quant.update_events, quant.comm_events))
              if quant.comm_events and quant.update_events else 0.0)
    quant.kernel_events = quant.comm_events = quant.update_events = None
    ms = ms_local
    per_rank = None
    if world > 1:
        t = torch.tensor([ms_local, kernel_ms, comm_ms, upd_ms, pre_ms], device=dev, dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        dist.barrier()
        per_rank = [[float(v) for v in a.tolist()] for a in allt]
        ms = max(p[0] for p in per_rank)
    return dict(ms=ms, ms_local=ms_local, kernel_ms=kernel_ms, comm_ms=comm_ms, upd_ms=upd_ms, n_warm=n_warm,
                per_rank=per_rank, wall=(t_wall0, t_wall1))


def replicas_identical(quant, world, dev):
    """True if every rank holds bit-identical codebooks and EMA state (what scripts/dist_check.py asserts)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return True
    sig = torch.stack([quant.codebooks.detach().view(torch.int32).to(torch.int64).sum(),
                       quant.ema_count.view(torch.int32).to(torch.int64).sum(),
                       quant.ema_sum.view(torch.int32).to(torch.int64).sum()])
    alls = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(alls, sig)
    return all(bool(torch.equal(a, alls[0])) for a in alls)


def collective_leg(args, rank, world, local, dev, sampler):
    """BASELINE configs[2] under the driver's own launch: c3 = encode + EMA statistics + all-reduce + refresh, the only
    step of the path that crosses frame shards (training.py:305-308,325-328 -> vae.py:315-318)."""
    import torch
    import torch.distributed as dist
    wl = dict(WORKLOADS["c3"])
    if args.frames:
        wl["frames"] = args.frames
    nq, K, d, N = wl["nq"], wl["K"], wl["d"], wl["frames"]
    quant = build_quantizer(wl, dev, args)
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    x = torch.randn(N, d, device=dev, generator=g)
    steps = max(5, min(args.steps, 20))
    r = timed_steps(quant, x, wl, steps, args.warmup, rank, world, dev)
    parity = replicas_identical(quant, world, dev)
    walls = [r["wall"]]
    if world > 1:
        walls = [None] * world
        dist.all_gather_object(walls, r["wall"])
    all_clocks = sampler.window(min(w[0] for w in walls), max(w[1] for w in walls), list(range(world))) if rank == 0 else None
    # the same step with the all-reduce skipped (every rank updates from its own shard): what this GPU does on
    # its own while its neighbours are just as busy - the denominator of the collective's efficiency.  Timed RIGHT
    # AFTER the synchronised steps with three warm-up steps only: the encode kernel's time depends on the codebooks
    # (profiles/r2b_c3_comm_probe_8gpu.log), and the two measurements must see the same ones.
    local_rate = local_kms = None
    if world > 1:
        quant.sync_stats = False
        r0 = timed_steps(quant, x, wl, steps, 3, rank, world, dev, settle=False)
        quant.sync_stats = True
        local_rate = N * world * steps / (r0["ms"] * 1e-3)
        local_kms = [p[1] for p in r0["per_rank"]]
    value = N * world * steps / (r["ms"] * 1e-3)
    flops_per_frame = nq * 2 * K * d
    peaks = load_peaks()
    kms = [p[1] for p in r["per_rank"]] if r["per_rank"] else [r["kernel_ms"]]
    cms = [p[2] for p in r["per_rank"]] if r["per_rank"] else [r["comm_ms"]]
    ums = [p[3] for p in r["per_rank"]] if r["per_rank"] else [r["upd_ms"]]
    pms = [p[4] for p in r["per_rank"]] if r["per_rank"] else [0.0]
    ms_step = r["ms"] / steps
    out = dict(workload="c3", desc=wl["desc"], value=value, unit="frames/s", per_gpu_value=value / world,
               steps=steps, ms_per_step=ms_step, frames_per_gpu=N,
               kernel_ms_min=min(kms), kernel_ms_max=max(kms), kernel_ms_per_rank=kms,
               allreduce_ms=max(cms), allreduce_ms_per_rank=cms,
               update_ms_per_rank=ums, kernel_end_to_allreduce_start_ms_per_rank=pms, k0_k3_ms=max(u - c for u, c in zip(ums, cms)),
               k0_k3_note="CUDA events from the end of the encode kernel to the end of the codebook maintenance "
                          "(update_ms: all-reduce + rvq_ema_finalize K3 [+ SOM / re-seeding]); k0_k3_ms = that minus the "
                          "all-reduce; rvq_prepare_codebooks (K0) runs at the start of the next call, inside ms_per_step",
               kernel_ms_per_rank_without_allreduce=local_kms,
               payload_bytes=nq * K * (d + 1) * 4,
               value_without_allreduce=local_rate,
               efficiency_vs_local_update=(value / local_rate) if local_rate else None,
               roofline_frac=N * flops_per_frame / (max(kms) * 1e-3) / 1e12 / peaks["bf16"],
               clocks_per_rank=all_clocks, parity="ok" if parity else "REPLICAS DIFFER")
    del quant, x
    torch.cuda.empty_cache()
    return out


def latency_leg(args, dev):
    """Microseconds per `forward` through the module at the call shapes the reference really makes, the CPU oracle
    beside each (same codebooks, same input)."""
    import torch
    from audio_generation_b200 import ResidualQuantizer
    from oracle import rvq_oracle as O
    shapes = [
        dict(name="training call (training.py:296-328): B=4, L=150, d=512, nq=10, K=512, (B,d,L)-backed view, "
                  "update_codebook=True", B=4, L=150, d=512, nq=10, K=512, update=True),
        dict(name="C1 fixture shape (om.wav through config/training.yml): 1 x 136 x 512, nq=10, K=512, encode",
             B=1, L=136, d=512, nq=10, K=512, update=False),
        dict(name="C5 share of one GPU (8 x 10 s @ 24 kHz): 8 x 500 x 512, model default nq=8, K=1024, encode",
             B=8, L=500, d=512, nq=8, K=1024, update=False),
    ]
    out = []
    for sh in shapes:
        torch.manual_seed(0)
        q = ResidualQuantizer(sh["nq"], sh["d"], "ema", sh["K"], use_som=False, vq_cutoff_freq=0.0, kernel=args.kernel)
        ref = O.ResidualQuantizerRef(sh["nq"], sh["d"], "ema", sh["K"], use_som=False, vq_cutoff_freq=0.0)
        ref.load_state_dict({k: v for k, v in q.state_dict().items() if k in ref.state_dict()}, strict=False)
        q = q.to(dev).train(sh["update"])
        ref.train(sh["update"])
        xc = torch.randn(sh["B"], sh["d"], sh["L"])
        xv = xc.to(dev).permute(0, 2, 1)              # the reference's "b c l -> b l c" view (vae.py:313)
        with torch.no_grad():
            for _ in range(20):
                q(xv, None, update_codebook=sh["update"])
            torch.cuda.synchronize()
            reps = 200
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(reps):
                q(xv, None, update_codebook=sh["update"])
            e1.record()
            torch.cuda.synchronize()
            wall_us = (time.perf_counter() - t0) / reps * 1e6
            dev_us = e0.elapsed_time(e1) / reps * 1e3
            graph_us = None
            if not sh["update"]:
                # the same call captured once in a CUDA graph and replayed (launch + Python overhead removed)
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    q(xv, None)
                for _ in range(5):
                    gr.replay()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(reps):
                    gr.replay()
                e1.record()
                torch.cuda.synchronize()
                graph_us = e0.elapsed_time(e1) / reps * 1e3
            xr = xc.permute(0, 2, 1)
            ref(xr, None, update_codebook=sh["update"])
            t0 = time.perf_counter()
            for _ in range(3):
                ref(xr, None, update_codebook=sh["update"])
            cpu_us = (time.perf_counter() - t0) / 3 * 1e6
        out.append(dict(shape=sh["name"], frames=sh["B"] * sh["L"], us_per_call_device=dev_us,
                        us_per_call_host_wall=wall_us, us_per_call_cuda_graph_replay=graph_us,
                        cpu_oracle_us_per_call=cpu_us, cpu_threads=torch.get_num_threads()))
        del q
    return out


def run_c5(args, wl, rank, world, local, dev):
    """BASELINE configs[4]: the whole model step around the drop-in (encoder -> "b c l -> b l c" view -> quantizer ->
    decoder), 8 waveforms of 10 s per GPU, inference forward; reports how much of the step the RVQ path is."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from c5_harness import SyntheticCausalVQAE
    torch.manual_seed(0)
    model = SyntheticCausalVQAE(num_quantizers=wl["nq"], codebook_size=wl["K"]).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    B, T = 8, 240000
    x = torch.randn(B, 1, T, device=dev, generator=g) * 0.1
    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
        sampler.wait_first()

    def step():
        with torch.no_grad():
            return model(x)

    for _ in range(max(3, args.warmup)):
        y, commit, index = step()
    torch.cuda.synchronize()
    frames = index.shape[0] * index.shape[1]
    if world > 1:
        dist.barrier()
    model.quantizer.kernel_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    kms = statistics.mean(a.elapsed_time(b) for a, b in model.quantizer.kernel_events)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.window(t0, t1, list(range(world))) if rank == 0 else None
    sampler.stop()
    if rank == 0:
        out = dict(metric="rvq_frames_per_sec", value=frames * world * args.steps / (ms * 1e-3), unit="frames/s",
                   n_gpus=world, steps=args.steps, warmup=max(3, args.warmup), ms_per_step=ms / args.steps,
                   higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32 convs (cuDNN) + f16-filter/f32-exact RVQ",
                   data="synthetic", config=dict(workload="c5", desc=wl["desc"], nq=wl["nq"], K=wl["K"], d=wl["d"],
                                                 waveforms_per_gpu=B, samples=T, frames_per_gpu=frames),
                   rvq=dict(kernel_ms=kms, share_of_step=kms / (ms / args.steps), frames_per_call=frames,
                            index_shape=list(index.shape)),
                   gpu_launches=args.steps, clocks=clocks)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--no-collective", action="store_true", help="skip the c3 (encode + EMA + all-reduce) leg")
    ap.add_argument("--no-latency", action="store_true", help="skip the small-call latency leg")
    ap.add_argument("--algo", default="tensor")
    ap.add_argument("--kernel", default="auto", help="auto | tmem | generic | frame (cross-check kernels)")
    ap.add_argument("--layout", default="rows", choices=["rows", "bdl"],
                    help="rows: frames [N, d] (default); bdl: the reference's '(B, d, L) -> b l c' view with L = 512 "
                         "(vae.py:313), device-resident legs only")
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
    from audio_generation_b200.quantizer import HostEncoder

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the RVQ path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created (at NCCL_DEBUG=VERSION and WARN
        # alike): stdout is pointed at stderr while that happens, so that rank 0 prints ONE JSON line and nothing else
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            t0 = torch.zeros(1, device=dev)
            dist.all_reduce(t0)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    if args.workload == "c5":
        return run_c5(args, wl, rank, world, local, dev)
    peaks = load_peaks()
    nq, K, d, N = wl["nq"], wl["K"], wl["d"], wl["frames"]

    quant = build_quantizer(wl, dev, args)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    if args.layout == "bdl":
        # the reference's boundary layout: a (B, L, d) VIEW of (B, d, L) storage, features strided by L
        x = torch.randn(max(N // 512, 1), d, 512, device=dev, generator=g).permute(0, 2, 1)
        N = x.shape[0] * x.shape[1]
        args.no_e2e = True
    else:
        x = torch.randn(N, d, device=dev, generator=g)
    # the clock sampler (nvidia-smi -lms) starts BEFORE the warm-up: its start-up (NVML initialisation) was seen to
    # stall kernel launches for tens of milliseconds when it fell into the timed region; rows are filtered by time.
    # One sampler per rank, each on its own GPU.
    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    r = timed_steps(quant, x, wl, args.steps, args.warmup, rank, world, dev)
    ms, kernel_ms, n_warm = r["ms"], r["kernel_ms"], r["n_warm"]
    clocks = None
    if rank == 0:
        per_gpu = sampler.window(*r["wall"], list(range(world)))
        clocks = dict(per_gpu[0] or {}, per_gpu=per_gpu) if per_gpu and per_gpu[0] else None
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
        n_e2e = min(N, 1 << 20)                          # one call = the workload's batch
        xh = torch.randn(n_e2e, d).pin_memory()
        he = HostEncoder(quant.eval(), packed=True)     # codes cross PCIe in their wire format (10 bits per code)
        bpf = he.bytes_per_frame(nq)
        ih = torch.empty((n_e2e, bpf), dtype=torch.uint8).pin_memory()
        for _ in range(3):
            he.encode(xh, ih)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = max(3, min(args.steps, 10))
        # HostEncoder.encode returns when the codes are on the host: host wall clock around the calls IS the end-to-end
        # time; the device events bracket the same region
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            he.encode(xh, ih)
        e1.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ems = max(e0.elapsed_time(e1), wall_ms)
        h2d_gbs = n_e2e * d * 4 * reps / (ems * 1e-3) / 1e9
        if world > 1:
            t = torch.tensor([ems], device=dev, dtype=torch.float64)
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            per_rank_gbs = [n_e2e * d * 4 * reps / (float(a.item()) * 1e-3) / 1e9 for a in allt]
            ems = max(float(a.item()) for a in allt)
        else:
            per_rank_gbs = [h2d_gbs]
        e2e = dict(value=n_e2e * world * reps / (ems * 1e-3), unit="frames/s",
                   h2d_bytes_per_step=n_e2e * d * 4, d2h_bytes_per_step=n_e2e * bpf,
                   frames_per_step=n_e2e, h2d_gbs_per_rank=per_rank_gbs, numa_node=numa,
                   api="audio_generation_b200.quantizer.HostEncoder(packed=True).encode: pinned host frames in, packed "
                       "codes on the host when the call returns")
        quant.train(wl["update"])

    del x
    torch.cuda.empty_cache()
    collective = None
    if not args.no_collective and args.workload == "c2":
        collective = collective_leg(args, rank, world, local, dev, sampler)
    sampler.stop()
    latency = None
    if rank == 0 and world == 1 and not args.no_latency and args.workload == "c2":
        latency = latency_leg(args, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops_per_frame = nq * 2 * K * d
    per_gpu_rate = value / world
    achieved = N * flops_per_frame / (kernel_ms * 1e-3) / 1e12      # the fused encode kernel alone (rank 0)
    traffic = traffic_src = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get(args.workload)
            traffic_src = tj.get("_source")
        except Exception:
            traffic = None
    kname = {"tmem": "rvq_encode_tr_kernel", "frame": "rvq_encode_fr_kernel", "generic": "rvq_encode_tc_kernel"}.get(
        args.kernel, "rvq_encode_tr_kernel" if d <= 128 else "rvq_encode_tc_kernel")
    roofline = dict(bound="tensor", achieved=achieved, peak=peaks["bf16"], unit="TFLOP/s", frac=achieved / peaks["bf16"],
                    traffic=traffic,
                    traffic_source=(traffic_src or "profiles/traffic.json") + " (dram__bytes_read+write of one earlier "
                                   "`ncu --set full` capture of this kernel at this size; NOT measured by this run)",
                    peak_source=peaks["source"] + " (burst bf16 cuBLAS 8192^3)",
                    frac_of_sustained=achieved / peaks["bf16_sustained"], frac_of_spec=achieved / SPEC_BF16_TFLOPS,
                    kernel=kname, flops_per_frame=flops_per_frame,
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
               warmup=args.warmup, warmup_extra=n_warm - args.warmup,
               warmup_note="the W requested warm-up steps plus `warmup_extra` untimed clock-settling steps (~0.4 s of "
                           "this kernel); the timed region is exactly `steps` steps",
               ms_per_step=ms_per_step, higher_is_better=True, scaling="weak",
               vs_baseline=None, dtype="f16-filter/f32-exact", data="synthetic",
               config=dict(workload=args.workload, desc=wl["desc"], nq=nq, K=K, d=d, frames_per_gpu=N,
                           update_codebook=wl["update"], parallelism=f"frames sharded x{world}, codebooks replicated",
                           l2="inputs (%.0f MB) larger than L2; no explicit flush" % (N * d * 4 / 1e6), algo=args.algo,
                           layout=args.layout,
                           kernel=args.kernel),
               roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=args.steps * launches_per_step,
               clocks=clocks, collective=collective, latency=latency)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
