"""Per-phase cycle counters of the encode kernel (RVQ_FLAG_COUNTERS).  python scripts/phase_profile.py c2"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from audio_generation_b200 import ResidualQuantizer

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WORKLOADS[name]
nq, K, d, N = wl["nq"], wl["K"], wl["d"], min(wl["frames"], 1 << 18)
q = ResidualQuantizer(nq, d, "ema", K, use_som=False, vq_cutoff_freq=0)
with torch.no_grad():
    q.codebooks.copy_(bench.synth_codebooks(nq, K, d))
upd = "--update" in sys.argv
nsteps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 3
with torch.no_grad():
    q.ema_sum.copy_(q.codebooks)
q = q.cuda().train(upd)
q.counters = True
x = torch.randn(N, d, device="cuda")
for _ in range(nsteps):
    with torch.no_grad():
        q(x if not upd else torch.randn(N, d, device="cuda"), None, update_codebook=upd)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
with torch.no_grad():
    q(x, None, update_codebook=upd)
ev1.record()
torch.cuda.synchronize()
ws = q._ws
prof = ws[(ws.numel() - 256) & ~7:][:256].view(torch.int64).cpu().tolist()
n = max(prof[5], 1)
ms = ev0.elapsed_time(ev1)
ctas = min(148, (N + 127) // 128)
cyc_total = ms * 1e-3 * 1.965e9 * ctas / n
print(f"workload {name} update={upd} after {nsteps} steps N={N} ms={ms:.3f} tile-stages={prof[5]}  wall cycles per tile-stage per CTA ~{cyc_total:.0f} "
      f"(MMA floor {K * d * 128 // 4096})")
if d <= 128:
    print(f"cycles per tile-stage: scan={prof[0]/n:.0f} (+a_ready wait {prof[1]/n:.0f}, of scan: tmem_full wait {prof[11]/n:.0f})")
    print(f"  update group (per job): total={prof[2]/n:.0f} (+scan_done wait {prof[7]/n:.0f})  staging-acquire={prof[3]/n:.0f} "
          f"rerank={prof[8]/n:.0f} gather-wait={prof[10]/n:.0f} apply={prof[9]/n:.0f} tail={prof[12]/n:.0f}")
    print(f"  rerank split: classify+barrier={prof[13]/n:.0f} expose-rows+barrier={prof[14]/n:.0f} score={prof[15]/n:.0f}")
    print(f"  control warps per tile-stage: producer waits for a free ring slot={prof[16]/n:.0f}  MMA waits for codebook data={prof[17]/n:.0f} "
          f"for the operand (a_ready)={prof[18]/n:.0f} for a free accumulator={prof[19]/n:.0f}")
    print(f"frames whose exact winner differs from the approximate argmin, per tile-stage: {prof[20]/n:.3f}")
    print(f"dirty rows per tile-stage={prof[4]/n:.3f}  multi-candidate rows per tile-stage={prof[6]/n:.2f}")
    sys.exit(0)
print(f"cycles per tile-stage: scan={prof[0]/n:.0f} (+wait {prof[1]/n:.0f})  update={prof[2]/n:.0f} "
      f"dirty={prof[3]/n:.0f} (+wait {prof[7]/n:.0f})")
print(f"  scan: of which waiting for the accumulator (tmem_full) = {prof[11]/n:.0f}")
print(f"  update breakdown per tile-stage (thread 0): score={prof[8]/n:.0f} apply={prof[9]/n:.0f} score-passes={prof[10]/n:.2f}")
print(f"dirty rows per tile-stage={prof[4]/n:.3f}  two-candidate rows per tile-stage={prof[6]/n:.2f}")
