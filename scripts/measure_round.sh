#!/bin/bash
# One-GPU measurement pass of a round: bench lines, ncu launch list, ncu full capture of the top kernel.
#   gpurun -- 'bash scripts/measure_round.sh r1b'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -u
tag=${1:-r1c}
out=gpurun_out
mkdir -p $out
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_c2.json 2> $out/${tag}_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-collective --no-latency --kernel frame > $out/${tag}_bench_c2_frame_kernel.json 2>> $out/${tag}_bench.err
python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench_c3.json 2>> $out/${tag}_bench.err
python bench.py --workload c3m --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench_c3m.json 2>> $out/${tag}_bench.err
python bench.py --workload c4s --steps 5 --warmup 3 --no-cpu --no-e2e > $out/${tag}_bench_c4s.json 2>> $out/${tag}_bench.err
python bench.py --workload c5q --steps 10 --warmup 3 --no-cpu --no-e2e > $out/${tag}_bench_c5q.json 2>> $out/${tag}_bench.err
python bench.py --workload c5 --steps 5 --warmup 3 > $out/${tag}_bench_c5.json 2>> $out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_cpu.json 2>> $out/${tag}_bench.err
# launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/${tag}_launches_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $out/${tag}_ncu_launch.log 2>&1
# full capture of the fused kernel at the bench size (1M frames)
ncu --set full --import-source on --clock-control none -k regex:rvq_encode_tr --launch-skip 2 --launch-count 1 \
    -o $out/${tag}_tr_c2_full -f python scripts/ncu_target.py c2 1048576 > $out/${tag}_ncu_full.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:rvq_encode_tc --launch-skip 2 --launch-count 1 \
    -o $out/${tag}_tc_c3_full -f python scripts/ncu_target.py c3 1048576 update > $out/${tag}_ncu_full_c3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/${tag}_launches_c3.csv \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --no-e2e > $out/${tag}_ncu_launch_c3.log 2>&1
for w in c2 c3; do python scripts/phase_profile.py $w >> $out/${tag}_phase_profile.log 2>&1; done
# row-major frames vs the reference's (B, d, L) storage; the reference's small call shapes; the codebook state 1 / 8
# shards converge to; bare pinned-copy rate of the box
python scripts/layout_time.py > $out/${tag}_layout_time.log 2>&1
python scripts/small_call_profile.py > $out/${tag}_small_profile.log 2>&1
python scripts/c3_shards_probe.py 1 8 > $out/${tag}_shards_probe.log 2>&1
python scripts/h2d_ceiling.py > $out/${tag}_h2d_ceiling.log 2>&1
tail -2 $out/${tag}_bench.err
for f in c2 c2_frame_kernel c3 c3m c4s c5q reference_cpu; do python - <<PY
import json
j = json.load(open("$out/${tag}_bench_$f.json"))
r = j.get("roofline") or {}
print("$f", round(j["value"] / 1e6, 3), "M frames/s", round(j["ms_per_step"], 3), "ms", "frac", r.get("frac"), "e2e", (j.get("e2e") or {}).get("value"), j.get("clocks"))
PY
done
