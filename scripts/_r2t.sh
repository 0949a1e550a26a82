timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2t_tests.log
(python scripts/phase_profile.py c5q; python scripts/phase_profile.py c3) > gpurun_out/r2t_phase.log 2>&1
python bench.py --workload c5q --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2t_bench_c5q.json 2>/dev/null
python bench.py --workload c4s --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2t_bench_c4s.json 2>/dev/null
python scripts/small_call_profile.py > gpurun_out/r2t_small_profile.log 2>&1
cat gpurun_out/r2t_tests.log gpurun_out/r2t_phase.log gpurun_out/r2t_small_profile.log
