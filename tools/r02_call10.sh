#!/bin/bash
# Round-2 GPU call 10 (one B200, the last ~10 GPU-minutes of the round): HEAD validation in priority order -- the driver's own round-end
# commands first (suite, smoke, default bench), then the c4 line and the per-shape GEMM sweep of the rewritten gemm_tc.cu, then ncu.
# Every step tees into gpurun_out/ as it goes: whatever finished before the budget clamp is kept.
set -u
mkdir -p gpurun_out
date -u +%T | tee gpurun_out/r02_c10_times.log
echo "== full GPU suite (includes tests/test_gemm_tc_gpu.py) =="
timeout 420 python -m pytest tests -m gpu -q --timeout 200 --timeout-method=thread 2>&1 | tail -12 | tee gpurun_out/r02_c10_gpu_tests.log
date -u +%T | tee -a gpurun_out/r02_c10_times.log
echo "== smoke =="
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | tee gpurun_out/r02_c10_smoke.log
date -u +%T | tee -a gpurun_out/r02_c10_times.log
echo "== bench c4 =="
timeout 200 python bench.py --workload c4 --steps 5 --warmup 3 --skip-cpu --skip-e2e 2>gpurun_out/r02_c10_bench_c4.err | tee gpurun_out/r02_c10_bench_c4.json | cut -c1-300
date -u +%T | tee -a gpurun_out/r02_c10_times.log
echo "== default bench (the driver's command) =="
timeout 400 python bench.py 2>gpurun_out/r02_c10_bench.err | tee gpurun_out/r02_c10_bench.json | cut -c1-300
date -u +%T | tee -a gpurun_out/r02_c10_times.log
echo "== GEMM sweep =="
timeout 150 python tools/gemm_tc_sweep.py 2>&1 | tee gpurun_out/r02_c10_gemm_sweep.log | cut -c1-260
date -u +%T | tee -a gpurun_out/r02_c10_times.log
echo "== ncu: c4 launch list (one eager epoch) =="
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --workload c4 --steps 1 --warmup 1 --no-graph --skip-cpu --skip-e2e > gpurun_out/r02_c10_ncu_launches_c4.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_c4.csv 2>&1 | head -30 | tee gpurun_out/r02_c4_launch_summary.txt
date -u +%T | tee -a gpurun_out/r02_c10_times.log
ls -la gpurun_out | tail -14
