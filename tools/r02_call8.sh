#!/bin/bash
# Round-2 GPU call 8 (one B200): gemm_tc.cu after the register-prefetch / transposed-epilogue rewrite, 64-bit grid-barrier counters.
# Safe evidence first (suite, headline), then the per-shape GEMM sweep, the c4 line and its ncu captures.
set -u
mkdir -p gpurun_out
echo "== gemm tests =="
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -q --timeout 120 --timeout-method=thread 2>&1 | tail -6 | tee gpurun_out/r02_c8_gemm_tests.log
echo "== full GPU suite =="
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -6 | tee gpurun_out/r02_c8_gpu_tests.log
echo "== GEMM sweep =="
timeout 200 python tools/gemm_tc_sweep.py 2>&1 | tee gpurun_out/r02_c8_gemm_sweep.log | cut -c1-260
echo "== bench c4 =="
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --skip-cpu --skip-e2e 2>gpurun_out/r02_c8_bench_c4.err | tee gpurun_out/r02_c8_bench_c4.json | cut -c1-300
echo "== default bench =="
timeout 400 python bench.py 2>gpurun_out/r02_c8_bench.err | tee gpurun_out/r02_c8_bench.json | cut -c1-300
echo "== ncu: c4 launch list (one eager epoch) =="
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv \
    python bench.py --workload c4 --steps 1 --warmup 1 --no-graph --skip-cpu --skip-e2e > gpurun_out/r02_c8_ncu_launches_c4.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_c4.csv 2>&1 | head -30 | tee gpurun_out/r02_c4_launch_summary.txt
echo "== ncu --set full: the three GEMM kernels at the update shapes =="
timeout 240 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc_kernel' --launch-skip 900 --launch-count 6 \
    -f -o gpurun_out/r02_prof_c4_gemm python bench.py --workload c4 --steps 1 --warmup 1 --no-graph --skip-cpu --skip-e2e > gpurun_out/r02_c8_ncu_full_c4.log 2>&1
ls -la gpurun_out | tail -12
