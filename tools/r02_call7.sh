#!/bin/bash
# Round-2 GPU call 7 (one B200): layer-wise GEMM path after the bf16 weight arena + 16 row splits; c4 line; suite.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py tests/test_agent_gpu.py -q --timeout 200 --timeout-method=thread -k "gemm or layerwise or linear_" 2>&1 | tail -8 | tee gpurun_out/r02_c7_gemm_tests.log
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --skip-cpu --skip-e2e 2>gpurun_out/r02_c7_bench_c4.err | tee gpurun_out/r02_c7_bench_c4.json | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -5 | tee gpurun_out/r02_c7_gpu_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu 2>/dev/null | tee gpurun_out/r02_c7_bench.json | cut -c1-200
