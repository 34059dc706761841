#!/bin/bash
# Round-2 GPU call 6 (one B200): the layer-wise tensor-core GEMM path (unit tests, agent tests, c4 line), recalibrated faithful tests.
set -u
mkdir -p gpurun_out
echo "== gemm_tc unit tests =="
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -q --timeout 120 --timeout-method=thread 2>&1 | tail -25 | tee gpurun_out/r02_c6_gemm_tests.log
echo "== layer-wise agent tests + faithful tests =="
timeout 400 python -m pytest tests/test_agent_gpu.py tests/test_tc_faithful_gpu.py -q --timeout 200 --timeout-method=thread -k "layerwise or faithful" -s 2>&1 | grep -E "faithful-reference|passed|failed|Error|assert" | cut -c1-330 | tee gpurun_out/r02_c6_agent_faithful.log | tail -30
echo "== c4 on the tensor cores =="
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --skip-cpu --skip-e2e 2>gpurun_out/r02_c6_bench_c4.err | tee gpurun_out/r02_c6_bench_c4.json | cut -c1-400
echo "== whole suite =="
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -12 | tee gpurun_out/r02_c6_gpu_tests.log
