#!/bin/bash
# Round-2 GPU call 9 (one B200): gemm_tc.cu with prefetched / activation-specialised epilogues and the 8-element-aligned parameter
# arena (16-byte bf16 weight loads): tests, smoke, per-shape sweep, c4 line.  (Headline + ncu evidence: call 10.)
set -u
mkdir -p gpurun_out
echo "== gemm tests =="
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -q --timeout 120 --timeout-method=thread 2>&1 | tail -6 | tee gpurun_out/r02_c9_gemm_tests.log
echo "== full GPU suite =="
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -6 | tee gpurun_out/r02_c9_gpu_tests.log
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | tee gpurun_out/r02_c9_smoke.log
echo "== GEMM sweep =="
timeout 200 python tools/gemm_tc_sweep.py 2>&1 | tee gpurun_out/r02_c9_gemm_sweep.log | cut -c1-260
echo "== bench c4 =="
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --skip-cpu --skip-e2e 2>gpurun_out/r02_c9_bench_c4.err | tee gpurun_out/r02_c9_bench_c4.json | cut -c1-300
