#!/bin/bash
# Round-2 GPU call 11 (one B200, the last ~4 GPU-minutes): re-validation of the final commit (product library without the test hooks --
# device SASS byte-identical to call 10's build -- plus the new GPU tests: adapter / observer route, min_sigma), then one `ncu --set full` capture of the layer-wise
# tcgen05 GEMMs at the c4 shapes.  Everything tees into gpurun_out/ as it goes.
set -u
mkdir -p gpurun_out
date -u +%T | tee gpurun_out/r02_c11_times.log
echo "== new tests first (verbose): manager-based adapter + observer, min_sigma golden run, min_sigma on the layer-wise path =="
timeout 150 python -m pytest tests/test_zz_env_adapters_gpu.py tests/test_agent_gpu.py -m gpu -v --timeout 100 --timeout-method=thread \
    -k "adapter or critic_group or minsigma or min_sigma" 2>&1 | tail -25 | tee gpurun_out/r02_c11_new_tests.log
date -u +%T | tee -a gpurun_out/r02_c11_times.log
echo "== full GPU suite (the driver's command) =="
timeout 300 python -m pytest tests/ -x -q -m gpu --timeout 200 --timeout-method=thread 2>&1 | tail -8 | tee gpurun_out/r02_c11_gpu_tests.log
date -u +%T | tee -a gpurun_out/r02_c11_times.log
echo "== smoke =="
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_c11_smoke.log
date -u +%T | tee -a gpurun_out/r02_c11_times.log
echo "== ncu --set full: gemm_tc_kernel<fwd|dgrad|wgrad> at three c4 shapes =="
timeout 150 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc_kernel' --launch-count 9 \
    -f -o gpurun_out/r02_prof_c4_gemm python tools/gemm_tc_ncu_target.py > gpurun_out/r02_c11_ncu_full_c4.log 2>&1
tail -4 gpurun_out/r02_c11_ncu_full_c4.log
date -u +%T | tee -a gpurun_out/r02_c11_times.log
echo "== default bench (the driver's command) =="
timeout 200 python bench.py 2>gpurun_out/r02_c11_bench.err | tee gpurun_out/r02_c11_bench.json | cut -c1-260
date -u +%T | tee -a gpurun_out/r02_c11_times.log
ls -la gpurun_out | tail -12
