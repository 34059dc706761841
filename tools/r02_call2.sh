#!/bin/bash
# Round-2 GPU call 2 (one B200): full -m gpu suite (gates removed), smoke, GAE sweep with the reference Triton comparator, headline bench with
# the c5 block + reference CPU arm, then profiling passes.  Usage: gpurun --timeout 1500 -- 'bash tools/r02_call2.sh'
set -u
mkdir -p gpurun_out
echo "== full gpu suite =="
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02_c2_gpu_tests.log
echo "== faithful-reference error table =="
timeout 200 python -m pytest tests/test_tc_faithful_gpu.py -q -s 2>&1 | grep -E "faithful-reference|passed|failed" | tee gpurun_out/r02_c2_faithful.log
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02_c2_smoke.log
echo "== gae sweep =="
timeout 300 python tools/gae_sweep.py 2>gpurun_out/r02_c2_gae_sweep.err | tee gpurun_out/r02_c2_gae_sweep.log | cut -c1-260 | tail -70
echo "== headline bench (c2 + c5 block + e2e + reference cpu arm) =="
timeout 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_c2_bench.err | tee gpurun_out/r02_c2_bench.json | cut -c1-400
echo "== reference arm =="
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 2>gpurun_out/r02_c2_ref.err | tee gpurun_out/r02_c2_bench_reference.json | cut -c1-600
echo "== ncu: GAE kernels at the c2 / c5 shapes (full set) =="
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:gae_tma_kernel|gae_fused_kernel' -c 6 -f -o gpurun_out/r02_prof_gae \
    python tools/gae_sweep.py --quick > gpurun_out/r02_c2_ncu_gae.log 2>&1
echo "== ncu: c5 launch list + l1_fwd_tc / reduce_adam full =="
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c5.csv \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-graph --skip-cpu --skip-e2e > gpurun_out/r02_c2_ncu_launches_c5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:l1_fwd_tc_kernel|l1_wgrad_tc_kernel|reduce_adam_kernel' --launch-skip 400 --launch-count 4 \
    -f -o gpurun_out/r02_prof_c5 python bench.py --workload c5 --steps 1 --warmup 3 --no-graph --skip-cpu --skip-e2e > gpurun_out/r02_c2_ncu_full_c5.log 2>&1
echo "== stage timelines (instrumented build) =="
if [ -f rl_games_b200/libb200rl_timing.so ]; then
  B200RL_LIB_PATH=$PWD/rl_games_b200/libb200rl_timing.so timeout 120 python tools/tc_stage_timing.py --old-bwd2 2>&1 | tail -80 | tee gpurun_out/r02_c2_stage_timing.log
fi
ls -la gpurun_out | tail -12
