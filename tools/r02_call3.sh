#!/bin/bash
# Round-2 GPU call 3 (one B200): A/B of the forward kernel's pieces (stage timelines of five instrumented builds), tanh hang hunt,
# GPU suite without the tanh cases, quick headline.
set -u
mkdir -p gpurun_out
for v in new epi1_old loss_old prologue_old all_old; do
  lib=rl_games_b200/libb200rl_timing_$v.so; [ "$v" = "new" ] && lib=rl_games_b200/libb200rl_timing.so
  echo "== stage timeline: variant $v =="
  B200RL_LIB_PATH=$PWD/$lib timeout 120 python tools/tc_stage_timing.py 2>&1 | grep -v "^iter 0" | head -64 | tee gpurun_out/r02_c3_stage_$v.log | grep -E "iter|\[fwd\]|\[bwd|flush|reduce_adam"
done
echo "== tanh cases, blocking launches, per-test timeout with stack dump =="
CUDA_LAUNCH_BLOCKING=1 timeout 200 python -X faulthandler -m pytest tests/test_tc_faithful_gpu.py -x -q -k "tanh" --timeout 60 -s 2>&1 | tail -60 | tee gpurun_out/r02_c3_tanh.log | tail -40
echo "== gpu suite without tanh cases =="
timeout 1000 python -m pytest tests -m gpu -q -x --timeout 300 --durations=12 -k "not tanh" 2>&1 | tail -40 | tee gpurun_out/r02_c3_gpu_tests.log
echo "== headline (c2 only) =="
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-secondary --skip-e2e 2>/dev/null | tee gpurun_out/r02_c3_bench.json | cut -c1-300
