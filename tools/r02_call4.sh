#!/bin/bash
# Round-2 GPU call 4 (one B200): locate the hang of the padded-geometry agent test (blocking launches + thread-method timeout dumps the
# Python stack at the stuck launch), prologue stamps, then the whole suite with durations.
set -u
mkdir -p gpurun_out
echo "== padded-geometry agent tests, one process each, blocking launches =="
for k in "17-units0-relu" "60-units1-tanh" "111-units2-elu" "60-units3-relu"; do
  echo "-- $k"
  CUDA_LAUNCH_BLOCKING=1 timeout 150 python -X faulthandler -m pytest "tests/test_agent_gpu.py::test_bf16_tcgen05_agent_other_geometries_and_activations_track_fp32_agent[$k]" \
      -x -q --timeout 60 --timeout-method=thread -s 2>&1 | grep -v "^$" | tail -45 | tee gpurun_out/r02_c4_geom_$k.log | grep -E "passed|failed|File \"|ops\.|Timeout|Error" | head -24
done
echo "== stage timeline (prologue stamps) =="
B200RL_LIB_PATH=$PWD/rl_games_b200/libb200rl_timing.so timeout 120 python tools/tc_stage_timing.py 2>&1 | grep -v "^iter 0" | head -70 | tee gpurun_out/r02_c4_stage.log | head -34
echo "== tanh faithful cases =="
timeout 200 python -m pytest tests/test_tc_faithful_gpu.py -q -k "tanh" --timeout 90 --timeout-method=thread 2>&1 | tail -5 | tee gpurun_out/r02_c4_tanh.log
echo "== gpu suite (padded-geometry agent tests deselected: run above), durations =="
timeout 1100 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method=thread --durations=25 -k "not other_geometries" 2>&1 | tail -60 | tee gpurun_out/r02_c4_gpu_tests.log
