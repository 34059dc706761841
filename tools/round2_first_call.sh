#!/bin/bash
# First GPU call of the next round (one B200).  Order: what must stay green first (validated suite, headline), then the code written after
# round 1's GPU budget ran out, least risky first -- a hung kernel in a late group cannot take the earlier results with it.
# Usage:  gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
set -u
mkdir -p gpurun_out
G="B200RL_UNVALIDATED=1"
echo "== validated suite =="
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_gpu_tests.log
echo "== headline =="
timeout 240 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e 2>/dev/null | tee gpurun_out/r02_bench_first.json | cut -c1-300
echo "== gated: validated kernels in new compositions / flag combinations (LSTM after the MLP, LSTM + autoreset masks, more config keys) =="
env $G timeout 300 python -m pytest tests/test_agent_gpu.py::test_lstm_after_mlp_agent_matches_reference_golden \
    tests/test_agent_gpu.py::test_lstm_on_next_step_autoreset_env_matches_reference_golden \
    tests/test_agent_gpu.py::test_agent_matches_reference_golden_more_config_keys tests/test_agent_gpu.py::test_masked_rows_contribute_zero_gradient tests/test_agent_gpu.py::test_bf16_tcgen05_agent_tracks_fp32_agent_flag_matrix -q 2>&1 | tail -30 | tee gpurun_out/r02_gated_compositions.log
echo "== gated: small new kernels (scheduler modes, lr_schedule_apply / resume, discrete PPO, central value) =="
env $G timeout 300 python -m pytest tests/test_kernels_gpu.py::test_per_mini_epoch_scheduler_modes tests/test_kernels_gpu.py::test_rnn_train_dones_vs_reference_expression tests/test_kernels_gpu.py::test_lr_schedule_apply_vs_oracle_scheduler \
    tests/test_agent_gpu.py::test_standard_schedule_agent_matches_reference_golden \
    tests/test_agent_gpu.py::test_resume_from_a_reference_checkpoint_continues_like_the_reference \
    tests/test_discrete_gpu.py tests/test_cv_gpu.py -q 2>&1 | tail -40 | tee gpurun_out/r02_gated_kernels.log
echo "== gated: wide-observation tcgen05 kernels (obs 256: BASELINE configs[4]) =="
env $G timeout 300 python -m pytest tests/test_mlp_tc_gpu.py tests/test_agent_gpu.py::test_bf16_tcgen05_wide_agent_tracks_fp32_agent -x -q 2>&1 | tail -40 | tee gpurun_out/r02_wide_tests.log
echo "== c5 on the wide tcgen05 path vs the fp32 path =="
timeout 240 python bench.py --workload c5 --steps 10 --warmup 3 --skip-cpu --skip-e2e --cfg b200_unvalidated=True 2>/dev/null | tee gpurun_out/r02_bench_c5_wide.json | cut -c1-300
timeout 240 python bench.py --workload c5 --steps 5 --warmup 3 --skip-cpu --skip-e2e --fp32 2>/dev/null | tee gpurun_out/r02_bench_c5_fp32.json | cut -c1-300
echo "== stage timelines incl. the fused reduce+Adam tail (instrumented build; run tools/tc_stage_timing.py --build beforehand) =="
if [ -f rl_games_b200/libb200rl_timing.so ]; then
  B200RL_LIB_PATH=$PWD/rl_games_b200/libb200rl_timing.so timeout 120 python tools/tc_stage_timing.py --old-bwd2 2>&1 | tail -80 | tee gpurun_out/r02_stage_timing.log
fi
