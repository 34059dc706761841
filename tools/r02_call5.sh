#!/bin/bash
# Round-2 GPU call 5 (one B200): prologue A/B, the whole -m gpu suite, smoke, the headline line (c2 + c5 block + e2e + reference CPU arm),
# the c4 (LSTM) line, then the profiling passes for profiles/ (launch list + one --set full capture of the step's main kernels).
set -u
mkdir -p gpurun_out
for v in new prologue_regs; do
  lib=rl_games_b200/libb200rl_timing_$v.so; [ "$v" = "new" ] && lib=rl_games_b200/libb200rl_timing.so
  echo "== stage timeline: $v =="
  B200RL_LIB_PATH=$PWD/$lib timeout 120 python tools/tc_stage_timing.py 2>&1 | grep -v "^iter 0" | head -70 | tee gpurun_out/r02_c5_stage_$v.log | grep -E "^iter|\[fwd\]|setup|loads|barrier|prologue"
done
echo "== gpu suite =="
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --durations=8 2>&1 | tail -30 | tee gpurun_out/r02_c5_gpu_tests.log
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_c5_smoke.log
echo "== headline =="
timeout 500 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_c5_bench.err | tee gpurun_out/r02_c5_bench.json | cut -c1-300
echo "== c4 (LSTM, fp32 kernels) =="
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --skip-cpu --skip-e2e 2>gpurun_out/r02_c5_bench_c4.err | tee gpurun_out/r02_c5_bench_c4.json | cut -c1-300
echo "== ncu: launch list of two eager c2 epochs =="
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-e2e --skip-secondary > gpurun_out/r02_c5_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_c2.csv 2>/dev/null | head -14 | tee gpurun_out/r02_c5_launch_summary.txt
echo "== ncu --set full: one launch of each main kernel of a c2 minibatch =="
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:gae_tma_kernel|mlp_fwd_tc_kernel.*Lb1ELb0|mlp_bwd_tc_kernel|reduce_adam_kernel|mlp_fwd_tc_kernel' --launch-skip 250 --launch-count 6 \
    -f -o gpurun_out/r02_prof_c2 python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-e2e --skip-secondary > gpurun_out/r02_c5_ncu_full.log 2>&1
ls -la gpurun_out | tail -6
