#!/bin/bash
# Profiling pass of the c2 bench step (run under gpurun, ONE GPU).  Outputs land in gpurun_out/ (scratch); summaries are
# committed under profiles/ by hand.  Numbers printed by bench.py under ncu are never bench values.
#   1. launch list: every kernel of two eager (--no-graph) epochs with its device time (cold-cache, serialised: compare shares)
#   2. one --set full capture of the dominant kernels of one minibatch (GAE, forward+loss, backward, reduce+Adam) of the 4th epoch:
#      97 matching launches per epoch (1 GAE + 32 x 3), so --launch-skip 291 starts at the GAE of epoch 4
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --skip-cpu --skip-e2e > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on \
    -k 'regex:gae_fused_kernel|mlp_fwd_tc_kernel.*true|mlp_fwd_tc_kernel.*1>|mlp_bwd_tc_kernel|reduce_adam_kernel' --launch-skip 291 --launch-count 7 \
    -f -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-e2e > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -8
