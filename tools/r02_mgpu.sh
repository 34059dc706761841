#!/bin/bash
# Round-2 multi-GPU call: N = $1 GPUs of one box (gpurun --gpus N --timeout 900 -- 'bash tools/r02_mgpu.sh N [full|lean] [pytest -k expr]').
#   1. hardware parity at N ranks: fused peer all-reduce + Adam vs one ncclAllReduce + Adam on the same minibatch (tests/test_multigpu_gpu.py)
#   2. tools/mgpu_check.py (identical ranks, eager / graph / fused variants, per-epoch device time)
#   3. bench.py --gpus N (c2 line + c5 block), max over ranks on the device
#   4. rank 0 under a single-pass ncu duration list (no replay: safe with cross-rank flag waits) for the exchange kernels
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== parity test at world $N =="
timeout 500 python -m pytest tests/test_multigpu_gpu.py -q -x -k "${3:-test}" 2>&1 | tail -8 | tee gpurun_out/r02_mgpu${N}_parity_test.log
if [ "${2:-full}" = "full" ]; then
echo "== mgpu_check =="
timeout 300 $TR --master-port 29511 tools/mgpu_check.py 2>gpurun_out/r02_mgpu${N}_check.err | tail -1 | tee gpurun_out/r02_mgpu${N}_check.json | cut -c1-1500
fi
echo "== bench --gpus $N =="
timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/r02_bench_${N}gpu.err | tee gpurun_out/r02_bench_${N}gpu.json | cut -c1-400
if [ "${2:-full}" = "full" ]; then
echo "== bench --gpus $N, NCCL exchange instead of the fused peer kernel (A/B) =="
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --skip-secondary --cfg b200_fused_allreduce=False 2>/dev/null | tee gpurun_out/r02_bench_${N}gpu_nccl.json | cut -c1-300
fi
echo "== rank 0 duration list (single pass) =="
cat > /tmp/ncu_rank0.sh <<'EOS'
#!/bin/bash
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_mgpu_rank0.csv python "$@"
else
  exec python "$@"
fi
EOS
chmod +x /tmp/ncu_rank0.sh
timeout 400 $TR --master-port 29514 --no-python /tmp/ncu_rank0.sh bench.py --gpus $N --steps 1 --warmup 3 --no-graph --skip-secondary > gpurun_out/r02_mgpu${N}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_mgpu_rank0.csv 2>/dev/null | head -14 | tee gpurun_out/r02_mgpu${N}_launch_summary.txt
ls -la gpurun_out | tail -8
