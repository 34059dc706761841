"""Per-shape timing of the layer-wise tcgen05 GEMMs (csrc/gemm_tc.cu) at the c4 workload's shapes (run on the B200 via gpurun; never
under a profiler).  CUDA events, median of 15, L2 flushed (256 MiB write) before every launch.  Reports TFLOP/s, and GB/s over the
ALGORITHMIC bytes of the call as the agent makes it (fp32 activations in and out, bf16 weight twin in) -- these GEMMs have fp32 I/O, so at
K <= 512 most of them are bound by HBM, not by the tensor pipe; `bound_us` is max(bytes / measured copy bandwidth, flops / measured bf16
peak).  Comparator on the same box: torch.matmul on bf16 copies of the operands (cuBLAS, bf16 in / bf16 out: half our input bytes, a
quarter of our output bytes, no bias / activation epilogue) -- a library lower bound, not the same contract."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_games_b200 import ops  # noqa: E402

HBM, TC = 6577.4, 1444.3
if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')):
    pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    HBM = pk.get('hbm_gbs', HBM)
    TC = pk.get('bf16_tflops_sustained', pk.get('bf16_tflops', TC))


def timed(fn, flush, iters=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        ops.fill_u32(flush, 1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = 'cuda'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # (label, M, K, N): the update minibatch (32768 rows, 4 LSTM steps of 8192 sequences) and the rollout step (8192 envs) of c4
    shapes = [('lstm_ih step', 8192, 348, 1024), ('lstm_hh step', 8192, 256, 1024), ('mlp1 update', 32768, 256, 512),
              ('mlp2 update', 32768, 512, 256), ('mlp3 update', 32768, 256, 128), ('mlp1 rollout', 8192, 256, 512),
              ('mlp3 rollout', 8192, 256, 128)]
    out = []
    for label, M, K, N in shapes:
        X = torch.randn(M, K, device=dev); dY = torch.randn(M, N, device=dev) * 0.1
        flat = torch.randn(N * K + 8, device=dev) / K ** 0.5
        fb = torch.empty(flat.numel(), dtype=torch.bfloat16, device=dev)
        ops.cast_bf16(flat, fb)
        W = flat[8:].view(N, K); b = torch.zeros(N, device=dev)
        Y = torch.empty(M, N, device=dev); dX = torch.empty(M, K, device=dev)
        # row splits as the agent uses them for c4: 16 per BPTT step for the gate GEMMs, all 64 partial rows for the MLP layers of the update
        splits = 64 if 'update' in label else max(1, min(16, M // 512))
        stride = N * K + N
        part = torch.empty(splits, stride, device=dev)
        Xb, Wb, dYb = X.bfloat16(), W.bfloat16().contiguous(), dY.bfloat16()
        flops = 2.0 * M * K * N
        calls = [
            ('fwd', lambda: ops.linear_fwd_tc(X, W, b, Y, 1, bf16_arena=(flat, fb)), 4 * M * K + 2 * N * K + 4 * M * N, lambda: Xb @ Wb.t()),
            ('dgrad', lambda: ops.linear_bwd_data_tc(dY, W, X, dX, 1, bf16_arena=(flat, fb)), 4 * M * N + 2 * N * K + 8 * M * K, lambda: dYb @ Wb),
            ('wgrad', lambda: ops.linear_bwd_weight_tc(dY, X, part, part[:, N * K:], K, N, splits, split_stride=stride),
             4 * M * N + 4 * M * K + 4 * splits * stride, lambda: dYb.t() @ Xb),
        ]
        for name, fn, nbytes, lib in calls:
            us = timed(fn, flush)
            us_lib = timed(lib, flush)
            bound = max(nbytes / (HBM * 1e3), flops / (TC * 1e6))
            rec = {'layer': label, 'op': name, 'M': M, 'K': K, 'N': N, 'us': round(us, 2), 'TFLOPs': round(flops / us / 1e6, 1),
                   'GBs_alg': round(nbytes / us / 1e3, 1), 'bound_us': round(bound, 2), 'frac_of_bound': round(bound / us, 3),
                   'cublas_bf16_us': round(us_lib, 2)}
            print(json.dumps(rec), flush=True)
            out.append(rec)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump({'hbm_gbs': HBM, 'bf16_tflops': TC, 'timing': 'CUDA events, median of 15, L2 flushed before every launch', 'rows': out},
              open(os.path.join(ROOT, 'gpurun_out', 'r02_gemm_tc_sweep.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
