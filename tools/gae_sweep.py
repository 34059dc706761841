"""GAE kernel roofline sweep (run on the B200 via gpurun; never under a profiler).

Algorithmic bytes: 13 B/element for the drop-in kernel with u8 dones (r 4 + V 4 + done 1 read, A 4 written), 17 B/element for
the fused kernel that also writes returns; the reference's Triton kernel is fed fp32 dones like its caller does (16 B/element).
Times with CUDA events on the launching stream, median of 20, two cache regimes per row:
  cold = L2 flushed (256 MiB write) before every launch -- what the HBM roofline is about;
  warm = 20 back-to-back launches inside one event pair / 20, L2-resident inputs -- what the kernel costs inside the epoch graph at the
         BASELINE shapes (4.5 / 9 MB working sets).
Comparators on the same box: the reference's own Triton kernel (rl_games/triton_kernels/gae_kernel.py:16-59 through _triton_gae, from the
vendored oracle/_ref) and its eager loop (_pytorch_gae) on the GPU."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_games_b200 import ops  # noqa: E402

PEAK = 6577.4
if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')):
    PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']


def time_kernel(fn, flush, iters=20, warm=3):
    """flush given: one launch per event pair, L2 flushed before it (the flush kernel is still running when the launch is enqueued, so no
    CPU launch latency is inside the pair).  flush None (L2-resident inputs): `iters` launches back to back inside ONE event pair, divided
    by iters -- stream-ordered launches overlap their launch latency, which is what the kernel costs inside a CUDA graph."""
    for _ in range(warm):
        fn()
    if flush is None:
        ts = []
        for _ in range(5):
            ops.fill_u32(flush_small, 1)          # something in front so the first launch is not a cold-start of the queue
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(iters):
                fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e) / iters)
        ts.sort()
        return ts[len(ts) // 2], ts[0]
    ts = []
    for _ in range(iters):
        ops.fill_u32(flush, 1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def reference_kernels():
    try:
        from oracle import ref_arm
        ref_arm.import_reference()
        from rl_games.triton_kernels.gae_kernel import _triton_gae, _pytorch_gae
        return _triton_gae, _pytorch_gae
    except Exception as e:      # oracle/_ref not vendored on this box
        print('reference kernels unavailable:', repr(e), file=sys.stderr)
        return None, None


def main():
    dev = 'cuda'
    global flush_small
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_small = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
    out = []
    quick = '--quick' in sys.argv
    shapes = [(16, 16384), (64, 4096), (32, 16384), (32, 131072), (16, 1 << 20), (32, 1 << 20), (64, 1 << 20), (32, 1 << 22)]
    if quick:
        shapes = [(16, 16384), (32, 16384), (32, 1 << 22)]
    triton_gae, eager_gae = reference_kernels()
    for H, N in shapes:
        r = torch.randn(H, N, device=dev); v = torch.randn(H, N, device=dev)
        d = (torch.rand(H, N, device=dev) < 0.05).to(torch.uint8)
        lv = torch.randn(N, device=dev); ld = (torch.rand(N, device=dev) < 0.05).to(torch.uint8)
        advs = torch.empty(H, N, device=dev); rets = torch.empty(H, N, device=dev)
        partials = torch.zeros(((N + 63) // 64, 8), dtype=torch.float64, device=dev)
        r3, v3, lv2 = r.unsqueeze(2), v.unsqueeze(2), lv.unsqueeze(1)
        df, ldf = d.float(), ld.float()
        el = H * N
        rows = []
        for tma in (True, False):
            tag = 'tma' if tma else 'regs'
            rows += [(f'gae_dropin_u8[{tag}]', tma, lambda: ops.compute_gae(r3, v3, d, lv2, ld, 0.99, 0.95), 13),
                     (f'gae_fused_returns_partials[{tag}]', tma, lambda: ops.gae_fused(r, v, d, lv, ld, None, advs, rets, partials, 0.99, 0.95), 17),
                     (f'gae_fused_returns[{tag}]', tma, lambda: ops.gae_fused(r, v, d, lv, ld, None, advs, rets, None, 0.99, 0.95), 17)]
        if triton_gae is not None:
            rows.append(('reference_triton_gae', True, lambda: triton_gae(r3, v3, df, lv2, ldf, 0.99, 0.95), 16))
            if el <= (1 << 22):
                rows.append(('reference_pytorch_loop_gae_on_gpu', True, lambda: eager_gae(r3, v3, df, lv2, ldf, 0.99, 0.95), 16))
        for name, tma, fn, bpe in rows:
            ops.gae_set_tma(tma)
            cold, cold_best = time_kernel(fn, flush)
            warm, warm_best = time_kernel(fn, None)
            gbs = el * bpe / (cold * 1e-3) / 1e9
            rec = {'kernel': name, 'H': H, 'N': N, 'elements': el, 'alg_bytes': el * bpe, 'ms_cold_median': cold, 'ms_cold_best': cold_best,
                   'ms_warm_median': warm, 'ms_warm_best': warm_best, 'GBs_cold': gbs, 'frac_of_measured_peak_cold': gbs / PEAK,
                   'GBs_warm': el * bpe / (warm * 1e-3) / 1e9}
            print(json.dumps(rec), flush=True)
            out.append(rec)
        ops.gae_set_tma(True)
        del r, v, d, lv, ld, advs, rets, partials, df, ldf
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump({'peak_gbs': PEAK, 'timing': 'CUDA events, median of 20 after 3 warm-up launches; not under a profiler', 'rows': out},
              open(os.path.join(ROOT, 'gpurun_out', 'r02_gae_sweep_quick.json' if quick else 'r02_gae_sweep.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
