"""GAE kernel roofline sweep (run on the B200 via gpurun).  Algorithmic bytes: 13 B/element for the
drop-in kernel with u8 dones (r 4 + V 4 + done 1 read, A 4 written), 17 B/element for the fused kernel that
also writes returns.  Times with CUDA events on the launching stream; L2 is flushed
(256 MiB write) between timed launches."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_games_b200 import ops  # noqa: E402

PEAK = 6577.4
if os.path.exists('MEASURED_PEAKS.json'):
    PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']


def time_kernel(fn, flush, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        ops.fill_u32(flush, 1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    dev = 'cuda'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    quick = '--quick' in sys.argv
    shapes = [(16, 16384), (64, 4096), (32, 16384), (32, 131072), (16, 1 << 20), (32, 1 << 20), (64, 1 << 20),
              (16, 1 << 22), (32, 1 << 22), (64, 1 << 22)]
    if quick:
        shapes = [(16, 16384), (32, 1 << 22)]
    for H, N in shapes:
        r = torch.randn(H, N, device=dev); v = torch.randn(H, N, device=dev)
        d = (torch.rand(H, N, device=dev) < 0.05).to(torch.uint8)
        lv = torch.randn(N, device=dev); ld = (torch.rand(N, device=dev) < 0.05).to(torch.uint8)
        advs = torch.empty(H, N, device=dev); rets = torch.empty(H, N, device=dev)
        partials = torch.zeros(((N + 127) // 128, 8), dtype=torch.float64, device=dev)
        r3, v3, lv2 = r.unsqueeze(2), v.unsqueeze(2), lv.unsqueeze(1)
        el = H * N
        for name, fn, bpe in [
            ('gae_dropin_u8', lambda: ops.compute_gae(r3, v3, d, lv2, ld, 0.99, 0.95), 13),
            ('gae_fused_returns_partials', lambda: ops.gae_fused(r, v, d, lv, ld, None, advs, rets, partials, 0.99, 0.95), 17),
            ('gae_fused_returns', lambda: ops.gae_fused(r, v, d, lv, ld, None, advs, rets, None, 0.99, 0.95), 17),
        ]:
            med, best = time_kernel(fn, flush)
            gbs = el * bpe / (med * 1e-3) / 1e9
            rec = {'kernel': name, 'H': H, 'N': N, 'elements': el, 'alg_bytes': el * bpe, 'ms_median': med, 'ms_best': best,
                   'GBs': gbs, 'frac_of_measured_peak': gbs / PEAK}
            print(json.dumps(rec), flush=True)
            out.append(rec)
        del r, v, d, lv, ld, advs, rets, partials
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/gae_sweep.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
