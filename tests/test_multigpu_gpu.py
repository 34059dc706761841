"""N > 1 on hardware (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py`; skipped on a one-GPU box): the
gradient exchange of this repo -- peer-memory all-reduce fused with clip + Adam (allreduce_adam_kernel) -- against the reference's
exchange, one NCCL all-reduce of the flat gradient (a2c_common.py:493-514), on the same rank-different minibatch:
  * the all-reduced gradient agrees to fp32 summation order (rel-L2 < 1e-6, max |diff| < 1e-5 x max |g|), the KL slot likewise;
  * the weights after that ONE optimiser step agree (Adam's first step is lr * g / (|g| + eps): only |g| ~ eps entries can move);
  * every rank holds bit-identical weights in both modes;
  * after whole epochs the two runs drift apart only at the rate the NCCL run drifts from itself under a last-bit perturbation
    (recorded, not asserted: profiles/r02_mgpu_parity.json) -- the 0.14 of round 1 is chaotic amplification of summation order by Adam
    at lr -> 1e-2, not an exchange error."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(world, extra=()):
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(HERE, '_mgpu_parity_worker.py'), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('MGPU_PARITY ')]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-3000:])
    out = json.loads(lines[-1][len('MGPU_PARITY '):])
    os.makedirs(os.path.join(os.path.dirname(HERE), 'gpurun_out'), exist_ok=True)
    with open(os.path.join(os.path.dirname(HERE), 'gpurun_out', f'mgpu_parity_w{world}{"".join(e.replace("--", "_") for e in extra)}.json'), 'w') as f:
        json.dump(out, f, indent=1)
    return out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs on one box (gpurun --gpus 2)')
@pytest.mark.parametrize('extra', [(), ('--fp32',)], ids=['bf16_tcgen05', 'fp32'])
def test_fused_peer_allreduce_matches_nccl_allreduce_one_minibatch(extra):
    world = 2 if torch.cuda.device_count() < 8 else 8
    o = _run(world, extra)
    assert o['world'] == world and o['ranks_identical'] == [True, True]
    assert o['grad_rel_l2'] < 1e-6 and o['grad_max_abs_diff'] <= 1e-5 * o['grad_max_abs'], o
    assert o['kl_sum'][0] == pytest.approx(o['kl_sum'][1], rel=1e-6), o
    assert o['w1_frac_gt_1e-6'] < 1e-3 and o['w1_max_abs_diff'] <= 2 * 3e-4 * 1.0001, o


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs on one box (gpurun --gpus 2)')
def test_divergence_of_whole_runs_comes_from_the_clip_scale_not_from_the_exchange():
    """round-1 finding: after 6 epochs the fused and the NCCL run differ by ~0.1 in some weights.  With the clip inactive (grad_norm 1.0 at this
    size) the two runs stay BIT-identical through six epochs at world 2 (a + b is order-free); with the clip active the all-reduced gradient
    is still identical, the first step differs by at most an ulp-sized change of the clip scale (two summation orders of ||g||^2), and that
    seed grows under Adam -- the exchange is exact, the optimiser tail's norm order is the only difference."""
    world = 2 if torch.cuda.device_count() < 8 else 8
    o = _run(world, ('--clip',))
    assert o['ranks_identical'] == [True, True]
    assert o['grad_rel_l2'] < 1e-6, o                        # the exchange itself
    assert o['w1_max_abs_diff'] <= 2e-6, o                   # one step: an ulp of the clip scale times lr
    if world == 2:
        plain = _run(world, ())
        assert plain['epochs_max_abs_diff'] == [0.0] * 6 and plain['grad_max_abs_diff'] == 0.0, plain
