"""torchrun worker of tests/test_multigpu_gpu.py (one process per GPU, NCCL): the fused peer-memory all-reduce + Adam kernel
(allreduce_adam_kernel, adam.cu) against ONE ncclAllReduce of the same flat gradient (+KL slot) followed by adam_step_kernel, the
reference's exchange (a2c_common.py:493-514), on the same minibatch of rank-different experience.  Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    rank, local_rank, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(local_rank)
    dev = f'cuda:{local_rank}'
    import test_agent_gpu as T
    from oracle import ppo_oracle as O
    T.DEV = dev
    N, H, D, A, units, mb = 1024, 8, 60, 8, [256, 128, 64], 4096
    mp = '--fp32' not in sys.argv
    params = O.init_params(D, units, A, seed=4)
    # every rank sees its own experience (different tapes and action noise), like sharded actors
    obs_tape, done_tape, tout_tape = O.make_tapes(8 * H + 1, N, D, seed=100 + rank)
    g = torch.Generator().manual_seed(50 + rank)
    noise = [torch.randn(H, N, A, generator=g).to(dev) for _ in range(8)]
    # '--clip': grad_norm small enough that the global-norm clip is ACTIVE (the default 1.0 never clips at this size).  The clip scale is
    # grad_norm / (||g|| + 1e-6) and the two optimiser kernels sum ||g||^2 in different orders (adam_step_kernel: every CTA over the whole
    # vector; allreduce_adam_kernel: per-CTA slices, then the partials) -> a last-bit difference in the scale, which Adam amplifies
    grad_norm = 0.05 if '--clip' in sys.argv else 1.0
    res, P = {}, None
    for mode, fused in (('nccl', False), ('fused', True)):
        env = T.TapeEnvGPU(obs_tape, done_tape, tout_tape, A)
        a = T.make_agent({'mixed_precision': mp, 'mini_epochs': 1, 'multi_gpu': True, 'device': dev, 'b200_cuda_graph': False,
                          'b200_fused_allreduce': fused, 'grad_norm': grad_norm}, N, H, D, A, units, mb, env, params)
        assert a.fused_allreduce == fused and a.world_size == world and a.use_tc == mp
        P = a.model.num_params
        # ---- one minibatch through the exchange ----
        a._rollout(noise[0])
        a._gae_and_prepare()
        a._minibatch_update(0, 0)
        torch.cuda.synchronize()
        red = (a.ar_red if fused else a._gv[0]['comm'])[:P + 1].clone()         # sum over ranks of the flat gradient, KL in the last slot
        w1 = a.model.flat.clone()
        a._minibatch_update(1, 1)
        torch.cuda.synchronize()
        w2 = a.model.flat.clone()
        # ---- then whole epochs: how a last-bit difference in the summation order grows under Adam ----
        growth = []
        for e in range(1, 7):
            a.epoch_num += 1
            a.train_epoch(noise=noise[e])
            growth.append(a.model.flat.clone())
        allw = [torch.empty_like(w2) for _ in range(world)]
        dist.all_gather(allw, a.model.flat)
        res[mode] = dict(red=red, w1=w1, w2=w2, growth=growth, same=all(torch.equal(allw[0], x) for x in allw[1:]), lr=a.last_lr)
        a._graph_update = a._graph_epoch = None
        del a
        torch.cuda.synchronize()
        dist.barrier()
    n, f = res['nccl'], res['fused']
    gscale = float(n['red'][:P].abs().max())
    out = {'world': world, 'P': P, 'mixed_precision': mp, 'grad_norm': grad_norm,
           'grad_max_abs': gscale,
           'grad_max_abs_diff': float((n['red'][:P] - f['red'][:P]).abs().max()),
           'grad_rel_l2': float((n['red'][:P] - f['red'][:P]).norm() / n['red'][:P].norm()),
           'kl_sum': [float(n['red'][P]), float(f['red'][P])],
           'w1_max_abs_diff': float((n['w1'] - f['w1']).abs().max()),
           'w1_frac_gt_1e-6': float(((n['w1'] - f['w1']).abs() > 1e-6).float().mean()),
           'w2_max_abs_diff': float((n['w2'] - f['w2']).abs().max()),
           'epochs_max_abs_diff': [float((x - y).abs().max()) for x, y in zip(n['growth'], f['growth'])],
           'epochs_rel_l2': [float((x - y).norm() / x.norm()) for x, y in zip(n['growth'], f['growth'])],
           'lr': [n['lr'], f['lr']], 'ranks_identical': [n['same'], f['same']]}
    if rank == 0:
        print('MGPU_PARITY ' + json.dumps(out), flush=True)
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == '__main__':
    main()
