"""GPU parity tests: every CUDA kernel (through the C ABI, via rl_games_b200.ops) against the CPU oracle
(oracle/ppo_oracle.py, itself pinned to the real reference by tests/golden) on the same seeded inputs.

Tolerances (stated per test): GAE / returns / normalise: BIT-EXACT vs the reference's fp32 loop;
reductions (moments, losses, grads): fp32 rtol 1e-5..1e-4 (different summation order, fp64 accumulators).
"""
import math
import os

import pytest
import torch

from oracle import ppo_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
DEV = 'cuda'


@pytest.fixture(scope='module')
def ops():
    from rl_games_b200 import ops as _ops
    return _ops


def load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


# ------------------------------------------------------------------------------------------ GAE
def test_gae_golden_bitexact(ops):
    for c in load('gae.pt'):
        inp = [t.to(DEV) for t in c['inputs']]
        out = ops.compute_gae(*inp, c['gamma'], c['tau'])
        assert torch.equal(out.cpu(), c['pytorch_gae']), c['shape']
        # uint8 dones, as stored by the experience buffer
        inp_u8 = [inp[0], inp[1], inp[2].to(torch.uint8), inp[3], inp[4].to(torch.uint8)]
        out2 = ops.compute_gae(*inp_u8, c['gamma'], c['tau'])
        assert torch.equal(out2.cpu(), c['pytorch_gae'])


def test_gae_view_inputs(ops):
    # tests/test_triton_gae.py:83-99: last-dim slices + transposed dones
    g = torch.Generator().manual_seed(0)
    H, N, V = 12, 8, 2
    r = torch.randn(H, N, V + 1, generator=g)
    v = torch.randn(H, N, V + 1, generator=g)
    d = (torch.rand(H, N, generator=g) < 0.15).float()
    lv = torch.randn(N, V + 1, generator=g)
    ld = (torch.rand(N, generator=g) < 0.15).float()
    ref = O.gae(r[:, :, :V].contiguous(), v[:, :, :V].contiguous(), d, lv[:, :V].contiguous(), ld, 0.99, 0.95)
    rc, vc, dc = r.to(DEV)[:, :, :V], v.to(DEV)[:, :, :V], d.to(DEV).t().contiguous().t()
    assert not rc.is_contiguous() and not dc.is_contiguous()
    out = ops.compute_gae(rc, vc, dc, lv.to(DEV)[:, :V], ld.to(DEV), 0.99, 0.95)
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize('H,N,masked', [(16, 1000, False), (64, 4099, True), (5, 130, False), (33, 257, True)])
def test_gae_fused_bitexact_and_partials(ops, H, N, masked):
    g = torch.Generator().manual_seed(H * 1000 + N)
    r = torch.randn(H, N, generator=g)
    v = torch.randn(H, N, generator=g) * 2 + 1
    d = (torch.rand(H, N, generator=g) < 0.1).to(torch.uint8)
    lv = torch.randn(N, generator=g)
    ld = (torch.rand(N, generator=g) < 0.1).to(torch.uint8)
    m = (torch.rand(H, N, generator=g) < 0.8).float() if masked else None
    ref = O.gae(r.unsqueeze(2), v.unsqueeze(2), d.float(), lv.unsqueeze(1), ld.float(), 0.99, 0.95).squeeze(2)
    advs = torch.empty(H, N, device=DEV)
    rets = torch.empty(H, N, device=DEV)
    partials = torch.zeros(((N + 127) // 128, 8), dtype=torch.float64, device=DEV)
    nb = ops.gae_fused(r.to(DEV), v.to(DEV), d.to(DEV), lv.to(DEV), ld.to(DEV), None if m is None else m.to(DEV), advs,
                       rets, partials, 0.99, 0.95)
    assert nb == partials.shape[0]
    assert torch.equal(advs.cpu(), ref)
    assert torch.equal(rets.cpu(), ref + v)
    p = partials.sum(0).cpu()
    w = torch.ones(H, N, dtype=torch.float64) if m is None else m.double()
    ret = (ref + v).double()
    a2 = ((ref + v) - v).double()    # advantages = returns - values (a2c_common.py:1598)
    exp = torch.stack([w.sum(), (w * v.double()).sum(), (w * v.double() ** 2).sum(), (w * ret).sum(), (w * ret ** 2).sum(),
                       (w * a2).sum(), (w * a2 ** 2).sum()])
    torch.testing.assert_close(p[:7], exp, rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize('H,N,masked', [(16, 1024, False), (16, 16384, False), (32, 16384, True), (1, 16, False), (17, 48, True), (33, 4096 + 16, False),
                                        (64, 4096, True), (100, 208, False), (16, 148 * 4 * 64 + 64, False)])
def test_gae_tma_kernel_bitexact_vs_oracle_and_register_kernel(ops, H, N, masked):
    """the TMA-staged kernel (dense [H,N], N % 16 == 0: every shape BASELINE.json names) -- single chunk, two chunks in flight, the ring
    (H > 32), ragged last chunk, tail tile, both tile widths -- bit-exact against the oracle loop and against the register-chunk kernel;
    partial sums agree to fp64 summation order"""
    g = torch.Generator().manual_seed(H * 1000 + N)
    r = torch.randn(H, N, generator=g); v = torch.randn(H, N, generator=g) * 2 + 1
    d = (torch.rand(H, N, generator=g) < 0.1).to(torch.uint8)
    lv = torch.randn(N, generator=g); ld = (torch.rand(N, generator=g) < 0.1).to(torch.uint8)
    m = (torch.rand(H, N, generator=g) < 0.8).float() if masked else None
    ref = O.gae(r.unsqueeze(2), v.unsqueeze(2), d.float(), lv.unsqueeze(1), ld.float(), 0.99, 0.95).squeeze(2)
    dev = lambda t: None if t is None else t.to(DEV)   # noqa: E731
    res = {}
    try:
        for tma in (True, False):
            ops.gae_set_tma(tma)
            advs = torch.full((H, N), float('nan'), device=DEV); rets = torch.full((H, N), float('nan'), device=DEV)
            partials = torch.zeros(((N + 63) // 64, 8), dtype=torch.float64, device=DEV)
            nb = ops.gae_fused(dev(r), dev(v), dev(d), dev(lv), dev(ld), dev(m), advs, rets, partials, 0.99, 0.95)
            drop = ops.compute_gae(dev(r).unsqueeze(2), dev(v).unsqueeze(2), dev(d), dev(lv).unsqueeze(1), dev(ld), 0.99, 0.95)
            torch.cuda.synchronize()
            res[tma] = (advs.cpu(), rets.cpu(), partials[:nb].sum(0).cpu(), nb, drop.squeeze(2).cpu())
    finally:
        ops.gae_set_tma(True)
    for tma in (True, False):
        assert torch.equal(res[tma][0], ref) and torch.equal(res[tma][1], ref + v) and torch.equal(res[tma][4], ref), tma
    assert res[True][3] == (N + 63) // 64 if N <= 148 * 4 * 64 else res[True][3] == (N + 127) // 128
    torch.testing.assert_close(res[True][2], res[False][2], rtol=1e-12, atol=1e-9)


def test_gae_full_size_properties(ops):
    """BASELINE c5 per-GPU size (H=32, N=16384) and 1M envs: compare with the eager loop on the GPU
    (same op order => bit-exact) and check the all-done closed form A = r - V."""
    for H, N in [(32, 16384), (16, 1 << 20)]:
        g = torch.Generator(device=DEV).manual_seed(1)
        r = torch.randn(H, N, 1, device=DEV, generator=g)
        v = torch.randn(H, N, 1, device=DEV, generator=g)
        d = (torch.rand(H, N, device=DEV, generator=g) < 0.05).float()
        lv = torch.randn(N, 1, device=DEV, generator=g)
        ld = (torch.rand(N, device=DEV, generator=g) < 0.05).float()
        out = ops.compute_gae(r, v, d, lv, ld, 0.99, 0.95)
        assert torch.equal(out, O.gae(r, v, d, lv, ld, 0.99, 0.95))
        ones = torch.ones_like(d)
        out1 = ops.compute_gae(r, v, ones, lv, torch.ones_like(ld), 0.99, 0.95)
        assert torch.equal(out1, r - v)


# ------------------------------------------------------------------------------------------ prepare / stats
@pytest.mark.parametrize('masked', [False, True])
def test_prepare_batch_vs_oracle(ops, masked):
    g = torch.Generator().manual_seed(7)
    H, N = 16, 520
    B = H * N
    r = torch.randn(H, N, generator=g) * 0.5
    v = torch.randn(H, N, generator=g) * 3 + 10
    d = (torch.rand(H, N, generator=g) < 0.1).to(torch.uint8)
    lv = torch.randn(N, generator=g) * 3 + 10
    ld = (torch.rand(N, generator=g) < 0.1).to(torch.uint8)
    m = (torch.rand(H, N, generator=g) < 0.85).float() if masked else None
    # oracle: prepare_dataset on flat [env*H+t] tensors
    advs_ref = O.gae(r.unsqueeze(2), v.unsqueeze(2), d.float(), lv.unsqueeze(1), ld.float(), 0.99, 0.95)
    ag = O.OracleAgent.__new__(O.OracleAgent)
    ag.cfg = dict(O.DEFAULT_CFG)
    ag.model = O.OracleModel(O.init_params(4, [8], 2), 4, [8], 2)
    ag.model.value_mean_std.load({'running_mean': torch.tensor([0.7]), 'running_var': torch.tensor([2.5]), 'count': torch.tensor(123)})
    batch = {'returns': O.swap_and_flatten01(advs_ref + v.unsqueeze(2)), 'values': O.swap_and_flatten01(v.unsqueeze(2)),
             'neglogpacs': torch.zeros(B), 'actions': torch.zeros(B, 2), 'obses': torch.zeros(B, 4), 'dones': torch.zeros(B),
             'mus': torch.zeros(B, 2), 'sigmas': torch.ones(B, 2)}
    if masked:
        batch['rnn_masks'] = O.swap_and_flatten01(m)
    ag.prepare_dataset(batch)
    # CUDA
    advs = torch.empty(H, N, device=DEV); rets = torch.empty(H, N, device=DEV)
    partials = torch.zeros(((N + 127) // 128, 8), dtype=torch.float64, device=DEV)
    md = None if m is None else m.to(DEV)
    nb = ops.gae_fused(r.to(DEV), v.to(DEV), d.to(DEV), lv.to(DEV), ld.to(DEV), md, advs, rets, partials, 0.99, 0.95)
    mean = torch.tensor([0.7], dtype=torch.float64, device=DEV); var = torch.tensor([2.5], dtype=torch.float64, device=DEV)
    cnt = torch.tensor([123], dtype=torch.int64, device=DEV)
    ov, rn, an = (torch.empty(H, N, device=DEV) for _ in range(3))
    ops.prepare_batch(v.to(DEV), rets, md, partials, nb, mean, var, cnt, ov, rn, an, True, True)
    vms = ag.model.value_mean_std
    assert int(cnt) == int(vms.count)
    torch.testing.assert_close(mean.cpu(), vms.running_mean, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(var.cpu(), vms.running_var, rtol=1e-5, atol=1e-7)
    fl = O.swap_and_flatten01
    torch.testing.assert_close(fl(ov.cpu().unsqueeze(2)), ag.dataset['old_values'], rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(fl(rn.cpu().unsqueeze(2)), ag.dataset['returns'], rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(fl(an.cpu()), ag.dataset['advantages'], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('D,rpc,nch', [(60, 2048, 4), (6, 37, 3), (256, 512, 2), (348, 100, 1)])
def test_moments_update_vs_oracle(ops, D, rpc, nch):
    g = torch.Generator().manual_seed(D)
    N = rpc * 3
    x = torch.randn(nch, N, D, generator=g) * 2.5 + torch.arange(D).float() * 0.1
    rms = O.RunningMeanStd((D,))
    rms.load({'running_mean': torch.randn(D, generator=g).double(), 'running_var': torch.rand(D, generator=g).double() + 0.5,
              'count': torch.tensor(1000)})
    mean = rms.running_mean.clone().to(DEV); var = rms.running_var.clone().to(DEV)
    cnt = torch.tensor([1000], dtype=torch.int64, device=DEV)
    mf, sf = torch.empty(D, device=DEV), torch.empty(D, device=DEV)
    scratch = torch.zeros(1024 * 2 * D, dtype=torch.float64, device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    xd = x.to(DEV)
    e0 = rpc   # minibatch = envs [rpc, 2*rpc) of every time step
    for _ in range(2):   # twice: checks counter reset + sequential merges
        rms.train()
        rms(x[:, e0:e0 + rpc].reshape(-1, D))
        ops.moments_update(xd[0, e0:], D, rpc, nch, N, mean, var, cnt, mf, sf, scratch, counter)
    assert int(cnt) == int(rms.count) and int(counter) == 0
    torch.testing.assert_close(mean.cpu(), rms.running_mean, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(var.cpu(), rms.running_var, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(mf.cpu(), rms.running_mean.float(), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(sf.cpu(), torch.sqrt(rms.running_var.float() + 1e-5), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('D', [60, 256, 8])
def test_obs_mb_moments_then_merge_vs_oracle(ops, D):
    """once-per-epoch batch sums of every minibatch + per-minibatch Chan merge == sequential training-mode updates"""
    g = torch.Generator().manual_seed(D + 1)
    H, N, epm = 4, 192, 64
    x = torch.randn(H, N, D, generator=g) * 2.0 + torch.arange(D).float() * 0.05
    rms = O.RunningMeanStd((D,))
    rms.load({'running_mean': torch.randn(D, generator=g).double(), 'running_var': torch.rand(D, generator=g).double() + 0.5,
              'count': torch.tensor(500)})
    mean = rms.running_mean.clone().to(DEV); var = rms.running_var.clone().to(DEV)
    cnt = torch.tensor([500], dtype=torch.int64, device=DEV)
    mf, sf = torch.empty(D, device=DEV), torch.empty(D, device=DEV)
    n_mb = N // epm
    mbmom = torch.zeros(n_mb, 2 * D, dtype=torch.float64, device=DEV); shift = torch.empty(D, device=DEV)
    scratch = torch.zeros(1024 * 2 * D, dtype=torch.float64, device=DEV); counters = torch.zeros(n_mb, dtype=torch.int32, device=DEV)
    xd = x.to(DEV)
    for rep in range(2):      # two "mini-epochs" reuse the same precomputed sums
        if rep == 0:
            ops.obs_mb_moments(xd, D, H, N, epm, mean, mbmom, shift, scratch, counters)
            assert int(counters.sum()) == 0
        for i in range(n_mb):
            rms.train()
            rms(x[:, i * epm:(i + 1) * epm].reshape(-1, D))
            ops.obs_stats_merge(mbmom[i], shift, D, H * epm, mean, var, cnt, mf, sf)
    assert int(cnt) == int(rms.count)
    torch.testing.assert_close(mean.cpu(), rms.running_mean, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(var.cpu(), rms.running_var, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(sf.cpu(), torch.sqrt(rms.running_var.float() + 1e-5), rtol=1e-5, atol=1e-6)


def test_normalize_bitexact(ops):
    g = load('math.pt')['rms']
    mean, var = g['state']['running_mean'].to(DEV), g['state']['running_var'].to(DEV)
    assert torch.equal(ops.normalize(g['x_eval'].to(DEV), mean, var).cpu(), g['y_eval'])
    assert torch.equal(ops.normalize(g['x_eval'].to(DEV), mean, var, denorm=True).cpu(), g['y_denorm'])


# ------------------------------------------------------------------------------------------ fp32 MLP blocks
@pytest.mark.parametrize('M,K,N', [(300, 60, 256), (1024, 256, 128), (77, 128, 64), (130, 13, 9)])
def test_linear_blocks_vs_torch(ops, M, K, N):
    g = torch.Generator().manual_seed(M + K + N)
    X = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / math.sqrt(K); b = torch.randn(N, generator=g)
    nm = torch.randn(K, generator=g) * 0.1; ns = torch.rand(K, generator=g) + 0.5
    Xn = torch.clamp((X - nm) / ns, -5, 5)
    Y_ref = torch.nn.functional.elu(Xn.double() @ W.double().t() + b.double()).float()
    Y = torch.empty(M, N, device=DEV)
    ops.linear_fwd(X.to(DEV), W.to(DEV), b.to(DEV), Y, 1, norm_mean=nm.to(DEV), norm_std=ns.to(DEV))
    torch.testing.assert_close(Y.cpu(), Y_ref, rtol=2e-5, atol=2e-5)
    dY = torch.randn(M, N, generator=g)
    A_prev = torch.nn.functional.elu(torch.randn(M, K, generator=g))
    dX_ref = ((dY.double() @ W.double()) * torch.where(A_prev > 0, torch.ones_like(A_prev), A_prev + 1).double()).float()
    dX = torch.empty(M, K, device=DEV)
    ops.linear_bwd_data(dY.to(DEV), W.to(DEV), A_prev.to(DEV), dX, 1)
    torch.testing.assert_close(dX.cpu(), dX_ref, rtol=2e-5, atol=2e-5)
    S = 3
    part = torch.empty(S, N * K + N, device=DEV)       # unified split layout: [dW | db] per split
    ops.linear_bwd_weight(dY.to(DEV), X.to(DEV), part, part[:, N * K:], K, N, S, norm_mean=nm.to(DEV), norm_std=ns.to(DEV),
                          split_stride=N * K + N)
    red = torch.empty(N * K + N, device=DEV)
    ops.reduce_splits(part, red, N * K + N, S)
    torch.testing.assert_close(red[:N * K].cpu().view(N, K), (dY.double().t() @ Xn.double()).float(), rtol=2e-5, atol=1e-4)
    torch.testing.assert_close(red[N * K:].cpu(), dY.double().sum(0).float(), rtol=2e-5, atol=1e-4)


def test_linear_fwd_chunked_rows(ops):
    g = torch.Generator().manual_seed(3)
    H, N, K, Nout, rpc = 4, 96, 20, 16, 32
    X = torch.randn(H, N, K, generator=g); W = torch.randn(Nout, K, generator=g); b = torch.randn(Nout, generator=g)
    e0 = 32
    Xs = X[:, e0:e0 + rpc].reshape(-1, K)
    Y = torch.empty(H * rpc, Nout, device=DEV)
    ops.linear_fwd(X.to(DEV)[0, e0:], W.to(DEV), b.to(DEV), Y, 0, rows_per_chunk=rpc, chunk_stride=N, M=H * rpc)
    torch.testing.assert_close(Y.cpu(), Xs @ W.t() + b, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------ loss head
def _loss_case(ops, cfg_over, masked, seed=11):
    from rl_games_b200.ops import LossCfg
    g = torch.Generator().manual_seed(seed)
    M, A, Hl = 300, 5, 24
    a_last = torch.nn.functional.elu(torch.randn(M, Hl, generator=g))
    Wh = torch.randn(A + 1, Hl, generator=g) * 0.3; bh = torch.randn(A + 1, generator=g) * 0.1
    logstd = torch.randn(A, generator=g) * 0.2
    actions = torch.randn(M, A, generator=g)
    old_mu = torch.randn(M, A, generator=g) * 0.5; old_sigma = torch.rand(M, A, generator=g) + 0.5
    old_v = torch.randn(M, 1, generator=g); ret = torch.randn(M, 1, generator=g)
    adv = torch.randn(M, generator=g)
    mask = (torch.rand(M, generator=g) < 0.7).float() if masked else None
    cfg = dict(O.DEFAULT_CFG); cfg.update(cfg_over)
    entropy_coef = cfg['entropy_coef']
    # ---- oracle (autograd) ----
    al = a_last.clone().requires_grad_(True); W = Wh.clone().requires_grad_(True); b = bh.clone().requires_grad_(True)
    ls = logstd.clone().requires_grad_(True)
    head = al @ W.t() + b
    value, mu = head[:, :1], head[:, 1:]
    sigma = torch.exp(mu * 0 + ls)
    nlp = O.neglogp_fn(actions, mu, sigma, mu * 0 + ls)
    old_nlp = (nlp + torch.randn(M, generator=g) * 0.3).detach()
    ag = O.OracleAgent.__new__(O.OracleAgent); ag.cfg = cfg; ag.entropy_coef = entropy_coef
    ent = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma)).sum(-1)
    mb = {'old_logp_actions': old_nlp, 'advantages': adv, 'old_values': old_v, 'returns': ret, 'rnn_masks': mask}
    loss, a_l, c_l, e_l, b_l = ag.calc_losses(mb, {'prev_neglogp': nlp, 'values': value, 'entropy': ent, 'mus': mu})
    loss.backward()
    kl = O.policy_kl(mu.detach(), sigma.detach(), old_mu, old_sigma, reduce=mask is None)
    if mask is not None:
        kl = (kl * mask).sum() / mask.sum().clamp(min=1.0)
    # ---- CUDA ----
    c = LossCfg(cfg['e_clip'], cfg['critic_coef'], cfg['bounds_loss_coef'] or 0.0, int(cfg['bounds_loss_coef'] is not None),
                {'bound': 1, 'regularisation': 2}.get(cfg['bound_loss_type'], 0), int(cfg['clip_value']),
                int(cfg['use_smooth_clamp']), int(cfg['ppo']))
    stride = ops.loss_partial_stride()
    partials = torch.zeros(((M + 127) // 128, stride), dtype=torch.float64, device=DEV)
    d_head = torch.empty(M, A + 1, device=DEV); d_al = torch.empty(M, Hl, device=DEV)
    omu, osg = old_mu.to(DEV).clone(), old_sigma.to(DEV).clone()
    inv = None
    if mask is not None:
        inv = torch.tensor([1.0 / max(float(mask.sum()), 1.0)], device=DEV)
    mu_o = torch.empty(M, A, device=DEV); v_o = torch.empty(M, device=DEV); nlp_o = torch.empty(M, device=DEV)
    nb = ops.ppo_head_loss(a_last.to(DEV), Wh.to(DEV), bh.to(DEV), logstd.to(DEV), actions.to(DEV), omu, osg,
                           old_v.squeeze(1).to(DEV), ret.squeeze(1).to(DEV), old_nlp.to(DEV), adv.to(DEV),
                           None if mask is None else mask.to(DEV), M, 0, M, A, c, inv, d_head, d_al, 1, partials,
                           mu_out=mu_o, value_out=v_o, neglogp_out=nlp_o)
    stats = torch.zeros(16, device=DEV); dls = torch.empty(A, device=DEV)
    ops.ppo_loss_finalize(partials, nb, A, torch.tensor([entropy_coef], device=DEV), stats, dls)
    st = stats.cpu()
    torch.testing.assert_close(mu_o.cpu(), mu.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(nlp_o.cpu(), nlp.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(st[0], a_l.detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(st[1], c_l.detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(st[2], e_l.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(st[3], b_l.detach().reshape(()), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(st[4], kl, rtol=1e-4, atol=1e-6)
    # gradients: d_head == dL/d(head) ; d_alast == dL/d(pre-activation of last layer) = dL/da_last * elu'(a_last)
    dhead_ref = torch.autograd.grad  # noqa (kept for readability)
    dl_dal = al.grad
    elu_d = torch.where(a_last > 0, torch.ones_like(a_last), a_last + 1)
    torch.testing.assert_close(d_al.cpu(), dl_dal * elu_d, rtol=2e-4, atol=1e-7)
    torch.testing.assert_close(dls.cpu(), ls.grad, rtol=2e-4, atol=1e-6)
    # dW_head via d_head^T a_last must equal autograd
    torch.testing.assert_close(d_head.cpu().t() @ a_last, W.grad, rtol=2e-4, atol=1e-6)
    torch.testing.assert_close(d_head.cpu().sum(0), b.grad, rtol=2e-4, atol=1e-6)
    # mu/sigma write-back (datasets.py:33-43)
    torch.testing.assert_close(omu.cpu(), mu.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(osg.cpu(), sigma.detach(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('masked', [False, True])
def test_loss_head_smooth_reg(ops, masked):
    _loss_case(ops, {}, masked)


@pytest.mark.parametrize('masked', [False, True])
def test_loss_head_hardclip_bound_entropy(ops, masked):
    _loss_case(ops, {'use_smooth_clamp': False, 'bound_loss_type': 'bound', 'bounds_loss_coef': 0.01, 'entropy_coef': 0.02,
                     'clip_value': False}, masked, seed=12)


def test_mask_inv_counts(ops):
    g = torch.Generator().manual_seed(2)
    H, N, epm = 8, 96, 32
    m = (torch.rand(H, N, generator=g) < 0.5).float()
    m[:, 32:64] = 0
    inv = torch.empty(N // epm, device=DEV)
    ops.mask_inv_counts(m.to(DEV), H, N, epm, inv)
    ref = torch.stack([1.0 / m[:, i * epm:(i + 1) * epm].sum().clamp(min=1.0) for i in range(N // epm)])
    torch.testing.assert_close(inv.cpu(), ref, rtol=1e-6, atol=0)


# ------------------------------------------------------------------------------------------ adam
@pytest.mark.parametrize('truncate,wd', [(True, 0.0), (False, 0.01)])
def test_adam_step_vs_oracle(ops, truncate, wd):
    from rl_games_b200.ops import OptCfg
    g = torch.Generator().manual_seed(5)
    n = 57361
    p0 = torch.randn(n, generator=g) * 0.1
    params = [p0.clone()]
    opt = O.Adam(params, 3e-4, eps=1e-8, weight_decay=wd)
    sched = O.AdaptiveScheduler(0.008)
    pd = p0.clone().to(DEV); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    state = torch.tensor([3e-4, 0.0, 0.0, 0.0], dtype=torch.float64, device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    cfg = OptCfg(0.9, 0.999, 1e-8, wd, 1.0, 0.008, 1e-6, 1e-2, 1.5, 0.5, int(truncate), 1)
    stats = torch.zeros(16, device=DEV)
    lr = 3e-4
    for it, klv in enumerate([0.001, 0.05, 0.01, 0.0001, 0.2]):
        grad = torch.randn(n, generator=g) * (0.02 if it % 2 else 0.001)
        gs = [grad * 0.5]          # grad_scale = 1/world_size = 0.5
        if truncate:
            gs, norm = O.clip_grad_norm(gs, 1.0)
        opt.lr = lr
        opt.step(gs)
        lr, _ = sched.update(lr, 0.0, 0, 0, float(torch.tensor(klv, dtype=torch.float32)))
        ops.adam_step(pd, grad.to(DEV), m, v, state, torch.tensor([klv], device=DEV), cfg, stats, counter)
        torch.cuda.synchronize()
        assert float(state[0]) == pytest.approx(lr, rel=1e-14)
        assert int(state[1]) == it + 1
        torch.testing.assert_close(pd.cpu(), params[0], rtol=1e-5, atol=1e-7)
        if truncate:
            assert float(stats[8]) == pytest.approx(float(norm), rel=1e-5)
    torch.testing.assert_close(m.cpu(), opt.m[0], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(v.cpu(), opt.v[0], rtol=5e-5, atol=1e-10)


@pytest.mark.parametrize('kernel', ['adam_step', 'reduce_adam'])
def test_per_mini_epoch_scheduler_modes(ops, kernel):
    """cfg.adaptive_lr 2 / 3 (schedule_type 'standard', a2c_common.py:1565-1571): the LR moves only at the last minibatch of a
    mini-epoch, on the mean of the mini-epoch's KLs; the accumulators return to zero"""
    from rl_games_b200.ops import OptCfg
    g = torch.Generator().manual_seed(11)
    n, A, nmb = 4099, 8, 3
    stride = ops.loss_partial_stride()
    sched = O.AdaptiveScheduler(0.008)
    cfg = OptCfg(0.9, 0.999, 1e-8, 0.0, 1.0, 0.008, 1e-6, 1e-2, 1.5, 1.0, 1, 2)
    p = (torch.randn(n, generator=g) * 0.1).to(DEV); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    state = torch.tensor([3e-4] + [0.0] * 7, dtype=torch.float64, device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV); stats = torch.zeros(16, device=DEV)
    grad = torch.zeros(n, device=DEV); klt = torch.zeros(1, device=DEV); ec = torch.tensor([0.0], device=DEV)
    nrm = torch.zeros(148, dtype=torch.float64, device=DEV); bar = torch.zeros(1, dtype=torch.int64, device=DEV)
    lr = 3e-4
    kl_rows = [[0.001, 0.002, 0.0015], [0.05, 0.03, 0.001], [0.006, 0.007, 0.008], [0.0001, 0.03, 0.0001]]
    for me, row in enumerate(kl_rows):
        seen = []
        for i, klv in enumerate(row):
            cfg.adaptive_lr = 3 if i == nmb - 1 else 2
            if kernel == 'adam_step':
                klt.fill_(klv)
                ops.adam_step(p, (torch.randn(n, generator=g) * 0.01).to(DEV), m, v, state, klt, cfg, stats, counter)
            else:
                part = (torch.randn(5, n, generator=g) * 0.01).to(DEV)
                lpart = torch.zeros(4, stride, dtype=torch.float64, device=DEV)
                lpart[0, 4] = klv                     # slot 4 of a partial row = weighted KL contribution (loss_math.cuh sc[4])
                ops.reduce_adam(part, 5, n, lpart, 4, A, ec, stats, klt, grad, p, m, v, n, state, cfg, counter, nrm, bar)
            torch.cuda.synchronize()
            seen.append(float(klt[0]))
            if i < nmb - 1:
                assert float(state[0]) == lr and float(state[5]) == i + 1          # unchanged inside the mini-epoch
                assert float(state[4]) == pytest.approx(sum(seen), rel=1e-12)
        lr, _ = sched.update(lr, 0.0, 0, 0, sum(seen) / nmb)
        assert float(state[0]) == pytest.approx(lr, rel=1e-14)
        assert float(state[4]) == 0.0 and float(state[5]) == 0.0
        assert int(state[1]) == (me + 1) * nmb


@pytest.mark.parametrize('n,n_splits,pad', [(57361, 148, False), (57361, 148, True), (700, 5, False), (703, 7, True), (201737, 148, True)])
def test_reduce_adam_vs_two_kernel_path(ops, n, n_splits, pad):
    """fused reduce+finalise+clip+Adam launch == reduce_finalize followed by adam_step (same maths, different summation
    tree over the splits: fp32 rounding-level agreement)"""
    from rl_games_b200.ops import OptCfg
    g = torch.Generator().manual_seed(n)
    A = 8
    stride = ops.loss_partial_stride()
    cfg = OptCfg(0.9, 0.999, 1e-8, 0.0, 1.0, 0.008, 1e-6, 1e-2, 1.5, 1.0, 1, 1)
    p0 = (torch.randn(n, generator=g) * 0.1).to(DEV)
    ec = torch.tensor([0.01], device=DEV)
    st = {}
    for name in ('two', 'fused'):
        st[name] = dict(p=p0.clone(), m=torch.zeros(n, device=DEV), v=torch.zeros(n, device=DEV),
                        state=torch.tensor([3e-4, 0.0, 0.0, 0.0], dtype=torch.float64, device=DEV),
                        counter=torch.zeros(1, dtype=torch.int32, device=DEV), stats=torch.zeros(16, device=DEV),
                        grad=torch.zeros(n, device=DEV), kl=torch.zeros(1, device=DEV))
    nrm = torch.zeros(148, dtype=torch.float64, device=DEV); bar = torch.zeros(1, dtype=torch.int64, device=DEV)
    for it in range(4):
        pstride = (n + 3) // 4 * 4 if pad else n         # padded rows take the float4 path; the pad / [0, A) entries are never read
        part = torch.full((n_splits, pstride), float('nan'), device=DEV)
        part[:, A:n] = (torch.randn(n_splits, n - A, generator=g) * (0.05 if it % 2 else 0.002)).to(DEV)
        lpart = torch.rand(37, stride, generator=g, dtype=torch.float64).to(DEV) * (0.01 if it < 2 else 0.0005)
        a = st['two']
        ops.reduce_finalize(part[0, A:], a['grad'][A:], n - A, n_splits, pstride, lpart, 37, A, ec, a['stats'], a['grad'][:A], a['kl'])
        ops.adam_step(a['p'], a['grad'], a['m'], a['v'], a['state'], a['kl'], cfg, a['stats'], a['counter'])
        b = st['fused']
        ops.reduce_adam(part, n_splits, pstride, lpart, 37, A, ec, b['stats'], b['kl'], b['grad'], b['p'], b['m'], b['v'], n, b['state'], cfg,
                        b['counter'], nrm, bar)
        torch.cuda.synchronize()
        torch.testing.assert_close(b['grad'], a['grad'], rtol=1e-4, atol=1e-6)
        assert torch.equal(b['grad'][:A], a['grad'][:A]) and torch.equal(b['kl'], a['kl'])
        assert torch.equal(b['state'], a['state'])
        torch.testing.assert_close(b['stats'][:9], a['stats'][:9], rtol=1e-5, atol=0)
        torch.testing.assert_close(b['p'], a['p'], rtol=1e-5, atol=2e-7)
    torch.testing.assert_close(b['m'], a['m'], rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(b['v'], a['v'], rtol=2e-4, atol=1e-11)


@pytest.mark.parametrize('masked', [False, True])
def test_adv_ema_normalize_vs_oracle(ops, masked):
    """EMA advantage normaliser against oracle.GeneralizedMovingStats over a sequence of batches incl. an
    all-invalid one (no state update); the batch moments come from fp64 partial sums instead of torch's fp32 mean: 1e-6 agreement"""
    g = torch.Generator().manual_seed(17)
    gms = O.GeneralizedMovingStats((1,), decay=0.5)
    state = torch.zeros(2, device=DEV); step = torch.ones(1, dtype=torch.int32, device=DEV)
    for it, B in enumerate([4096, 1000, 777, 20000]):
        v = torch.randn(B, generator=g); r = v + torch.randn(B, generator=g) * (1 + it) + 0.2 * it
        if masked:
            mask = torch.zeros(B) if it == 2 else (torch.rand(B, generator=g) < 0.6).float()
        else:
            mask = None
        adv = r - v
        ref = gms(adv, mask=mask) if mask is not None else gms(adv)
        vd, rd = v.to(DEV), r.to(DEV)
        md = None if mask is None else mask.to(DEV)
        part = torch.zeros(64, 8, dtype=torch.float64, device=DEV)
        nb = ops.batch_moments(vd, rd, md, part)
        out = (rd - vd).clone()
        ops.adv_ema_normalize(out, part, nb, state, step, 0.5, training=True)
        torch.cuda.synchronize()
        torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(state.cpu(), torch.cat([gms.mean, gms.sqrs]), rtol=1e-6, atol=1e-7)
        assert int(step) == int(gms.step)
    # eval mode: normalise only
    x = torch.randn(513, generator=g) * 10
    gms.eval()
    out = x.to(DEV).clone()
    ops.adv_ema_normalize(out, part, nb, state, step, 0.5, training=False)
    torch.testing.assert_close(out.cpu(), gms(x), rtol=1e-5, atol=2e-6)
    assert int(step) == int(gms.step)


# ------------------------------------------------------------------------------------------ rollout pieces
def test_policy_head_sample_vs_oracle(ops):
    g = torch.Generator().manual_seed(9)
    N, A, Hl = 515, 8, 64
    a_last = torch.nn.functional.elu(torch.randn(N, Hl, generator=g))
    Wh = torch.randn(A + 1, Hl, generator=g) * 0.2; bh = torch.randn(A + 1, generator=g) * 0.1
    logstd = torch.randn(A, generator=g) * 0.3
    noise = torch.randn(N, A, generator=g)
    vm, vv = torch.tensor([1.5], dtype=torch.float64), torch.tensor([4.0], dtype=torch.float64)
    head = a_last @ Wh.t() + bh
    mu, val = head[:, 1:], head[:, :1]
    sigma = torch.exp(logstd).expand_as(mu)
    act = mu + sigma * noise
    nlp = O.neglogp_fn(act, mu, sigma, logstd.expand_as(mu))
    vms = O.RunningMeanStd((1,)); vms.load({'running_mean': vm, 'running_var': vv, 'count': torch.tensor(10)})
    val_d = vms(val, denorm=True)
    outs = {k: torch.empty(N, A, device=DEV) for k in ('actions', 'mus', 'sigmas', 'env')}
    nl = torch.empty(N, device=DEV); vals = torch.empty(N, device=DEV)
    dones_cur = (torch.rand(N, generator=g) < 0.3).to(torch.uint8).to(DEV)
    dones_out = torch.zeros(N, dtype=torch.uint8, device=DEV)
    prev = (torch.rand(N, generator=g) < 0.3).float().to(DEV); valid = torch.empty(N, device=DEV)
    lo = torch.full((A,), -2.0, device=DEV); hi = torch.full((A,), 4.0, device=DEV)
    ops.policy_head_sample(a_last.to(DEV), Wh.to(DEV), bh.to(DEV), logstd.to(DEV), vm.to(DEV), vv.to(DEV), True, noise.to(DEV),
                           123, None, 0, outs['actions'], outs['mus'], outs['sigmas'], nl, vals, outs['env'], True, lo, hi,
                           dones_cur, dones_out, prev, valid, N, A)
    torch.testing.assert_close(outs['mus'].cpu(), mu, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(outs['sigmas'].cpu(), sigma, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(outs['actions'].cpu(), act, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(nl.cpu(), nlp, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(vals.cpu(), val_d.squeeze(1), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(outs['env'].cpu(), torch.clamp(act, -1, 1) * 3.0 + 1.0, rtol=1e-5, atol=1e-5)
    assert torch.equal(dones_out, dones_cur) and torch.equal(valid, 1 - prev)
    # Philox path: moments of the implied noise
    N2 = 1 << 16
    a2 = torch.zeros(N2, Hl, device=DEV)
    o = {k: torch.empty(N2, A, device=DEV) for k in ('a', 'm', 's')}
    nl2 = torch.empty(N2, device=DEV); v2 = torch.empty(N2, device=DEV)
    ep = torch.zeros(1, dtype=torch.int64, device=DEV)
    ops.policy_head_sample(a2, Wh.to(DEV), bh.to(DEV), torch.zeros(A, device=DEV), None, None, False, None, 77, ep, 3,
                           o['a'], o['m'], o['s'], nl2, v2, None, False, None, None, None, None, None, None, N2, A)
    eps = (o['a'] - o['m'])
    assert abs(float(eps.mean())) < 0.01 and abs(float(eps.var()) - 1.0) < 0.02
    assert abs(float((eps ** 4).mean()) - 3.0) < 0.15
    # different step index => different draws; same index => identical (graph replay determinism)
    o2 = torch.empty(N2, A, device=DEV)
    ops.policy_head_sample(a2, Wh.to(DEV), bh.to(DEV), torch.zeros(A, device=DEV), None, None, False, None, 77, ep, 3,
                           o2, o['m'], o['s'], nl2, v2, None, False, None, None, None, None, None, None, N2, A)
    assert torch.equal(o2, o['a'])
    ops.policy_head_sample(a2, Wh.to(DEV), bh.to(DEV), torch.zeros(A, device=DEV), None, None, False, None, 77, ep, 4,
                           o2, o['m'], o['s'], nl2, v2, None, False, None, None, None, None, None, None, N2, A)
    assert not torch.equal(o2, o['a'])


def test_post_step_vs_oracle(ops):
    from rl_games_b200.ops import ShaperCfg
    g = torch.Generator().manual_seed(4)
    N, T = 700, 12
    cfg = ShaperCfg(0.5, 0.1, -float('inf'), float('inf'), 0.99, 0, 1)
    ep_state = torch.zeros(3, N, device=DEV); meter = torch.zeros(8, dtype=torch.float64, device=DEV)
    scratch = torch.zeros(64 * 4, dtype=torch.float64, device=DEV); counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    dones_cur = torch.zeros(N, dtype=torch.uint8, device=DEV); prev = torch.zeros(N, device=DEV)
    cr = torch.zeros(N, 1); cs = torch.zeros(N, 1); cl = torch.zeros(N)
    mr, msr, ml = O.AverageMeter(1, 100), O.AverageMeter(1, 100), O.AverageMeter(1, 100)
    for t in range(T):
        rew = torch.randn(N, generator=g); dn = (torch.rand(N, generator=g) < 0.08).to(torch.uint8)
        to = ((torch.rand(N, generator=g) < 0.5) & dn.bool())
        vals = torch.randn(N, generator=g)
        shaped = O.shape_rewards(rew.unsqueeze(1), 0.5, 0.1) + 0.99 * vals.unsqueeze(1) * to.unsqueeze(1).float()
        cr += rew.unsqueeze(1); cs += shaped; cl += 1
        idx = dn.nonzero(as_tuple=False)
        mr.update(cr[idx]); msr.update(cs[idx]); ml.update(cl[idx])
        nd = (1.0 - dn.float()).unsqueeze(1)
        cr *= nd; cs *= nd; cl *= nd.squeeze(1)
        out = torch.empty(N, device=DEV)
        ops.post_step(rew.to(DEV), dn.to(DEV), to.to(DEV), vals.to(DEV), None, out, dones_cur, prev, ep_state, meter, 100,
                      scratch, counter, N, cfg)
        torch.testing.assert_close(out.cpu(), shaped.squeeze(1), rtol=1e-6, atol=1e-6)
        assert torch.equal(dones_cur.cpu(), dn)
    torch.testing.assert_close(ep_state[0].cpu(), cr.squeeze(1), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ep_state[2].cpu(), cl, rtol=0, atol=0)
    mt = meter.cpu()
    assert mt[3] == mr.current_size
    torch.testing.assert_close(mt[0].float(), mr.mean.squeeze(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(mt[1].float(), msr.mean.squeeze(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(mt[2].float(), ml.mean.squeeze(), rtol=1e-5, atol=1e-5)


def test_synth_env_step(ops):
    N, D, A = 4096, 60, 8
    obs = torch.empty(N, D, device=DEV); rew = torch.empty(N, device=DEV)
    dones = torch.empty(N, dtype=torch.uint8, device=DEV); tos = torch.empty(N, dtype=torch.uint8, device=DEV)
    ep_t = torch.zeros(N, dtype=torch.int32, device=DEV); ep = torch.zeros(1, dtype=torch.int64, device=DEV)
    acts = torch.randn(N, A, device=DEV)
    nd = 0
    for t in range(120):
        ops.synth_env_step(acts, obs, rew, dones, tos, ep_t, N, D, A, 100, 0.01, 5, ep, t)
        nd += int(dones.sum())
        if t == 0:
            torch.testing.assert_close(rew, -(acts * acts).sum(-1), rtol=1e-5, atol=1e-5)
            assert abs(float(obs.mean())) < 0.02 and abs(float(obs.var()) - 1) < 0.03
    assert int(ep_t.max()) < 100
    assert 0.005 < nd / (120 * N) < 0.03


def test_fused_allreduce_adam_world1_matches_adam_step(ops):
    """The peer-memory all-reduce + clip + Adam kernel with a world of one rank (its own IPC buffer as the only peer)
    must reproduce b200rl_adam_step_f32 (only the fp64 summation order of the norm partials differs)."""
    from rl_games_b200.ops import OptCfg
    g = torch.Generator().manual_seed(8)
    n = 57361
    S = ((n + 1 + 3) // 4) * 4
    base, _ = ops.ipc_alloc(2 * S * 4 + 64)
    table = ops.PeerTable([[base + par * S * 4] for par in (0, 1)], [base + 2 * S * 4])
    comm = [ops.tensor_from_ptr(base + par * S * 4, n + 1, torch.float32, DEV) for par in (0, 1)]
    p0 = torch.randn(n, generator=g) * 0.1
    pa, pb = p0.clone().to(DEV), p0.clone().to(DEV)
    ma, va, mb_, vb = (torch.zeros(n, device=DEV) for _ in range(4))
    sa = torch.tensor([3e-4, 0, 0, 0], dtype=torch.float64, device=DEV); sb = sa.clone()
    ca, cb = torch.zeros(1, dtype=torch.int32, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV)
    cfg = OptCfg(0.9, 0.999, 1e-8, 0.0, 1.0, 0.008, 1e-6, 1e-2, 1.5, 1.0, 1, 1)
    seq = torch.zeros(1, dtype=torch.int64, device=DEV); red = torch.zeros(n + 1, device=DEV)
    nrm = torch.zeros(128, dtype=torch.float64, device=DEV); bar = torch.zeros(1, dtype=torch.int64, device=DEV)
    sta, stb = torch.zeros(16, device=DEV), torch.zeros(16, device=DEV)
    for it, klv in enumerate([0.001, 0.05, 0.01, 0.0001]):
        grad = torch.randn(n, generator=g) * (0.02 if it % 2 else 0.001)
        buf = comm[it & 1]
        buf[:n].copy_(grad.to(DEV)); buf[n:].fill_(klv)
        ops.adam_step(pa, buf[:n], ma, va, sa, buf[n:], cfg, sta, ca)
        ops.allreduce_adam(table, it & 1, 0, base + 2 * S * 4, seq, red, nrm, bar, pb, mb_, vb, n, sb, cfg, stb, cb)
        torch.cuda.synchronize()
        assert int(seq) == it + 1
        torch.testing.assert_close(sb, sa, rtol=1e-14, atol=0)
        torch.testing.assert_close(pb, pa, rtol=1e-5, atol=1e-7)
        assert float(stb[8]) == pytest.approx(float(sta[8]), rel=1e-6)
    torch.testing.assert_close(vb, va, rtol=1e-5, atol=1e-12)




@pytest.mark.parametrize('H,N', [(1, 7), (8, 520), (16, 4099)])
def test_rnn_train_dones_vs_reference_expression(ops, H, N):
    """a2c_common.py:1180-1191: rnn_dones[1:] = max(rnn_dones[1:], (mb_valid == 0)[:-1])"""
    g = torch.Generator().manual_seed(H + N)
    dones = (torch.rand(H, N, generator=g) < 0.3).to(torch.uint8)
    valid = (torch.rand(H, N, generator=g) < 0.7).float()
    out = torch.full((H, N), 9, dtype=torch.uint8, device=DEV)
    ops.rnn_train_dones(dones.to(DEV), valid.to(DEV), out)
    ref = dones.clone()
    ref[1:] = torch.maximum(ref[1:], (valid == 0.0)[:-1].to(torch.uint8))
    assert torch.equal(out.cpu(), ref)


def test_lr_schedule_apply_vs_oracle_scheduler(ops):
    from rl_games_b200.ops import OptCfg
    sched = O.AdaptiveScheduler(0.008, 1e-6, 1e-2, 1.5)
    cfg = OptCfg(0.9, 0.999, 1e-8, 0.0, 1.0, 0.008, 1e-6, 1e-2, 1.5, 0.5, 1, 0)       # adaptive_lr of the struct is ignored by this entry point
    state = torch.tensor([123.0, 7.0, 0.5, 0.25, 0.0, 0.0, 0.0, 0.0], dtype=torch.float64, device=DEV)
    for base, klv in ((3e-4, 0.001), (3e-4, 0.05), (3e-4, 0.01), (9e-3, 0.0001), (1.2e-6, 0.5)):
        kl = torch.tensor([klv * 2.0], device=DEV)          # summed over two ranks; kl_scale = 1/world
        ops.lr_schedule_apply(state, kl, 0.5, base, cfg)
        want, _ = sched.update(base, 0.0, 0, 0, float(torch.tensor(klv * 2.0, dtype=torch.float32)) * 0.5)
        assert float(state[0]) == pytest.approx(want, rel=1e-14)
        assert state[1:].tolist() == [7.0, 0.5, 0.25, 0.0, 0.0, 0.0, 0.0]
