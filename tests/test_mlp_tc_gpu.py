"""bf16 tcgen05 MLP kernels (fwd+loss, rollout fwd, backward) against the fp32 CUDA-core kernels on identical inputs.

Tolerance: bf16 operands (8-bit mantissa) with fp32 accumulation -- activations / heads atol 3e-2 of the O(1) values,
loss scalars rtol 3e-2, gradients: relative L2 error < 3e-2 and cosine > 0.999 per parameter tensor (the reference's
own bf16-autocast path has the same error class; a2c_continuous.py:173)."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'
D, UNITS, A = 60, [256, 128, 64], 8


def decode_tiles(buf, n_tiles, C):
    """tiled INTERLEAVE bf16 buffer -> dense [n_tiles*128, C] float"""
    r = torch.arange(128, device=buf.device).view(128, 1)
    c = torch.arange(C, device=buf.device).view(1, C)
    off = ((c // 8) * 2048 + (r // 8) * 128 + (r % 8) * 16 + (c % 8) * 2) // 2
    t = torch.arange(n_tiles, device=buf.device).view(n_tiles, 1, 1) * (C * 128)
    idx = (t + off.unsqueeze(0)).reshape(-1)
    return buf.view(torch.bfloat16)[idx].view(n_tiles * 128, C).float()


def make_net(g, D=D):
    ins, W, b = D, [], []
    for u in UNITS:
        W.append((torch.randn(u, ins, generator=g) / math.sqrt(ins)).to(DEV))
        b.append((torch.randn(u, generator=g) * 0.1).to(DEV))
        ins = u
    Wh = (torch.randn(A + 1, ins, generator=g) / math.sqrt(ins)).to(DEV)
    bh = (torch.randn(A + 1, generator=g) * 0.1).to(DEV)
    logstd = (torch.randn(A, generator=g) * 0.2).to(DEV)
    return W, b, Wh, bh, logstd


def rel_l2(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


def cosine(a, b):
    return float((a.flatten() @ b.flatten()) / (a.norm() * b.norm() + 1e-20))


@pytest.mark.parametrize('H,N,epm,masked', [(4, 512, 256, False), (2, 384, 128, True), (1, 1000, 1000, False)])
def test_tc_fwd_loss_bwd_vs_fp32(H, N, epm, masked):
    _fwd_loss_bwd_case(H, N, epm, masked, D)




@pytest.mark.parametrize('H,N,epm,masked,Dw', [(4, 512, 256, False, 256), (2, 384, 128, True, 105), (1, 1000, 1000, False, 65),
                                               (2, 19072, 19072, False, 256)])
def test_tc_wide_fwd_loss_bwd_vs_fp32(H, N, epm, masked, Dw):
    """BASELINE configs[4] geometry (obs 256) and ragged widths: layer 1 in l1_fwd_tc_kernel / l1_wgrad_tc_kernel, the chain kernels
    in their external-layer-1 form; the last case gives every CTA two or three tiles (298 tiles)"""
    _fwd_loss_bwd_case(H, N, epm, masked, Dw)


def _fwd_loss_bwd_case(H, N, epm, masked, D):
    from rl_games_b200 import ops
    from rl_games_b200.ops import LossCfg
    wide = D > 64
    assert ops.tc_kind(D, UNITS, A) == (2 if wide else 1)
    g = torch.Generator().manual_seed(H * 100 + N)
    W, b, Wh, bh, logstd = make_net(g, D)
    M = H * epm
    e0 = 128 if N > epm else 0
    obs = (torch.randn(H, N, D, generator=g) * 2 + 0.5).to(DEV)
    nm = (torch.randn(D, generator=g) * 0.3).to(DEV); ns = (torch.rand(D, generator=g) + 0.7).to(DEV)
    actions = torch.randn(H, N, A, generator=g).to(DEV)
    old_mu = (torch.randn(H, N, A, generator=g) * 0.5).to(DEV); old_sigma = (torch.rand(H, N, A, generator=g) + 0.5).to(DEV)
    old_v = torch.randn(H, N, generator=g).to(DEV); ret = torch.randn(H, N, generator=g).to(DEV)
    old_nlp = (torch.randn(H, N, generator=g) * 0.3 + 9.0).to(DEV); adv = torch.randn(H, N, generator=g).to(DEV)
    mask = (torch.rand(H, N, generator=g) < 0.7).float().to(DEV) if masked else None
    inv = None
    if masked:
        inv = torch.tensor([1.0 / max(float(mask[:, e0:e0 + epm].sum()), 1.0)], device=DEV)
    cfg = LossCfg(0.2, 2.0, 0.0, 1, 2, 1, 1, 1)
    ec = torch.tensor([0.0], device=DEV)
    stride = ops.loss_partial_stride()
    sl = lambda t: t[0, e0:]   # noqa: E731
    # ---------------- fp32 reference (CUDA-core kernels) ----------------
    ta = [torch.empty(M, u, device=DEV) for u in UNITS]
    ops.linear_fwd(sl(obs), W[0], b[0], ta[0], 1, rows_per_chunk=epm, chunk_stride=N, x_ld=D, norm_mean=nm, norm_std=ns, M=M)
    ops.linear_fwd(ta[0], W[1], b[1], ta[1], 1, M=M)
    ops.linear_fwd(ta[1], W[2], b[2], ta[2], 1, M=M)
    mu_r, sg_r = old_mu.clone(), old_sigma.clone()
    d_head = torch.empty(M, A + 1, device=DEV); dA = [torch.empty(M, u, device=DEV) for u in UNITS]
    partials = torch.zeros((M + 127) // 128, stride, dtype=torch.float64, device=DEV)
    nb = ops.ppo_head_loss(ta[2], Wh, bh, logstd, sl(actions), sl(mu_r), sl(sg_r), sl(old_v), sl(ret), sl(old_nlp), sl(adv),
                           None if mask is None else sl(mask), epm, N, M, A, cfg, inv, d_head, dA[2], 1, partials)
    stats_r = torch.zeros(16, device=DEV); dls_r = torch.empty(A, device=DEV)
    ops.ppo_loss_finalize(partials, nb, A, ec, stats_r, dls_r)
    gW = [torch.empty_like(w) for w in W]; gb = [torch.empty_like(x) for x in b]
    gWh = torch.empty_like(Wh); gbh = torch.empty_like(bh)

    def wgrad(dY, X, gw, gbias, **kw):
        n, k = gw.shape
        part = torch.empty(1, n * k + n, device=DEV)
        ops.linear_bwd_weight(dY, X, part, part[:, n * k:], k, n, 1, M=M, split_stride=n * k + n, **kw)
        gw.copy_(part[0, :n * k].view(n, k)); gbias.copy_(part[0, n * k:])
    wgrad(d_head, ta[2], gWh, gbh)
    wgrad(dA[2], ta[1], gW[2], gb[2])
    ops.linear_bwd_data(dA[2], W[2], ta[1], dA[1], 1, M=M)
    wgrad(dA[1], ta[0], gW[1], gb[1])
    ops.linear_bwd_data(dA[1], W[1], ta[0], dA[0], 1, M=M)
    wgrad(dA[0], sl(obs), gW[0], gb[0], rows_per_chunk=epm, chunk_stride=N, x_ld=D, norm_mean=nm, norm_std=ns)
    # ---------------- bf16 tcgen05 path ----------------
    n_tiles = (M + 127) // 128
    tb = ops.tc_tile_bytes(D, UNITS, A)
    wpack = torch.zeros(ops.tc_pack_bytes(D, UNITS, A), dtype=torch.uint8, device=DEV)
    ops.tc_pack_weights(W[0], W[1], W[2], Wh, D, UNITS, A, wpack)
    act = [torch.zeros(n_tiles * tb[i], dtype=torch.uint8, device=DEV) for i in range(3)]
    dhead = torch.zeros(n_tiles * tb[3], dtype=torch.uint8, device=DEV)
    delta2 = torch.zeros(n_tiles * tb[1], dtype=torch.uint8, device=DEV); delta1 = torch.zeros(n_tiles * tb[0], dtype=torch.uint8, device=DEV)
    mu_t, sg_t = old_mu.clone(), old_sigma.clone()
    partials_t = torch.zeros(148, stride, dtype=torch.float64, device=DEV)
    xt = torch.zeros(n_tiles * ops.tc_xtile_bytes(D, UNITS, A), dtype=torch.uint8, device=DEV)      # 128 x 64 (obs <= 64) or 128 x 256 bf16 per tile
    nbt = ops.tc_mlp_fwd_train(sl(obs), epm, N, D, nm, ns, wpack, b, bh, logstd, UNITS, M, A, sl(actions), sl(mu_t), sl(sg_t),
                               sl(old_v), sl(ret), sl(old_nlp), sl(adv), None if mask is None else sl(mask), cfg, inv, act, dhead,
                               partials_t, xtile=xt)
    stats_t = torch.zeros(16, device=DEV); dls_t = torch.empty(A, device=DEV)
    ops.ppo_loss_finalize(partials_t, nbt, A, ec, stats_t, dls_t)
    torch.cuda.synchronize()
    # activations
    for i, C in enumerate(UNITS):
        got = decode_tiles(act[i], n_tiles, C)[:M]
        assert rel_l2(got, ta[i]) < 1.5e-2, (i, rel_l2(got, ta[i]))
    # heads / loss
    torch.testing.assert_close(mu_t, mu_r, rtol=0, atol=4e-2)
    assert torch.equal(sg_t, sg_r)
    for k in (0, 1, 2, 3, 4):
        assert float(stats_t[k]) == pytest.approx(float(stats_r[k]), rel=4e-2, abs=2e-3), (k, stats_t[:8], stats_r[:8])
    assert float(stats_t[5]) == float(stats_r[5])
    dh_t = decode_tiles(dhead, n_tiles, 16)[:M, :A + 1]
    assert rel_l2(dh_t, d_head) < 6e-2 and cosine(dh_t, d_head) > 0.998, (rel_l2(dh_t, d_head), cosine(dh_t, d_head))
    assert rel_l2(dls_t, dls_r) < 5e-2
    # backward
    P = A + sum(w.numel() + x.numel() for w, x in zip(W, b)) + Wh.numel() + bh.numel()
    offs, o, ins = {}, A, D
    for i, u in enumerate(UNITS):
        offs[f'W{i}'] = o; o += u * ins
        offs[f'b{i}'] = o; o += u
        ins = u
    offs['W_head'] = o; o += (A + 1) * ins
    offs['b_head'] = o; o += A + 1
    assert o == P
    # normalised observation tile emitted by the forward (consumed by the pipelined weight-gradient kernel)
    xn = torch.clamp((torch.cat([obs[t, e0:e0 + epm] for t in range(H)]) - nm) / ns, -5.0, 5.0)
    xt_dec = decode_tiles(xt, n_tiles, 256 if wide else 64)[:M, :D]
    torch.testing.assert_close(xt_dec, xn.to(torch.bfloat16).float(), rtol=0, atol=4e-2)
    part = torch.full((148, P), float('nan'), device=DEV)
    # three editions of the layer-1 / layer-2 weight gradients must agree: X tiles from the forward (default), pipelined ring, re-derived X
    npart = ops.tc_mlp_bwd(sl(obs), epm, N, D, nm, ns, wpack, UNITS, M, A, act, dhead, delta2, delta1, part, P, offs, xtile=xt,
                           pipelined_wgrad=not wide)
    grad = torch.zeros(P, device=DEV)
    ops.reduce_splits(part[0, A:], grad[A:], P - A, npart, split_stride=P)
    # the non-pipelined kernel (no xtile: re-normalises the observations itself) must give the same weight gradients
    part_b = torch.full((148, P), float('nan'), device=DEV)
    ops.tc_mlp_bwd(sl(obs), epm, N, D, nm, ns, wpack, UNITS, M, A, act, dhead, delta2, delta1, part_b, P, offs)
    grad_b = torch.zeros(P, device=DEV)
    ops.reduce_splits(part_b[0, A:], grad_b[A:], P - A, npart, split_stride=P)
    torch.cuda.synchronize()
    assert torch.isfinite(grad).all()
    torch.testing.assert_close(grad, grad_b, rtol=1e-4, atol=1e-6)
    if True:        # wide observations too: l1_wgrad_tc_kernel TMA-loads the tiles l1_fwd_tc_kernel emitted
        part_c = torch.full((148, P), float('nan'), device=DEV)
        ops.tc_mlp_bwd(sl(obs), epm, N, D, nm, ns, wpack, UNITS, M, A, act, dhead, delta2, delta1, part_c, P, offs, xtile=xt)
        grad_c = torch.zeros(P, device=DEV)
        ops.reduce_splits(part_c[0, A:], grad_c[A:], P - A, npart, split_stride=P)
        torch.cuda.synchronize()
        assert torch.equal(grad_c, grad_b)        # same single-buffered kernel, same bf16 X tile bytes -> bit-identical
    d2 = decode_tiles(delta2, n_tiles, UNITS[1])[:M]; d1 = decode_tiles(delta1, n_tiles, UNITS[0])[:M]
    assert rel_l2(d2, dA[1]) < 8e-2 and cosine(d2, dA[1]) > 0.997, (rel_l2(d2, dA[1]), cosine(d2, dA[1]))
    assert rel_l2(d1, dA[0]) < 8e-2 and cosine(d1, dA[0]) > 0.997, (rel_l2(d1, dA[0]), cosine(d1, dA[0]))
    refs = {'W0': gW[0], 'b0': gb[0], 'W1': gW[1], 'b1': gb[1], 'W2': gW[2], 'b2': gb[2], 'W_head': gWh, 'b_head': gbh}
    for k, r in refs.items():
        got = grad[offs[k]:offs[k] + r.numel()].view_as(r)
        assert rel_l2(got, r) < 5e-2 and cosine(got, r) > 0.998, (k, rel_l2(got, r), cosine(got, r))


def test_tc_rollout_vs_fp32():
    _rollout_case(D, 1000)


@pytest.mark.parametrize('Dw,N', [(256, 1000), (105, 128), (72, 20000)])
def test_tc_wide_rollout_vs_fp32(Dw, N):
    _rollout_case(Dw, N)


def _rollout_case(D, N):
    from rl_games_b200 import ops
    wide = D > 64
    g = torch.Generator().manual_seed(3)
    W, b, Wh, bh, logstd = make_net(g, D)
    scratch = torch.zeros((N + 127) // 128 * ops.tc_tile_bytes(D, UNITS, A)[0], dtype=torch.uint8, device=DEV) if wide else None
    obs = (torch.randn(N, D, generator=g) * 2).to(DEV)
    nm = (torch.randn(D, generator=g) * 0.3).to(DEV); ns = (torch.rand(D, generator=g) + 0.7).to(DEV)
    noise = torch.randn(N, A, generator=g).to(DEV)
    vm, vv = torch.tensor([1.5], dtype=torch.float64, device=DEV), torch.tensor([4.0], dtype=torch.float64, device=DEV)
    lo, hi = torch.full((A,), -1.0, device=DEV), torch.full((A,), 1.0, device=DEV)
    dones_cur = (torch.rand(N, generator=g) < 0.3).to(torch.uint8).to(DEV)

    def outs():
        return dict(a=torch.empty(N, A, device=DEV), m=torch.empty(N, A, device=DEV), s=torch.empty(N, A, device=DEV),
                    nl=torch.empty(N, device=DEV), v=torch.empty(N, device=DEV), e=torch.empty(N, A, device=DEV),
                    d=torch.zeros(N, dtype=torch.uint8, device=DEV))
    r, t = outs(), outs()
    ra = [torch.empty(N, u, device=DEV) for u in UNITS]
    ops.linear_fwd(obs, W[0], b[0], ra[0], 1, norm_mean=nm, norm_std=ns)
    ops.linear_fwd(ra[0], W[1], b[1], ra[1], 1); ops.linear_fwd(ra[1], W[2], b[2], ra[2], 1)
    ops.policy_head_sample(ra[2], Wh, bh, logstd, vm, vv, True, noise, 1, None, 0, r['a'], r['m'], r['s'], r['nl'], r['v'], r['e'], True,
                           lo, hi, dones_cur, r['d'], None, None, N, A)
    wpack = torch.zeros(ops.tc_pack_bytes(D, UNITS, A), dtype=torch.uint8, device=DEV)
    ops.tc_pack_weights(W[0], W[1], W[2], Wh, D, UNITS, A, wpack)
    ops.tc_mlp_fwd_rollout(obs, D, nm, ns, wpack, b, bh, logstd, UNITS, N, A, vm, vv, True, noise, 1, None, 0, t['a'], t['m'], t['s'],
                           t['nl'], t['v'], t['e'], True, lo, hi, dones_cur, t['d'], None, None, l1_scratch=scratch)
    torch.cuda.synchronize()
    torch.testing.assert_close(t['m'], r['m'], rtol=0, atol=4e-2)
    torch.testing.assert_close(t['a'], r['a'], rtol=0, atol=4e-2)
    torch.testing.assert_close(t['v'], r['v'], rtol=0, atol=8e-2)
    assert torch.equal(t['s'], r['s']) and torch.equal(t['d'], r['d'])
    # neglogp is evaluated at the sampled action: z = eps exactly in both paths
    torch.testing.assert_close(t['nl'], r['nl'], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(t['e'], torch.clamp(t['a'], -1, 1), rtol=0, atol=1e-6)
    # values_only
    v2 = torch.empty(N, device=DEV)
    ops.tc_mlp_fwd_rollout(obs, D, nm, ns, wpack, b, bh, logstd, UNITS, N, A, vm, vv, True, None, 0, None, 0, None, None, None, None, v2,
                           None, False, None, None, None, None, None, None, values_only=True, l1_scratch=scratch)
    assert torch.equal(v2, t['v'])
    if wide:        # the scratch holds the layer-1 activations of all N rows
        a1 = decode_tiles(scratch, (N + 127) // 128, UNITS[0])[:N]
        assert rel_l2(a1, ra[0]) < 1.5e-2, rel_l2(a1, ra[0])
