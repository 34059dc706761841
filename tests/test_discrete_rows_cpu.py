"""The per-row arithmetic of the categorical kernels (csrc/discrete.cu: cat_sample_row / cat_loss_row are __host__ __device__), run ON
THE CPU through the library's host test entry points and checked against the oracle (sampling rule, neglogp) and against autograd
(loss pieces and gradients).  No GPU is involved: this validates the arithmetic the kernels execute, not their launch/indexing code."""
import ctypes

import pytest
import torch

from oracle import ppo_discrete_oracle as DO
from oracle import ppo_oracle as O
from tests import _hooks


class CatLossCfg(ctypes.Structure):
    _fields_ = [('e_clip', ctypes.c_float), ('critic_coef', ctypes.c_float), ('entropy_coef', ctypes.c_float),
                ('clip_value', ctypes.c_int), ('use_smooth_clamp', ctypes.c_int), ('ppo', ctypes.c_int)]


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


@pytest.fixture(scope='module')
def lib():
    cdll = _hooks.load()
    cdll.b200rl_hosttest_categorical_sample_rows.restype = ctypes.c_int
    cdll.b200rl_hosttest_categorical_sample_rows.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                             ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    cdll.b200rl_hosttest_categorical_loss_rows.restype = ctypes.c_int
    cdll.b200rl_hosttest_categorical_loss_rows.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p] + \
        [ctypes.c_void_p] * 8 + [ctypes.c_int] + [ctypes.c_void_p] * 4
    return cdll


@pytest.mark.parametrize('masked', [False, True])
def test_sample_rows_match_oracle(lib, masked):
    g = torch.Generator().manual_seed(3)
    N, K = 5000, 7
    logits = (torch.randn(N, K, generator=g) * 2).contiguous()
    masks = None
    if masked:
        masks = torch.rand(N, K, generator=g) < 0.6
        masks[torch.arange(N), torch.randint(0, K, (N,), generator=g)] = True
    u = torch.rand(N, generator=g)
    u[:8] = 0.0                                  # u == 0 must pick the first LEGAL action
    nl, probs, _ = DO.categorical_masked(logits, masks)
    a_ref = DO.sample_inverse_cdf(probs, u)
    m8 = None if masks is None else masks.to(torch.uint8).contiguous()
    actions = torch.zeros(N, dtype=torch.int64); nlp = torch.zeros(N)
    assert lib.b200rl_hosttest_categorical_sample_rows(_p(logits), K, 0, None, _p(m8), _p(u), N, _p(actions), _p(nlp)) == 0
    same = actions == a_ref
    assert same.float().mean() > 0.999           # only cdf-boundary ties (fp32 summation order) may differ
    if masks is not None:
        assert masks.gather(1, actions.unsqueeze(1)).all()
    torch.testing.assert_close(nlp[same], -nl.gather(1, a_ref.unsqueeze(1)).squeeze(1)[same], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('masked,smooth,clip_value,ppo', [(False, False, True, True), (True, False, True, True), (True, True, False, True),
                                                         (False, False, True, False)])
def test_loss_rows_match_autograd(lib, masked, smooth, clip_value, ppo):
    g = torch.Generator().manual_seed(11)
    M, K, e_clip, critic_coef, ent_coef = 700, 6, 0.2, 1.5, 0.02
    logits = (torch.randn(M, K, generator=g) * 1.5).requires_grad_(True)
    value = torch.randn(M, generator=g).requires_grad_(True)
    amask = None
    if masked:
        amask = torch.rand(M, K, generator=g) < 0.7
        amask[torch.arange(M), torch.randint(0, K, (M,), generator=g)] = True
    nl, probs, ent = DO.categorical_masked(logits, amask)
    actions = DO.sample_inverse_cdf(probs.detach(), torch.rand(M, generator=g))
    old_nlp = -nl.detach().gather(1, actions.unsqueeze(1)).squeeze(1) + torch.randn(M, generator=g) * 0.2
    adv, old_v, ret = torch.randn(M, generator=g), torch.randn(M, generator=g), torch.randn(M, generator=g)
    rmask = (torch.rand(M, generator=g) < 0.8).float() if masked else None
    w = (torch.full((M,), 1.0 / M) if rmask is None else rmask / rmask.sum()).contiguous()
    nlp = -nl.gather(1, actions.unsqueeze(1)).squeeze(1)
    a = O.actor_loss(old_nlp, nlp, adv, ppo, e_clip, smooth=smooth)
    c = O.critic_loss(old_v.unsqueeze(1), value.unsqueeze(1), e_clip, ret.unsqueeze(1), clip_value).squeeze(1)
    la, lc, le = (a * w).sum(), (c * w).sum(), (ent * w).sum()
    (la + 0.5 * critic_coef * lc - ent_coef * le).backward()
    kl = (0.5 * (old_nlp - nlp.detach()) ** 2 * w).sum()
    cfg = CatLossCfg(e_clip, critic_coef, ent_coef, int(clip_value), int(smooth), int(ppo))
    dl, dv = torch.zeros(M, K), torch.zeros(M)
    sums = torch.zeros(4, dtype=torch.float64)
    z = logits.detach().contiguous()
    m8 = None if amask is None else amask.to(torch.uint8).contiguous()
    rc = lib.b200rl_hosttest_categorical_loss_rows(_p(z), K, 0, None, _p(value.detach().contiguous()), _p(actions), _p(m8), _p(old_v), _p(ret),
                                                   _p(old_nlp.contiguous()), _p(adv), _p(w), M, ctypes.byref(cfg), _p(dl), _p(dv), _p(sums))
    assert rc == 0
    torch.testing.assert_close(sums.float(), torch.stack([la, lc, le, kl]).detach(), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(dl, logits.grad, rtol=2e-4, atol=1e-8)
    torch.testing.assert_close(dv, value.grad, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize('masked', [False, True])
def test_multi_head_rows_match_oracle_and_autograd(lib, masked):
    """multi-discrete (Tuple) space: heads [3, 4, 2] concatenated along the row; one categorical / uniform per head, neglogp and entropy
    summed over heads (models.py:128-206)"""
    g = torch.Generator().manual_seed(21)
    sizes, M = [3, 4, 2], 900
    K, nh = sum(sizes), len(sizes)
    tab = (ctypes.c_int * nh)(*sizes)
    logits = (torch.randn(M, K, generator=g) * 1.5).requires_grad_(True)
    value = torch.randn(M, generator=g).requires_grad_(True)
    amask = None
    if masked:
        amask = torch.rand(M, K, generator=g) < 0.6
        off = 0
        for k in sizes:
            amask[torch.arange(M), torch.randint(0, k, (M,), generator=g) + off] = True
            off += k
    zs = torch.split(logits, sizes, dim=1)
    ms = [None] * nh if amask is None else torch.split(amask, sizes, dim=1)
    heads = [DO.categorical_masked(z, m) for z, m in zip(zs, ms)]
    u = torch.rand(nh, M, generator=g).contiguous()
    a_ref = torch.stack([DO.sample_inverse_cdf(h[1].detach(), u[j]) for j, h in enumerate(heads)], dim=-1)
    m8 = None if amask is None else amask.to(torch.uint8).contiguous()
    z = logits.detach().contiguous()
    actions = torch.zeros(M, nh, dtype=torch.int64); nlp_k = torch.zeros(M)
    assert lib.b200rl_hosttest_categorical_sample_rows(_p(z), K, nh, tab, _p(m8), _p(u), M, _p(actions), _p(nlp_k)) == 0
    same = (actions == a_ref).all(dim=1)
    assert same.float().mean() > 0.998
    nlp = sum(-h[0].gather(1, a_ref[:, j:j + 1]).squeeze(1) for j, h in enumerate(heads))
    torch.testing.assert_close(nlp_k[same], nlp.detach()[same], rtol=1e-5, atol=1e-5)
    # loss + gradients
    ent = sum(h[2] for h in heads)
    old_nlp = nlp.detach() + torch.randn(M, generator=g) * 0.2
    adv, old_v, ret = torch.randn(M, generator=g), torch.randn(M, generator=g), torch.randn(M, generator=g)
    w = torch.full((M,), 1.0 / M)
    a = O.actor_loss(old_nlp, nlp, adv, True, 0.2, smooth=False)
    c = O.critic_loss(old_v.unsqueeze(1), value.unsqueeze(1), 0.2, ret.unsqueeze(1), True).squeeze(1)
    la, lc, le = (a * w).sum(), (c * w).sum(), (ent * w).sum()
    (la + 0.5 * 1.5 * lc - 0.02 * le).backward()
    cfg = CatLossCfg(0.2, 1.5, 0.02, 1, 0, 1)
    dl, dv, sums = torch.zeros(M, K), torch.zeros(M), torch.zeros(4, dtype=torch.float64)
    rc = lib.b200rl_hosttest_categorical_loss_rows(_p(z), K, nh, tab, _p(value.detach().contiguous()), _p(a_ref.contiguous()), _p(m8), _p(old_v),
                                                   _p(ret), _p(old_nlp.contiguous()), _p(adv), _p(w), M, ctypes.byref(cfg), _p(dl), _p(dv), _p(sums))
    assert rc == 0
    kl = (0.5 * (old_nlp - nlp.detach()) ** 2 * w).sum()
    torch.testing.assert_close(sums.float(), torch.stack([la, lc, le, kl]).detach(), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(dl, logits.grad, rtol=2e-4, atol=1e-8)
    torch.testing.assert_close(dv, value.grad, rtol=1e-5, atol=1e-9)
