"""Torch (CPU) stand-ins for the C-ABI ops used by ``rl_games_b200.agent_discrete`` -- TEST INFRASTRUCTURE ONLY.

They restate, tensor for tensor, what each kernel does (argument meaning, arena addressing through (rows_per_chunk, chunk_stride),
split-partial layout, in-place state updates), so that the HOST logic of the agent (arena views, flat-parameter offsets, call order,
scheduler, checkpoints) can be exercised against the reference's golden runs without a GPU.  They are installed by monkeypatching
``rl_games_b200.ops`` inside one test; nothing in the product imports this module, and it proves nothing about the CUDA kernels
themselves (those are checked on a GPU against the same oracle).
"""
import torch

from oracle import ppo_discrete_oracle as DO
from oracle import ppo_oracle as O

ACT = {0: lambda x: x, 1: torch.nn.functional.elu, 2: torch.relu, 3: torch.tanh}


def _act_grad_from_out(a, act):
    if act == 1:
        return torch.where(a > 0, torch.ones_like(a), a + 1.0)
    if act == 2:
        return (a > 0).float()
    if act == 3:
        return 1.0 - a * a
    return torch.ones_like(a)


def _flat(t):
    """1-D view of the storage from t's first element to the end of the storage (what a raw device pointer addresses)"""
    n = t.untyped_storage().nbytes() // t.element_size() - t.storage_offset()
    return torch.as_strided(t, (n,), (1,))


def _rows(X, M, K, rows_per_chunk, chunk_stride, x_ld):
    m = torch.arange(M)
    if rows_per_chunk is None:
        r = m
    else:
        c = m // rows_per_chunk
        r = c * chunk_stride + (m - c * rows_per_chunk)
    idx = r.unsqueeze(1) * x_ld + torch.arange(K).unsqueeze(0)
    return _flat(X)[idx]


def _norm(x, mean, std):
    return x if mean is None else torch.clamp((x - mean) / std, -5.0, 5.0)


def linear_fwd(X, W, b, Y, act, rows_per_chunk=None, chunk_stride=0, x_ld=None, norm_mean=None, norm_std=None, M=None, accumulate=False):
    Nout, K = W.shape
    M = Y.shape[0] if M is None else M
    x = _norm(_rows(X, M, K, rows_per_chunk, chunk_stride, K if x_ld is None else x_ld), norm_mean, norm_std)
    y = x @ W.t() + b
    if accumulate:
        y = y + Y[:M]
    Y[:M] = ACT[act](y)


def linear_bwd_data(dY, W, A_prev, dX, act_prev, M=None):
    M = dY.shape[0] if M is None else M
    g = dY[:M] @ W
    dX[:M] = g if A_prev is None else g * _act_grad_from_out(A_prev[:M], act_prev)


def linear_bwd_weight(dY, X, dW_part, db_part, K, Nout, n_splits, rows_per_chunk=None, chunk_stride=0, x_ld=None, norm_mean=None,
                      norm_std=None, M=None, split_stride=None):
    M = dY.shape[0] if M is None else M
    x = _norm(_rows(X, M, K, rows_per_chunk, chunk_stride, K if x_ld is None else x_ld), norm_mean, norm_std)
    dW, db = dY[:M].t() @ x, dY[:M].sum(0)
    fw, fb = _flat(dW_part), _flat(db_part)
    for s in range(n_splits):          # the kernel spreads the rows over the splits; the sum is what matters: all of it in split 0
        fw[s * split_stride:s * split_stride + Nout * K] = dW.reshape(-1) if s == 0 else 0.0
        fb[s * split_stride:s * split_stride + Nout] = db if s == 0 else 0.0


def reduce_splits(part, out, n, n_splits, split_stride=None):
    stride = n if split_stride is None else split_stride
    f = _flat(part)
    out[:n] = sum(f[s * stride:s * stride + n] for s in range(n_splits))


def refresh_norm(mean, var, mean_f32, std_f32, eps=1e-5):
    mean_f32.copy_(mean.float())
    std_f32.copy_(torch.sqrt(var.float() + eps))


def moments_update(x, D, rows_per_chunk, n_chunks, chunk_stride, mean, var, count, mean_f32, std_f32, scratch, counter, eps=1e-5):
    rows = _rows(x, rows_per_chunk * n_chunks, D, rows_per_chunk, chunk_stride, D).double()
    n = rows.shape[0]
    bm, bv = rows.mean(0), rows.var(0, unbiased=False)
    tot = count.double() + n
    delta = bm - mean
    m2 = var * count.double() + bv * n + delta ** 2 * count.double() * n / tot
    mean.add_(delta * n / tot)
    var.copy_(m2 / tot)
    count.add_(n)
    refresh_norm(mean, var, mean_f32, std_f32, eps)


def mask_inv_counts(mask, H, N, envs_per_mb, inv_count):
    for i in range(N // envs_per_mb):
        inv_count[i] = 1.0 / max(float(mask[:, i * envs_per_mb:(i + 1) * envs_per_mb].sum()), 1.0)


def gae_fused(rewards, values, dones_u8, last_values, last_dones_u8, mask, advs, returns, partials, gamma, tau):
    a = O.gae(rewards.unsqueeze(2), values.unsqueeze(2), dones_u8.float(), last_values.unsqueeze(1), last_dones_u8.float(), gamma, tau).squeeze(2)
    advs.copy_(a)
    returns.copy_(a + values)
    w = torch.ones_like(values) if mask is None else (mask != 0).float()
    da = (returns - values).double()
    v, r, wd = values.double(), returns.double(), w.double()
    partials.zero_()
    partials[0, :7] = torch.stack([wd.sum(), (wd * v).sum(), (wd * v * v).sum(), (wd * r).sum(), (wd * r * r).sum(), (wd * da).sum(),
                                   (wd * da * da).sum()])
    return 1


def batch_moments(values, returns, mask, partials):
    """the moment partials of gae_fused alone (prepare_dataset after a caller edited values / returns)"""
    w = torch.ones_like(values) if mask is None else (mask != 0).float()
    da = (returns - values).double()
    v, r, wd = values.double(), returns.double(), w.double()
    # masked entries contribute exact zeros whatever they hold (the kernel predicates; a product with a zero weight would turn inf into nan)
    z = lambda t: torch.where(wd != 0, t, torch.zeros_like(t))       # noqa: E731
    partials.zero_()
    partials[0, :7] = torch.stack([wd.sum(), z(v).sum(), z(v * v).sum(), z(r).sum(), z(r * r).sum(), z(da).sum(), z(da * da).sum()])
    return 1


def prepare_batch(values, returns, mask, partials, n_partials, vms_mean, vms_var, vms_count, old_values_n, returns_n, advs_n,
                  normalize_value, normalize_advantage, freeze_stats=False):
    acc = partials[:n_partials].sum(0)
    n = float(acc[0])
    mean, var, cnt = float(vms_mean[0]), float(vms_var[0]), float(vms_count[0])

    def merge(mean, var, cnt, bm, bv, bn):
        tot = cnt + bn
        d = bm - mean
        return mean + d * bn / tot, (var * cnt + bv * bn + d * d * cnt * bn / tot) / tot, tot
    mean_v, var_v = mean, var
    if normalize_value and not freeze_stats and n > 0:
        mv = float(acc[1]) / n
        mean, var, cnt = merge(mean, var, cnt, mv, max(float(acc[2]) / n - mv * mv, 0.0), n)
        mean_v, var_v = mean, var
        mr = float(acc[3]) / n
        mean, var, cnt = merge(mean, var, cnt, mr, max(float(acc[4]) / n - mr * mr, 0.0), n)
        if mask is not None:
            mean_v, var_v = mean, var
        vms_mean[0], vms_var[0] = mean, var
        vms_count[0] = int(vms_count[0]) + 2 * int(round(n))

    def nc(x, m, v):
        return torch.clamp((x - float(torch.tensor(m, dtype=torch.float32))) / torch.sqrt(torch.tensor(v, dtype=torch.float32) + 1e-5), -5.0, 5.0)
    a = returns - values
    old_values_n.copy_(nc(values, mean_v, var_v) if normalize_value else values)
    returns_n.copy_(nc(returns, mean, var) if normalize_value else returns)
    if normalize_advantage:
        if mask is None:
            m = float(acc[5]) / n
            v = max((float(acc[6]) - n * m * m) / (n - 1.0), 0.0)
        else:
            sm = max(n, 1.0)
            m = float(acc[5]) / sm
            v = max((float(acc[6]) / sm - m * m) * sm / max(sm - 1.0, 1.0), 0.0)
        a = (a - m) / (v ** 0.5 + 1e-8)
    advs_n.copy_(a)


def post_step(rewards, dones, time_outs, values_t, valid_t, rewards_out_t, dones_cur, prev_dones, ep_state, meter, games_to_track, scratch,
              counter, N, cfg):
    r = rewards.float()
    sh = torch.clamp((r + cfg.shift_value) * cfg.scale_value, cfg.min_val, cfg.max_val)
    if cfg.log_val:
        sh = torch.log(sh)
    if cfg.value_bootstrap and time_outs is not None:
        sh = sh + cfg.gamma * values_t * time_outs.float()
    rewards_out_t.copy_(sh)
    live = torch.ones(N) if valid_t is None else valid_t
    cr, cs, cl = ep_state[0] + r * live, ep_state[1] + sh * live, ep_state[2] + live
    d = dones.float()
    done = d != 0
    n = int(done.sum())
    if n > 0:          # AverageMeter.update (torch_ext.py:333-342)
        size = min(n, games_to_track)
        old = min(games_to_track - size, float(meter[3]))
        tot = old + size
        for slot, x in ((0, cr), (1, cs), (2, cl)):
            meter[slot] = (float(meter[slot]) * old + float(x[done].double().mean()) * size) / tot
        meter[3] = tot
        meter[4] += n
    nd = 1.0 - d
    ep_state[0], ep_state[1], ep_state[2] = cr * nd, cs * nd, cl * nd
    dones_cur.copy_(done.to(torch.uint8))
    if prev_dones is not None:
        prev_dones.copy_(d)


def _cat_heads(z, masks, head_sizes):
    """per-head (normalised logits, probs, entropy) of the concatenated logits; one head when head_sizes is None"""
    sizes = [z.shape[1]] if not head_sizes else list(head_sizes)
    zs = torch.split(z, sizes, dim=1)
    ms = [None] * len(sizes) if masks is None else torch.split(masks.bool(), sizes, dim=1)
    return [DO.categorical_masked(zz, mm) for zz, mm in zip(zs, ms)]


def categorical_sample(logits, ld, K, value_raw, value_ld, action_masks, u_tape, seed, rng_epoch, step_index, vms_mean, vms_var,
                       normalize_value, actions, neglogp, values, dones_cur, dones_out, prev_dones, valid_out, N, values_only=False,
                       head_sizes=None):
    val = _flat(value_raw)[torch.arange(N) * value_ld]
    if normalize_value:
        val = torch.sqrt(vms_var.float() + 1e-5) * torch.clamp(val, -5.0, 5.0) + vms_mean.float()
    values.copy_(val)
    if values_only:
        return
    z = _flat(logits)[(torch.arange(N) * ld).unsqueeze(1) + torch.arange(K).unsqueeze(0)]
    heads = _cat_heads(z, action_masks, head_sizes)
    nh = len(heads)
    u = torch.rand(nh, N) if u_tape is None else u_tape.reshape(nh, N)      # the kernel draws Philox uniforms when no tape is given
    acts = [DO.sample_inverse_cdf(h[1], u[j]) for j, h in enumerate(heads)]
    actions.copy_(torch.stack(acts, dim=-1).reshape(actions.shape))
    neglogp.copy_(sum(-h[0].gather(1, a.unsqueeze(1)).squeeze(1) for h, a in zip(heads, acts)))
    if dones_out is not None:
        dones_out.copy_(dones_cur)
    if valid_out is not None:
        valid_out.copy_(1.0 - prev_dones if prev_dones is not None else torch.ones(N))


def categorical_loss(logits, ld, K, values, value_ld, actions, action_masks, old_values_n, returns_n, old_neglogp, advs_n, mask,
                     rows_per_chunk, chunk_stride, M, cfg, inv_count, d_logits, d_ld, d_value, dv_ld, partials, head_sizes=None):
    mrow = torch.arange(M)
    z = _flat(logits)[(mrow * ld).unsqueeze(1) + torch.arange(K).unsqueeze(0)].clone().requires_grad_(True)
    v = _flat(values)[mrow * value_ld].clone().requires_grad_(True)
    nh = len(head_sizes) if head_sizes else 1

    def arena(t, width=1):
        return _rows(t, M, width, rows_per_chunk, chunk_stride, width)
    act = arena(actions, nh)
    am = None if action_masks is None else arena(action_masks, K).bool()
    rm = None if mask is None else arena(mask).squeeze(1)
    old_nlp, adv = arena(old_neglogp).squeeze(1), arena(advs_n).squeeze(1)
    heads = _cat_heads(z, am, head_sizes)
    nlp = sum(-h[0].gather(1, act[:, j:j + 1]).squeeze(1) for j, h in enumerate(heads))
    ent = sum(h[2] for h in heads)
    a = O.actor_loss(old_nlp, nlp, adv, bool(cfg.ppo), cfg.e_clip, smooth=bool(cfg.use_smooth_clamp))
    c = O.critic_loss(arena(old_values_n), v.unsqueeze(1), cfg.e_clip, arena(returns_n), bool(cfg.clip_value))
    w = torch.full((M,), 1.0 / M) if inv_count is None else rm * inv_count[0]
    la, lc, le = (a * w).sum(), (c.squeeze(1) * w).sum(), (ent * w).sum()
    (la + 0.5 * cfg.critic_coef * lc - cfg.entropy_coef * le).backward()
    _flat(d_logits)[(mrow * d_ld).unsqueeze(1) + torch.arange(K).unsqueeze(0)] = z.grad
    _flat(d_value)[mrow * dv_ld] = v.grad
    kl = (0.5 * (old_nlp - nlp.detach()) ** 2 * w).sum()
    partials.zero_()
    partials[0, :4] = torch.stack([la.detach(), lc.detach(), le.detach(), kl]).double()
    return 1


def adam_step(params, grads, exp_avg, exp_avg_sq, state_d, kl_dev, cfg, stats_out, counter, n=None, wpack=None, pack_table=None,
              merge_next=None):
    n = params.numel() if n is None else n
    g = grads[:n] * float(cfg.grad_scale)
    if cfg.truncate_grads:
        g = g * min(float(cfg.grad_norm) / (float(g.norm()) + 1e-6), 1.0)
    lr, step = float(state_d[0]), float(state_d[1]) + 1.0
    p1 = (float(state_d[2]) if float(state_d[2]) > 0 else 1.0) * cfg.beta1
    p2 = (float(state_d[3]) if float(state_d[3]) > 0 else 1.0) * cfg.beta2
    if cfg.weight_decay:
        g = g + cfg.weight_decay * params[:n]
    exp_avg[:n] = exp_avg[:n] + (g - exp_avg[:n]) * (1.0 - cfg.beta1)
    exp_avg_sq[:n] = cfg.beta2 * exp_avg_sq[:n] + (1.0 - cfg.beta2) * g * g
    denom = exp_avg_sq[:n].sqrt() / (1.0 - p2) ** 0.5 + cfg.eps
    params[:n] -= (lr / (1.0 - p1)) * exp_avg[:n] / denom
    state_d[1], state_d[2], state_d[3] = step, p1, p2


def bump_u64(p):
    p.add_(1)


def install(monkeypatch):
    """replace the ops the discrete agent calls by the stand-ins above"""
    from rl_games_b200 import ops
    for name in ('linear_fwd', 'linear_bwd_data', 'linear_bwd_weight', 'reduce_splits', 'refresh_norm', 'moments_update', 'mask_inv_counts',
                 'gae_fused', 'batch_moments', 'prepare_batch', 'post_step', 'categorical_sample', 'categorical_loss', 'adam_step', 'bump_u64'):
        monkeypatch.setattr(ops, name, globals()[name])


# ---------------------------------------------------------------------------------------------- continuous (Gaussian) heads, fp32 path
def policy_head_sample(a_last, W_head, b_head, logstd, vms_mean, vms_var, normalize_value, noise, seed, rng_epoch, step_index, actions, mus,
                       sigmas, neglogp, values, env_actions, clip_actions, act_low, act_high, dones_cur, dones_out, prev_dones, valid_out, N, A,
                       values_only=False):
    head = a_last[:N] @ W_head.t() + b_head
    val = head[:, 0]
    if normalize_value:
        val = torch.sqrt(vms_var.float() + 1e-5) * torch.clamp(val, -5.0, 5.0) + vms_mean.float()
    values.copy_(val)
    if values_only:
        return
    assert noise is not None, 'the host-logic test always supplies the normal tape'
    mu, sg = head[:, 1:], torch.exp(logstd).expand(N, A)
    act = mu + sg * noise
    actions.copy_(act); mus.copy_(mu); sigmas.copy_(sg)
    neglogp.copy_(O.neglogp_fn(act, mu, sg, logstd.expand(N, A)))
    if env_actions is not None:
        ea = act
        if clip_actions:
            ea = torch.clamp(act, -1.0, 1.0) * ((act_high - act_low) * 0.5) + (act_high + act_low) * 0.5
        env_actions.copy_(ea)
    if dones_out is not None:
        dones_out.copy_(dones_cur)
    if valid_out is not None:
        valid_out.copy_(1.0 - prev_dones if prev_dones is not None else torch.ones(N))


_LOSS_SIDE = {}


def ppo_head_loss(a_last, W_head, b_head, logstd, actions, old_mu, old_sigma, old_values_n, returns_n, old_neglogp, advs_n, mask, rows_per_chunk,
                  chunk_stride, M, A, cfg, inv_count, d_head, d_alast, act_last, partials, mu_out=None, value_out=None, neglogp_out=None):
    """heads + calc_losses + backward to (head, a_last, logstd); new mu/sigma overwrite the old ones in the arena (datasets.py:33-43).
    The loss scalars / d_logstd travel to ppo_loss_finalize through a side channel instead of the kernel's partial rows."""
    def arena(t, width=1):
        return _rows(t, M, width, rows_per_chunk, chunk_stride, width)
    al = a_last[:M].clone().requires_grad_(True)
    ls = logstd.clone().requires_grad_(True)
    Wh, bh = W_head.clone().requires_grad_(True), b_head.clone().requires_grad_(True)
    head = al @ Wh.t() + bh
    head.retain_grad()
    value, mu = head[:, :1], head[:, 1:]
    sg = torch.exp(ls).expand(M, A)
    act, omu, osg = arena(actions, A), arena(old_mu, A), arena(old_sigma, A)
    rm = None if mask is None else arena(mask).squeeze(1)
    nlp = O.neglogp_fn(act, mu, sg, ls.expand(M, A))
    ent = (0.5 + 0.5 * torch.log(torch.tensor(2 * torch.pi)) + torch.log(sg)).sum(-1)
    a = O.actor_loss(arena(old_neglogp).squeeze(1), nlp, arena(advs_n).squeeze(1), bool(cfg.ppo), cfg.e_clip, smooth=bool(cfg.use_smooth_clamp))
    c = O.critic_loss(arena(old_values_n), value, cfg.e_clip, arena(returns_n), bool(cfg.clip_value)).squeeze(1)
    if cfg.has_bounds_loss and cfg.bound_loss_type == 1:
        b = O.bound_loss(mu, cfg.bounds_loss_coef)
    elif cfg.has_bounds_loss and cfg.bound_loss_type == 2:
        b = O.reg_loss(mu, cfg.bounds_loss_coef)
    else:
        b = torch.zeros(M)
    w = torch.full((M,), 1.0 / M) if inv_count is None else rm * inv_count[0]
    la, lc, le, lb = (a * w).sum(), (c * w).sum(), (ent * w).sum(), (b * w).sum()
    # entropy's gradient is applied in ppo_loss_finalize (entropy_coef lives in device memory there)
    (la + 0.5 * cfg.critic_coef * lc + cfg.bounds_loss_coef * lb).backward(retain_graph=True)
    d_head[:M] = head.grad
    d_alast[:M] = al.grad * _act_grad_from_out(a_last[:M], act_last)
    dls = ls.grad.clone()
    ls.grad = None
    le.backward()
    kl = O.policy_kl(mu.detach(), sg.detach(), omu, osg, reduce=False)
    kl = (kl * w).sum()
    lr_ = arena(old_neglogp).squeeze(1) - nlp.detach()
    lo, hi = torch.log1p(torch.tensor(-cfg.e_clip)), torch.log1p(torch.tensor(cfg.e_clip))
    clipped = ((lr_ < lo) | (lr_ > hi)).float()
    mk = torch.ones(M) if rm is None else rm
    _LOSS_SIDE['last'] = dict(stats=[la.detach(), lc.detach(), le.detach(), lb.detach(), kl, mk.sum(), (mk * clipped).sum() / mk.sum().clamp(min=1.0)],
                              dls=dls, dent=ls.grad.clone())
    # write-back of the new mu / sigma
    m_ = torch.arange(M)
    cidx = m_ // rows_per_chunk
    r = cidx * chunk_stride + (m_ - cidx * rows_per_chunk)
    idx = r.unsqueeze(1) * A + torch.arange(A).unsqueeze(0)
    _flat(old_mu)[idx] = mu.detach()
    _flat(old_sigma)[idx] = sg.detach()
    return 1


def ppo_loss_finalize(partials, n_partials, A, entropy_coef_dev, stats, d_logstd, kl_out=None):
    side = _LOSS_SIDE['last']
    for i, v in enumerate(side['stats']):
        stats[i] = float(v)
    d_logstd.copy_(side['dls'] - float(entropy_coef_dev[0]) * side['dent'])
    if kl_out is not None:
        kl_out[0] = float(side['stats'][4])


def adam_step_full(params, grads, exp_avg, exp_avg_sq, state_d, kl_dev, cfg, stats_out, counter, n=None, wpack=None, pack_table=None,
                   merge_next=None):
    """adam_step + the on-device adaptive-KL schedule (schedulers.py:19-33) and the LR / grad-norm stats slots"""
    lr = float(state_d[0])
    n = params.numel() if n is None else n
    gnorm = float((grads[:n] * float(cfg.grad_scale)).norm())
    adam_step(params, grads, exp_avg, exp_avg_sq, state_d, kl_dev, cfg, stats_out, counter, n=n)
    if cfg.adaptive_lr and kl_dev is not None and _USE_HOST_SCHED['on']:
        state_d[0] = lr_schedule_step_host(lr, float(kl_dev[0]) * float(cfg.grad_scale), cfg, state_d)
    elif cfg.adaptive_lr and kl_dev is not None:
        kl = float(kl_dev[0]) * float(cfg.grad_scale)
        apply = True
        if cfg.adaptive_lr >= 2:        # schedule_type 'standard': accumulate (2) / accumulate + step on the mean + reset (3)
            s_, c_ = float(state_d[4]) + kl, float(state_d[5]) + 1.0
            apply = cfg.adaptive_lr == 3
            state_d[4], state_d[5] = (0.0, 0.0) if apply else (s_, c_)
            kl = s_ / c_
        new_lr = lr
        if apply and kl > 2.0 * cfg.kl_threshold:
            new_lr = max(lr / cfg.lr_multiplier, cfg.min_lr)
        if apply and kl < 0.5 * cfg.kl_threshold:
            new_lr = min(lr * cfg.lr_multiplier, cfg.max_lr)
        state_d[0] = new_lr
    if stats_out is not None:
        stats_out[7] = lr
        stats_out[8] = gnorm
    if merge_next is not None:          # the optimiser kernel's last CTA merges the NEXT minibatch's observation moments
        o = merge_next
        obs_stats_merge(o.mbmom, o.shift, o.D, o.n_rows, o.mean, o.var, o.count, o.mean_f32, o.std_f32, o.eps)


def rnn_train_dones(dones_u8, valid, out_u8):
    """the kernel's own element function compiled for the host (csrc/rnn.cu rnn_train_done), checked against the reference's expression"""
    H, N = dones_u8.shape
    assert dones_u8.is_contiguous() and valid.is_contiguous() and out_u8.is_contiguous()
    rc = _host_lib().b200rl_hosttest_rnn_train_dones(_p(dones_u8), _p(valid), _p(out_u8), int(H), int(N))
    assert rc == 0
    ref = dones_u8.clone()
    ref[1:] = torch.maximum(dones_u8[1:], (valid[:-1] == 0).to(torch.uint8))
    assert torch.equal(out_u8, ref)


def lr_schedule_apply(state_d, kl_dev, kl_scale, base_lr, cfg):
    """b200rl_lr_schedule_apply: the kernel forces mode 1 and calls the optimiser kernels' scheduler step -- here that same function, on the host"""
    import copy
    c = copy.copy(cfg)
    c.adaptive_lr = 1
    state_d[0] = lr_schedule_step_host(base_lr, float(kl_dev[0]) * kl_scale, c, state_d)


def adv_ema_normalize(advs, partials, n_partials, ema_state, ema_step, decay, training=True):
    acc = partials[:n_partials].sum(0)
    n = float(acc[0])
    if training and n > 0:
        ema_state[0] = ema_state[0] * decay + (1.0 - decay) * float(acc[5] / n)
        ema_state[1] = ema_state[1] * decay + (1.0 - decay) * float(acc[6] / n)
        ema_step.add_(1)
    std = torch.sqrt(torch.clamp_min(ema_state[1] - ema_state[0] ** 2, 1e-10))
    advs.copy_(torch.clamp((advs - ema_state[0]) / std, -5.0, 5.0))


def normalize(x, mean, var, denorm=False, eps=1e-5, out=None):
    std = torch.sqrt(var.float() + eps)
    y = std * torch.clamp(x, -5.0, 5.0) + mean.float() if denorm else torch.clamp((x - mean.float()) / std, -5.0, 5.0)
    if out is None:
        return y
    out.copy_(y.reshape(out.shape))
    return out


def value_loss(values, value_ld, old_values_n, returns_n, mask, rows_per_chunk, chunk_stride, M, e_clip, clip_value, inv_count, d_value, dv_ld,
               partials):
    mrow = torch.arange(M)
    v = _flat(values)[mrow * value_ld].clone().requires_grad_(True)

    def arena(t):
        return _rows(t, M, 1, rows_per_chunk, chunk_stride, 1)
    c = O.critic_loss(arena(old_values_n), v.unsqueeze(1), e_clip, arena(returns_n), bool(clip_value)).squeeze(1)
    w = torch.full((M,), 1.0 / M) if inv_count is None else arena(mask).squeeze(1) * inv_count[0]
    loss = (c * w).sum()
    loss.backward()
    _flat(d_value)[mrow * dv_ld] = v.grad
    partials.zero_()
    partials[0, 0] = loss.detach().double()
    return 1


# ---------------------------------------------------------------------------------------------- LSTM-with-dones cell ops (csrc/rnn.cu)
def _crow(S, rpc, stride):
    s_ = torch.arange(S)
    if rpc <= 0:
        return s_
    c = s_ // rpc
    return c * stride + (s_ - c * rpc)


def lstm_cell_fwd(gates, cin, c_out, h_out, S, Hd, h_scatter=None, scatter_rpc=0, scatter_stride=0, hin_next=None, cin_next=None,
                  done_next=None, done_rpc=0, done_stride=0):
    g = gates[:S]
    i, f, gg, o = torch.sigmoid(g[:, :Hd]), torch.sigmoid(g[:, Hd:2 * Hd]), torch.tanh(g[:, 2 * Hd:3 * Hd]), torch.sigmoid(g[:, 3 * Hd:])
    gates[:S] = torch.cat([i, f, gg, o], dim=1)
    c = f * cin[:S] + i * gg
    h = o * torch.tanh(c)
    c_out[:S] = c
    h_out[:S] = h
    if h_scatter is not None:
        rows = (torch.arange(S) // scatter_rpc) * scatter_stride + torch.arange(S) % scatter_rpc
        _flat(h_scatter)[rows.unsqueeze(1) * Hd + torch.arange(Hd).unsqueeze(0)] = h
    if hin_next is not None:
        m = torch.ones(S, 1)
        if done_next is not None:
            m = 1.0 - _flat(done_next)[_crow(S, done_rpc, done_stride)].float().unsqueeze(1)
        hin_next[:S] = h * m
        cin_next[:S] = c * m


def lstm_cell_bwd(gates_act, c_t, cin, dgates, dcin, S, Hd, dH=None, scatter_rpc=0, scatter_stride=0, dhin_next=None, dcin_next=None,
                  done_next=None, done_rpc=0, done_stride=0):
    g = gates_act[:S]
    i, f, gg, o = g[:, :Hd], g[:, Hd:2 * Hd], g[:, 2 * Hd:3 * Hd], g[:, 3 * Hd:]
    m = torch.ones(S, 1)
    if done_next is not None:
        m = 1.0 - _flat(done_next)[_crow(S, done_rpc, done_stride)].float().unsqueeze(1)
    dh = torch.zeros(S, Hd)
    if dH is not None:
        rows = (torch.arange(S) // scatter_rpc) * scatter_stride + torch.arange(S) % scatter_rpc
        dh = _flat(dH)[rows.unsqueeze(1) * Hd + torch.arange(Hd).unsqueeze(0)].clone()
    dc = torch.zeros(S, Hd)
    if dhin_next is not None:
        dh = dh + dhin_next[:S] * m
        dc = dc + dcin_next[:S] * m
    tc = torch.tanh(c_t[:S])
    dc = dc + dh * o * (1.0 - tc * tc)
    dgates[:S] = torch.cat([dc * gg * i * (1.0 - i), dc * cin[:S] * f * (1.0 - f), dc * i * (1.0 - gg * gg), dh * tc * o * (1.0 - o)], dim=1)
    dcin[:S] = dc * f


def rnn_mask_rows(inp, in_rpc, in_stride, out, S, Hd, done=None, done_rpc=0, done_stride=0):
    m = torch.ones(S, 1) if done is None else 1.0 - _flat(done)[_crow(S, done_rpc, done_stride)].float().unsqueeze(1)
    rows = _crow(S, in_rpc, in_stride)
    out[:S] = _flat(inp)[rows.unsqueeze(1) * Hd + torch.arange(Hd).unsqueeze(0)] * m


def install_continuous(monkeypatch):
    """stand-ins for everything rl_games_b200.agent.A2CAgent calls on its fp32 path (mixed_precision: False, no CUDA graph)"""
    from rl_games_b200 import ops
    install(monkeypatch)
    for name in ('policy_head_sample', 'ppo_head_loss', 'ppo_loss_finalize', 'adv_ema_normalize', 'normalize', 'value_loss', 'make_obs_merge', 'lr_schedule_apply', 'rnn_train_dones',
                 'obs_mb_moments', 'obs_stats_merge', 'lstm_cell_fwd', 'lstm_cell_bwd', 'rnn_mask_rows'):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, 'adam_step', adam_step_full)
    monkeypatch.setattr(ops, 'set_pdl', lambda enable: False)
    # layer-wise tensor-core GEMMs (mixed_precision: True on LSTM policies / geometries without fused kernels): same contract, fp32 here
    for name in ('linear_fwd', 'linear_bwd_data', 'linear_bwd_weight'):
        monkeypatch.setattr(ops, name + '_tc', (lambda f: (lambda *a, bf16_arena=None, **k: f(*a, **k)))(globals()[name]))
    monkeypatch.setattr(ops, 'cast_bf16', lambda src, dst: dst.copy_(src))


# ---------------------------------------------------------------------------------------------- tcgen05 path (host logic only: fp32 maths)
_TC = {}          # wpack.data_ptr() -> the fp32 weight views tc_pack_weights was given (the kernels read the packed bf16 copy instead)


class _ObsMergeRef:
    """what ops.make_obs_merge packs into a ctypes struct, kept as tensor references"""

    def __init__(self, mbmom_i, shift, D, n_rows, mean, var, count, mean_f32, std_f32, eps=1e-5):
        self.mbmom, self.shift, self.D, self.n_rows = mbmom_i, shift, D, n_rows
        self.mean, self.var, self.count, self.mean_f32, self.std_f32, self.eps = mean, var, count, mean_f32, std_f32, eps


def make_obs_merge(*a, **k):
    return _ObsMergeRef(*a, **k)


def obs_mb_moments(x, D, H, N, envs_per_mb, run_mean, mbmom, mb_shift, scratch, counters):
    """per-minibatch batch sums of the observations, shifted by the running mean at the start of the update phase"""
    mb_shift.copy_(run_mean.float())
    for i in range(N // envs_per_mb):
        rows = x[:, i * envs_per_mb:(i + 1) * envs_per_mb].reshape(-1, D).double() - mb_shift.double()
        mbmom[i, :D] = rows.sum(0)
        mbmom[i, D:] = (rows * rows).sum(0)


def obs_stats_merge(mbmom_i, mb_shift, D, n_rows, mean, var, count, mean_f32, std_f32, eps=1e-5):
    n, cnt0 = float(n_rows), float(count[0])
    ms = mbmom_i[:D] / n
    bm = mb_shift.double() + ms
    bv = torch.clamp_min(mbmom_i[D:] / n - ms * ms, 0.0)
    tot = cnt0 + n
    delta = bm - mean
    m2 = var * cnt0 + bv * n + delta * delta * cnt0 * n / tot
    mean.add_(delta * n / tot)
    var.copy_(m2 / tot)
    count.add_(int(n_rows))
    refresh_norm(mean, var, mean_f32, std_f32, eps)


def tc_pack_weights(W1, W2, W3, W_head, D, units, A, wpack):
    _TC[wpack.data_ptr()] = (W1, W2, W3, W_head)


def _tc_forward(x, ws, b, b_head, act=1):
    h = x
    for W, bb in zip(ws[:3], b):
        h = ACT[act](h @ W.t() + bb)
    return h @ ws[3].t() + b_head


def tc_mlp_fwd_rollout(obs, D, nm, ns, wpack, b, b_head, logstd, units, N, A, vms_mean, vms_var, normalize_value, noise, seed, rng_epoch,
                       step_index, actions, mus, sigmas, neglogp, values, env_actions, clip_actions, act_low, act_high, dones_cur, dones_out,
                       prev_dones, valid_out, values_only=False, l1_scratch=None, activation=1):
    ws = _TC[wpack.data_ptr()]
    if _TC.get('kind', 1) == 2:      # wide observations: the layer-1 kernel parks N rows of a1 tiles in the scratch between the two launches
        assert l1_scratch is not None and l1_scratch.numel() >= (N + 127) // 128 * units[0] * 256
    h = obs[:N].reshape(N, D)
    for W, bb in zip(ws[:3], b):
        h = ACT[activation]((_norm(h, nm, ns) if W is ws[0] else h) @ W.t() + bb)
    policy_head_sample(h, ws[3], b_head, logstd, vms_mean, vms_var, normalize_value, noise, seed, rng_epoch, step_index, actions, mus, sigmas,
                       neglogp, values, env_actions, clip_actions, act_low, act_high, dones_cur, dones_out, prev_dones, valid_out, N, A,
                       values_only=values_only)


def tc_mlp_fwd_train(obs, rows_per_chunk, chunk_stride, D, nm, ns, wpack, b, b_head, logstd, units, M, A, actions, old_mu, old_sigma, old_values_n,
                     returns_n, old_neglogp, advs_n, mask, cfg, inv_count, act, dhead, partials, xtile=None, activation=1):
    """whole training forward + loss + backward of the MLP in fp32 with autograd; the gradients wait in a side channel for tc_mlp_bwd"""
    ws = [w.clone().requires_grad_(True) for w in _TC[wpack.data_ptr()]]
    bs = [t.clone().requires_grad_(True) for t in b]
    x = _norm(_rows(obs, M, D, rows_per_chunk, chunk_stride, D), nm, ns)
    h = x
    for W, bb in zip(ws[:3], bs):
        h = ACT[activation](h @ W.t() + bb)
    hl = h
    hl.retain_grad()
    d_head = torch.zeros(M, A + 1)
    d_alast = torch.zeros(M, hl.shape[1])
    ppo_head_loss(hl.detach(), ws[3].detach(), b_head, logstd, actions, old_mu, old_sigma, old_values_n, returns_n, old_neglogp, advs_n, mask,
                  rows_per_chunk, chunk_stride, M, A, cfg, inv_count, d_head, d_alast, 0, partials)
    # backward of the trunk from d(loss)/d(a_last) (act_last = 0 above: d_alast is the gradient at the activation OUTPUT)
    hl.backward(d_alast)
    _LOSS_SIDE['tc_grads'] = {'W0': ws[0].grad, 'b0': bs[0].grad, 'W1': ws[1].grad, 'b1': bs[1].grad, 'W2': ws[2].grad, 'b2': bs[2].grad,
                              'W_head': d_head.t() @ hl.detach(), 'b_head': d_head.sum(0)}
    return 1


def tc_mlp_bwd(obs, rows_per_chunk, chunk_stride, D, nm, ns, wpack, units, M, A, act, dhead, delta2, delta1, part, P, offs, xtile=None, activation=1,
               pipelined_wgrad=False):
    """split partial rows: everything in row 0, zeros elsewhere; P is the row stride"""
    g = _LOSS_SIDE['tc_grads']
    f = _flat(part)
    n_rows = 3
    for s_ in range(n_rows):
        for k, v in g.items():
            f[s_ * P + offs[k]:s_ * P + offs[k] + v.numel()] = v.reshape(-1) if s_ == 0 else 0.0
    return n_rows


def reduce_adam(part, n_splits, split_stride, loss_partials, n_loss_partials, A, entropy_coef_dev, stats, kl_out, grads, params, exp_avg,
                exp_avg_sq, n, state_d, cfg, counter, nrm_part, grid_bar, wpack=None, pack_table=None, merge_next=None):
    f = _flat(part)
    grads[A:n] = sum(f[s_ * split_stride + A:s_ * split_stride + n] for s_ in range(n_splits))
    ppo_loss_finalize(loss_partials, n_loss_partials, A, entropy_coef_dev, stats, grads[:A], kl_out)
    adam_step_full(params, grads, exp_avg, exp_avg_sq, state_d, kl_out, cfg, stats, counter, n=n, merge_next=merge_next)


def install_tc(monkeypatch, kind=1):
    """stand-ins for the bf16 tcgen05 path of A2CAgent (mixed_precision: True), computed in fp32; kind 2 = the wide-observation
    edition (layer 1 in kernels of its own: only the host-visible contract differs -- scratch buffer, 128 x 256 X tiles)"""
    from rl_games_b200 import ops
    install_continuous(monkeypatch)
    _TC['kind'] = kind
    for name in ('tc_pack_weights', 'tc_mlp_fwd_rollout', 'tc_mlp_fwd_train', 'tc_mlp_bwd', 'reduce_adam'):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, 'tc_kind', lambda D, units, A: kind if len(units) == 3 else 0)
    monkeypatch.setattr(ops, 'tc_supported', lambda D, units, A, allow_wide=False: len(units) == 3 and (kind == 1 or allow_wide))
    monkeypatch.setattr(ops, 'tc_tile_bytes', lambda D, units, A: [units[0] * 256, units[1] * 256, units[2] * 256, 16 * 256])
    monkeypatch.setattr(ops, 'tc_pack_bytes', lambda D, units, A: 1024)
    monkeypatch.setattr(ops, 'tc_xtile_bytes', lambda D, units, A: 64 * 256)
    monkeypatch.setattr(ops, 'tc_pack_table', lambda D, units, A, offs: object())


# ---------------------------------------------------------------------------------------------- kernels' own thread bodies on the host
# The per-thread bodies of the discrete / critic kernels are __host__ __device__; the library exports entry points that run them over
# HOST arrays laid out like the device arena (csrc/discrete.cu, csrc/critic.cu: b200rl_hosttest_*_arena).  Swapping them in for the torch
# stand-ins runs the agents' golden tests on the kernels' REAL indexing and arithmetic -- everything except launch geometry and the
# block reduction.
def _host_lib():
    try:
        from tests import _hooks
    except ImportError:
        import _hooks
    return _hooks.load()


def _p(t):
    import ctypes
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def categorical_sample_hostkernel(logits, ld, K, value_raw, value_ld, action_masks, u_tape, seed, rng_epoch, step_index, vms_mean, vms_var,
                                  normalize_value, actions, neglogp, values, dones_cur, dones_out, prev_dones, valid_out, N, values_only=False,
                                  head_sizes=None):
    import ctypes
    assert values_only or u_tape is not None, 'the host body has no Philox: the tests supply the uniform tape'
    hs = None if not head_sizes else (ctypes.c_int * len(head_sizes))(*[int(k) for k in head_sizes])
    rc = _host_lib().b200rl_hosttest_categorical_sample_arena(
        _p(logits), int(ld), int(K), len(head_sizes) if head_sizes else 0, hs, _p(value_raw), int(value_ld), _p(action_masks), _p(u_tape),
        _p(vms_mean), _p(vms_var), int(normalize_value), _p(actions), _p(neglogp), _p(values), _p(dones_cur), _p(dones_out), _p(prev_dones),
        _p(valid_out), int(N), int(values_only))
    assert rc == 0, rc


def categorical_loss_hostkernel(logits, ld, K, values, value_ld, actions, action_masks, old_values_n, returns_n, old_neglogp, advs_n, mask,
                                rows_per_chunk, chunk_stride, M, cfg, inv_count, d_logits, d_ld, d_value, dv_ld, partials, head_sizes=None):
    import ctypes
    hs = None if not head_sizes else (ctypes.c_int * len(head_sizes))(*[int(k) for k in head_sizes])
    partials.zero_()
    rc = _host_lib().b200rl_hosttest_categorical_loss_arena(
        _p(logits), int(ld), int(K), len(head_sizes) if head_sizes else 0, hs, _p(values), int(value_ld), _p(actions), _p(action_masks),
        _p(old_values_n), _p(returns_n), _p(old_neglogp), _p(advs_n), _p(mask), int(rows_per_chunk), ctypes.c_int64(int(chunk_stride)), int(M),
        ctypes.c_void_p(ctypes.addressof(cfg)), _p(inv_count), _p(d_logits), int(d_ld), _p(d_value), int(dv_ld), _p(partials))
    assert rc == 0, rc
    return 1


def value_loss_hostkernel(values, value_ld, old_values_n, returns_n, mask, rows_per_chunk, chunk_stride, M, e_clip, clip_value, inv_count, d_value,
                          dv_ld, partials):
    import ctypes
    partials.zero_()
    rc = _host_lib().b200rl_hosttest_value_loss_arena(
        _p(values), int(value_ld), _p(old_values_n), _p(returns_n), _p(mask), int(rows_per_chunk), ctypes.c_int64(int(chunk_stride)), int(M),
        ctypes.c_float(float(e_clip)), int(clip_value), _p(inv_count), _p(d_value), int(dv_ld), _p(partials))
    assert rc == 0, rc
    return 1


def lr_schedule_step_host(lr, kl, cfg, state_d):
    """the optimiser kernels' scheduler step (csrc/adam.cu lr_schedule_step, compiled for the host); state_d: fp64 CPU tensor (>= 6 entries
    for the per-mini-epoch modes)"""
    import ctypes
    fn = _host_lib().b200rl_hosttest_lr_schedule_step
    fn.restype = ctypes.c_double
    return float(fn(ctypes.c_double(float(lr)), ctypes.c_double(float(kl)), ctypes.c_void_p(ctypes.addressof(cfg)), _p(state_d)))


_USE_HOST_SCHED = {'on': False}


def install_host_kernels(monkeypatch):
    """on top of install*(): ops whose kernels have host-compilable thread bodies run THOSE (compiled for the host) instead of the torch
    stand-ins -- the categorical sample / loss kernels, the central-value loss kernel, the scheduler step of the optimiser kernels"""
    from rl_games_b200 import ops
    monkeypatch.setattr(ops, 'categorical_sample', categorical_sample_hostkernel)
    monkeypatch.setattr(ops, 'categorical_loss', categorical_loss_hostkernel)
    monkeypatch.setattr(ops, 'value_loss', value_loss_hostkernel)
    monkeypatch.setitem(_USE_HOST_SCHED, 'on', True)
