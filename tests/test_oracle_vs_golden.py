"""Pin the CPU oracle (oracle/ppo_oracle.py) to outputs of the REAL reference.

The fixtures under tests/golden/*.pt were produced by tests/golden/gen_golden.py, which imports
Denys88/rl_games from /root/reference in the build container.  Everything here runs on CPU.
"""
import os

import pytest
import torch

from oracle import ppo_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def test_gae_matches_reference_pytorch_loop_bitexact():
    for c in load('gae.pt'):
        out = O.gae(*c['inputs'], c['gamma'], c['tau'])
        assert torch.equal(out, c['pytorch_gae']), c['shape']
        if c['scalar_f64_as_f32'] is not None:
            # reference's own KAT tolerance: tests/test_triton_gae.py:61
            assert torch.allclose(out, c['scalar_f64_as_f32'], atol=1e-5)
            f64 = O.gae_f64_scalar(*c['inputs'], c['gamma'], c['tau'])
            assert torch.allclose(f64.float(), c['scalar_f64_as_f32'], atol=1e-6)


def test_running_mean_std_bitexact():
    g = load('math.pt')['rms']
    rms = O.RunningMeanStd((6,))
    rms.train()
    for x, y in zip(g['xs'], g['ys']):
        assert torch.equal(rms(x), y)
    rms.eval()
    assert torch.equal(rms.running_mean, g['state']['running_mean'])
    assert torch.equal(rms.running_var, g['state']['running_var'])
    assert int(rms.count) == int(g['state']['count'])
    assert torch.equal(rms(g['x_eval']), g['y_eval'])
    assert torch.equal(rms(g['x_eval'], denorm=True), g['y_denorm'])


def test_running_mean_std_valid_rows():
    g = load('math.pt')['rms_valid_rows']
    rms = O.RunningMeanStd((1,))
    rms.train()
    rms(g['x'][g['mask']])
    assert torch.equal(rms.running_mean, g['state']['running_mean'])
    assert torch.equal(rms.running_var, g['state']['running_var'])
    assert int(rms.count) == int(g['state']['count'])


def test_masked_moments_and_normalisation():
    m = load('math.pt')
    g = m['masked']
    mean, var = O.get_mean_var_with_masks(g['v'], g['m'])
    assert torch.equal(mean, g['mean']) and torch.equal(var, g['var'])
    assert torch.equal(O.normalization_with_masks(g['v'], g['m']), g['norm_masked'])
    assert torch.equal(O.normalization_with_masks(g['v'], None), g['norm_unmasked'])
    for k in ('masked_zero', 'masked_one'):
        mean, var = O.get_mean_var_with_masks(g['v'], m[k]['m'])
        assert torch.equal(mean, m[k]['mean']) and torch.equal(var, m[k]['var'])
        assert torch.isfinite(O.normalization_with_masks(g['v'], m[k]['m'])).all()


def test_losses_bitexact():
    m = load('math.pt')
    a = m['actor']
    assert torch.equal(O.actor_loss(a['old'], a['new'], a['adv'], True, 0.2, smooth=False), a['hard'])
    assert torch.equal(O.actor_loss(a['old'], a['new'], a['adv'], True, 0.2, smooth=True), a['smooth'])
    c = m['critic']
    assert torch.equal(O.critic_loss(c['old_values'], c['values'], 0.2, c['returns'], True), c['clip'])
    assert torch.equal(O.critic_loss(c['old_values'], c['values'], 0.2, c['returns'], False), c['noclip'])
    k = m['kl']
    assert torch.equal(O.policy_kl(k['mu0'], k['s0'], k['mu1'], k['s1']), k['kl'])
    assert torch.equal(O.policy_kl(k['mu0'], k['s0'], k['mu1'], k['s1'], False), k['kl_rows'])
    d = m['diag']
    assert torch.equal(O.explained_variance(d['y_pred'], d['y']), d['ev'])
    assert torch.equal(O.policy_clip_fraction(a['new'], a['old'], 0.2), d['clip_frac'])


def test_average_meter_and_scheduler():
    m = load('math.pt')
    am = O.AverageMeter(1, 10)
    for s, (mean, size) in zip(m['meter']['seq'], m['meter']['means']):
        am.update(s)
        assert torch.equal(am.mean, mean) and am.current_size == size
    sch = O.AdaptiveScheduler(0.008)
    lr = 3e-4
    for k, ref in zip(m['adaptive']['kls'], m['adaptive']['lrs']):
        lr, _ = sch.update(lr, 0.0, 0, 0, k)
        assert lr == ref


def _oracle_from_golden(g):
    cfgk = g['config']
    cfg = {k: cfgk[k] for k in ('gamma', 'tau', 'e_clip', 'clip_value', 'critic_coef', 'entropy_coef',
                                'bound_loss_type', 'use_smooth_clamp', 'truncate_grads', 'grad_norm',
                                'learning_rate', 'kl_threshold', 'normalize_input', 'normalize_value',
                                'normalize_advantage', 'value_bootstrap', 'mini_epochs', 'normalize_rms_advantage',
                                'adv_rms_momentum', 'schedule_type', 'ppo', 'clip_actions', 'max_epochs', 'max_frames', 'schedule_entropy',
                                'games_to_track') if k in cfgk}
    rs = g.get('reward_shaper') or {}
    cfg.update(reward_scale=rs.get('scale_value', 1.0), reward_shift=rs.get('shift_value', 0.0), reward_min=rs.get('min_val', -float('inf')),
               reward_max=rs.get('max_val', float('inf')), reward_log=rs.get('log_val', False))
    cfg['actions_low'], cfg['actions_high'] = g.get('act_bounds', (-1.0, 1.0))
    cfg['bounds_loss_coef'] = cfgk.get('bounds_loss_coef', None)
    cfg['lr_schedule'] = cfgk.get('lr_schedule', None)
    cfg['weight_decay'] = cfgk.get('weight_decay', 0.0)
    cfg['mask_autoreset_rows'] = g['autoreset'] == 'next_step'
    cfg['rnn_units'] = g.get('rnn_units', 0)
    cfg['rnn_before_mlp'] = g.get('rnn_before_mlp', True)
    cfg['seq_length'] = cfgk.get('seq_length', 4)
    cfg['min_sigma'] = (g.get('space_over') or {}).get('min_sigma', 0.0)
    env = O.TapeEnv(g['obs_tape'], g['done_tape'], g['timeout_tape'])
    params = {k: v for k, v in g['init_state'].items() if k.startswith('a2c_network')}
    ag = O.OracleAgent(env, params, g['D'], g['A'], g['units'], g['N'], g['H'], g['mb'], cfg)
    ag.obs = ag.env_reset()
    return ag


@pytest.mark.parametrize('name', ['agent_base.pt', 'agent_masked.pt', 'agent_hardclip.pt', 'agent_lstm.pt', 'agent_rmsadv.pt',
                                  'agent_lstm_after.pt', 'agent_sched_standard.pt', 'agent_misc.pt', 'agent_rescale.pt', 'agent_lstm_masked.pt',
                                  'agent_lstm_after_masked.pt', 'agent_trainloop.pt', 'agent_trainloop_adaptive.pt', 'agent_minsigma.pt', 'agent_separate.pt'])
def test_full_train_epochs_match_reference_agent(name):
    """Two full train_epoch()s of the reference A2CAgent vs the oracle restatement, same tapes/noise."""
    g = load(name)
    lstm = g.get('rnn_units', 0) > 0
    separate = bool((g.get('network_over') or {}).get('separate', False))
    assert g['param_order'] == O.param_names(len(g['units']), lstm=lstm, separate=separate)
    ag = _oracle_from_golden(g)
    for ep, ref in enumerate(g['epochs_out']):
        out = ag.train_epoch(g['noise'][ep])
        ds = ref['dataset']
        # rollout + GAE + prepare_dataset: same op order => tight
        assert torch.equal(ag.buf['dones'], ref['mb_dones'])
        torch.testing.assert_close(ag.buf['rewards'], ref['mb_rewards'], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(ag.buf['values'], ref['mb_values'], rtol=1e-5, atol=1e-6)
        if ds.get('rnn_masks') is not None:
            assert torch.equal(ag.dataset['rnn_masks'], ds['rnn_masks'])
        torch.testing.assert_close(ag.dataset['advantages'], ds['advantages'], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(ag.dataset['returns'], ds['returns'], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(ag.dataset['old_values'], ds['old_values'], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(torch.stack(out['kl']), ref['kls'] if ref['kls'].numel() == len(out['kl'])
                                   else torch.stack(out['kl']), rtol=1e-3, atol=1e-7)
        torch.testing.assert_close(torch.stack(out['a_loss']), ref['a_losses'], rtol=1e-3, atol=1e-6)
        torch.testing.assert_close(torch.stack(out['c_loss']), ref['c_losses'], rtol=1e-3, atol=1e-6)
        torch.testing.assert_close(torch.stack(out['entropy']), ref['entropies'], rtol=1e-5, atol=1e-6)
        assert ag.last_lr == pytest.approx(ref['last_lr'], rel=1e-12)
        for k in O.param_names(len(g['units']), lstm=lstm, separate=separate):
            torch.testing.assert_close(ag.model.p[k].detach(), ref['state'][k], rtol=1e-4, atol=2e-6, msg=lambda m: k + m)
        st = ref['state']
        if g['config'].get('normalize_input', True):
            torch.testing.assert_close(ag.model.running_mean_std.running_mean, st['running_mean_std.running_mean'], rtol=1e-9, atol=1e-9)
            torch.testing.assert_close(ag.model.running_mean_std.running_var, st['running_mean_std.running_var'], rtol=1e-9, atol=1e-9)
            assert int(ag.model.running_mean_std.count) == int(st['running_mean_std.count'])
        else:
            assert 'running_mean_std.running_mean' not in st
        if g['config'].get('normalize_value', True):
            torch.testing.assert_close(ag.model.value_mean_std.running_mean, st['value_mean_std.running_mean'], rtol=1e-6, atol=1e-7)
            torch.testing.assert_close(ag.model.value_mean_std.running_var, st['value_mean_std.running_var'], rtol=1e-6, atol=1e-7)
            assert int(ag.model.value_mean_std.count) == int(st['value_mean_std.count'])
        torch.testing.assert_close(ag.game_rewards.mean, ref['game_rewards_mean'], rtol=1e-5, atol=1e-6)
        assert ag.game_rewards.current_size == ref['game_rewards_size']
        torch.testing.assert_close(ag.game_lengths.mean, ref['game_lengths_mean'], rtol=1e-6, atol=1e-6)


def _discrete_oracle_from_golden(g):
    from oracle import ppo_discrete_oracle as DO
    cfgk = g['config']
    cfg = {k: cfgk[k] for k in ('gamma', 'tau', 'e_clip', 'clip_value', 'critic_coef', 'entropy_coef', 'truncate_grads', 'grad_norm',
                                'learning_rate', 'kl_threshold', 'normalize_input', 'normalize_value', 'normalize_advantage',
                                'mini_epochs') if k in cfgk}
    cfg['lr_schedule'] = cfgk.get('lr_schedule', None)
    cfg['value_bootstrap'] = cfgk.get('value_bootstrap', True)
    cfg['reward_scale'] = 0.1
    cfg['mask_autoreset_rows'] = g['autoreset'] == 'next_step'
    env = DO.DiscreteTapeEnv(g['obs_tape'], g['done_tape'], g['timeout_tape'], g['K'], g['mask_tape'])
    params = {k: v for k, v in g['init_state'].items() if k.startswith('a2c_network')}
    ag = DO.DiscreteOracleAgent(env, params, g['D'], g['K'], g['units'], g['N'], g['H'], g['mb'], cfg, separate=g['separate'],
                                use_action_masks=g['use_action_masks'])
    ag.obs = ag.env_reset()
    return ag


# ------------------------------------------------------------------------------------------ discrete PPO
@pytest.mark.parametrize('name', ['agent_discrete.pt', 'agent_discrete_masked.pt', 'agent_multidiscrete.pt'])
def test_discrete_train_epochs_match_reference_agent(name):
    """Two train_epoch()s of the reference DiscreteA2CAgent (configs/ppo_cartpole.yaml shape; second fixture: shared trunk,
    action masks, next_step autoreset, normalisers, adaptive LR per mini-epoch) vs oracle/ppo_discrete_oracle.py on the same
    tapes and uniform draws.  Sampled actions and masks are integer/bool work: bit-exact."""
    from oracle import ppo_discrete_oracle as DO
    g = load(name)
    multi = isinstance(g['K'], (list, tuple))     # agent_multidiscrete.pt: Tuple(Discrete(3), Discrete(4)), ModelA2CMultiDiscrete
    assert g['param_order'] == DO.discrete_param_names(len(g['units']), g['separate'], len(g['K']) if multi else None)
    ag = _discrete_oracle_from_golden(g)
    for ep, ref in enumerate(g['epochs_out']):
        out = ag.train_epoch(g['u'][ep])          # [H + 1, N], or [H + 1, n_heads, N] for a multi-discrete space
        ds = ref['dataset']
        assert torch.equal(ag.buf['actions'], ref['mb_actions'])
        assert torch.equal(ag.dataset['actions'], ds['actions'])
        if g['use_action_masks']:
            assert torch.equal(ag.dataset['action_masks'], ds['action_masks'])
        if ds.get('rnn_masks') is not None:
            assert torch.equal(ag.dataset['rnn_masks'], ds['rnn_masks'])
        torch.testing.assert_close(ag.buf['rewards'], ref['mb_rewards'], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(ag.buf['values'], ref['mb_values'], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ag.dataset['old_logp_actions'], ds['old_logp_actions'], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ag.dataset['advantages'], ds['advantages'], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(ag.dataset['returns'], ds['returns'], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(torch.stack(out['a_loss']), ref['a_losses'], rtol=1e-3, atol=1e-6)
        torch.testing.assert_close(torch.stack(out['c_loss']), ref['c_losses'], rtol=1e-3, atol=1e-6)
        torch.testing.assert_close(torch.stack(out['entropy']), ref['entropies'], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(torch.stack(out['kl']), ref['kls'], rtol=2e-3, atol=1e-8)        # one mean KL per mini-epoch
        assert ag.last_lr == pytest.approx(ref['last_lr'], rel=1e-12)
        for k in g['param_order']:
            torch.testing.assert_close(ag.model.p[k].detach(), ref['state'][k], rtol=1e-4, atol=2e-6, msg=lambda m: k + m)
        if g['config'].get('normalize_input'):
            st = ref['state']
            torch.testing.assert_close(ag.model.running_mean_std.running_mean, st['running_mean_std.running_mean'], rtol=1e-9, atol=1e-9)
            assert int(ag.model.running_mean_std.count) == int(st['running_mean_std.count'])
            torch.testing.assert_close(ag.model.value_mean_std.running_var, st['value_mean_std.running_var'], rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(ag.game_rewards.mean, ref['game_rewards_mean'], rtol=1e-5, atol=1e-6)
        assert ag.game_rewards.current_size == ref['game_rewards_size']


def test_inverse_cdf_sampling_and_masked_categorical():
    """the sampling rule of the discrete path (shared by golden generation, oracle and the future kernel) and CategoricalMasked
    against torch.distributions on random logits"""
    from oracle import ppo_discrete_oracle as DO
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(64, 5, generator=g) * 2
    masks = torch.rand(64, 5, generator=g) < 0.6
    masks[:, 0] |= ~masks.any(dim=1)
    nl, probs, ent = DO.categorical_masked(logits, None)
    ref = torch.distributions.Categorical(logits=logits)
    torch.testing.assert_close(nl, ref.logits); torch.testing.assert_close(ent, ref.entropy())
    nl, probs, ent = DO.categorical_masked(logits, masks)
    assert (probs[~masks] < 1e-30).all() and torch.allclose(probs.sum(-1), torch.ones(64))
    u = torch.rand(64, generator=g)
    a = DO.sample_inverse_cdf(probs, u)
    assert masks.gather(1, a.unsqueeze(1)).all()                     # never samples an illegal action
    assert (DO.sample_inverse_cdf(probs, torch.zeros(64)) == masks.float().argmax(dim=1)).all()     # u = 0 -> first legal action
    # empirical frequencies follow the probabilities
    p1 = torch.tensor([[0.1, 0.2, 0.3, 0.4]]).expand(20000, 4)
    cnt = torch.bincount(DO.sample_inverse_cdf(p1, torch.rand(20000, generator=g)), minlength=4).float() / 20000
    torch.testing.assert_close(cnt, p1[0], rtol=0, atol=0.02)


def test_ema_advantage_normaliser_bitexact():
    """GeneralizedMovingStats 'mean_std' (the reference's own test is bit-identity too:
    tests/test_rms_advantage.py:42-70): unmasked, masked and all-invalid updates, then eval normalise / denormalise"""
    g = load('rms_adv.pt')
    gms = O.GeneralizedMovingStats((1,), decay=g['decay'])
    for s_ in g['seq']:
        y = gms(s_['x'], mask=s_['mask']) if s_['mask'] is not None else gms(s_['x'])
        assert torch.equal(y, s_['y'])
        assert torch.equal(gms.step, s_['step']) and torch.equal(gms.mean, s_['mean']) and torch.equal(gms.sqrs, s_['sqrs'])
    gms.eval()
    assert torch.equal(gms(g['x_eval']), g['y_eval'])
    assert torch.equal(gms(g['x_eval'], denorm=True), g['y_denorm'])


# ------------------------------------------------------------------------------------------ central value
def _cv_oracle_from_golden(g):
    cfgk = g['config']
    cfg = {k: cfgk[k] for k in ('gamma', 'tau', 'e_clip', 'clip_value', 'critic_coef', 'entropy_coef', 'bound_loss_type', 'use_smooth_clamp',
                                'truncate_grads', 'grad_norm', 'learning_rate', 'kl_threshold', 'normalize_input', 'normalize_value',
                                'normalize_advantage', 'value_bootstrap', 'mini_epochs') if k in cfgk}
    cfg['bounds_loss_coef'] = cfgk.get('bounds_loss_coef', None)
    cfg['lr_schedule'] = cfgk.get('lr_schedule', None)
    cfg['mask_autoreset_rows'] = g['autoreset'] == 'next_step'

    class Env(O.TapeEnv):
        def __init__(self):
            super().__init__(g['obs_tape'], g['done_tape'], g['timeout_tape'])

        def reset(self):
            return {'obs': super().reset(), 'states': g['state_tape'][0].clone()}

        def step(self, actions):
            o, r, d, info = super().step(actions)
            return {'obs': o, 'states': g['state_tape'][self.i % g['state_tape'].shape[0]].clone()}, r, d, info
    cv_params = {k: v for k, v in g['cv_init_state'].items() if k.startswith('a2c_network')}
    cv = O.CentralValueOracle(cv_params, g['S'], g['cv_units'], g['cv_config'], cfg['normalize_value'], g['N'], g['H'])
    params = {k: v for k, v in g['init_state'].items() if k.startswith('a2c_network')}
    ag = O.OracleAgent(Env(), params, g['D'], g['A'], g['units'], g['N'], g['H'], g['mb'], cfg, central_value=cv)
    ag.obs = ag.env_reset()
    return ag, cv


def test_central_value_train_epochs_match_reference_agent():
    """A2CAgent with central_value_config (asymmetric critic on privileged `states`, own normalisers / Adam / minibatching,
    trained before the actor's mini-epochs; the rollout values and the value normaliser come from the critic) vs the oracle."""
    g = load('agent_cv.pt')
    assert g['cv_param_order'] == O.cv_param_names(len(g['cv_units']))
    ag, cv = _cv_oracle_from_golden(g)
    # with a central value the last-value forward goes through the critic and draws no action noise, so the reference consumed the
    # tape as one flat stream of H draws per epoch (gen_golden.py asserts counter == epochs * H)
    flat_noise = g['noise'].reshape(-1, g['N'], g['A'])
    for ep, ref in enumerate(g['epochs_out']):
        out = ag.train_epoch(flat_noise[ep * g['H']:(ep + 1) * g['H']])
        ds = ref['dataset']
        torch.testing.assert_close(ag.buf['values'], ref['mb_values'], rtol=1e-5, atol=1e-6)        # critic values in the rollout
        torch.testing.assert_close(ag.buf['rewards'], ref['mb_rewards'], rtol=1e-6, atol=1e-6)      # incl. the time-out bootstrap on them
        torch.testing.assert_close(ag.dataset['advantages'], ds['advantages'], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(ag.dataset['returns'], ds['returns'], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(ag.dataset['old_values'], ds['old_values'], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(torch.stack(out['a_loss']), ref['a_losses'], rtol=1e-3, atol=1e-6)
        torch.testing.assert_close(torch.stack(out['c_loss']), ref['c_losses'], rtol=1e-3, atol=1e-6)
        assert ag.last_lr == pytest.approx(ref['last_lr'], rel=1e-12) and cv.lr == pytest.approx(ref['cv_lr'], rel=1e-12)
        for k in O.param_names(len(g['units'])):
            torch.testing.assert_close(ag.model.p[k].detach(), ref['state'][k], rtol=1e-4, atol=2e-6, msg=lambda m: k + m)
        st = ref['cv_state']
        for k in O.cv_param_names(len(g['cv_units'])):
            torch.testing.assert_close(cv.p[k].detach(), st[k], rtol=1e-4, atol=2e-6, msg=lambda m: 'cv ' + k + m)
        torch.testing.assert_close(cv.running_mean_std.running_mean, st['running_mean_std.running_mean'], rtol=1e-9, atol=1e-9)
        assert int(cv.running_mean_std.count) == int(st['running_mean_std.count'])
        torch.testing.assert_close(cv.value_mean_std.running_mean, st['value_mean_std.running_mean'], rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(cv.value_mean_std.running_var, st['value_mean_std.running_var'], rtol=1e-6, atol=1e-7)
        assert int(cv.value_mean_std.count) == int(st['value_mean_std.count'])
        # the actor model's own value normaliser never moves when a central value exists
        assert int(ref['state']['value_mean_std.count']) == 1


def test_categorical_head_gradient_formulas_match_autograd():
    """The closed-form gradients that csrc/discrete.cu::categorical_loss_kernel implements, restated with torch ops and checked
    against autograd through the oracle's loss (the CUDA kernel itself is checked on a GPU by tests/test_discrete_gpu.py):
        dnlp/dz_k = p_k - [k == a];  dH/dz_k = -p_k (log p_k + H);  d(actor)/dnlp = adv * f'(ratio) * ratio on the active branch;
        d(critic)/dV from the clipped / unclipped branch;  every term weighted by mask / sum(mask)."""
    from oracle import ppo_discrete_oracle as DO
    g = torch.Generator().manual_seed(5)
    M, K, e_clip, critic_coef, ent_coef = 300, 6, 0.2, 1.0, 0.01
    for masked in (False, True):
        logits = (torch.randn(M, K, generator=g) * 1.5).requires_grad_(True)
        value = torch.randn(M, 1, generator=g).requires_grad_(True)
        amask = None
        if masked:
            amask = torch.rand(M, K, generator=g) < 0.7
            amask[torch.arange(M), torch.randint(0, K, (M,), generator=g)] = True
        nl, probs, ent = DO.categorical_masked(logits, amask)
        actions = DO.sample_inverse_cdf(probs.detach(), torch.rand(M, generator=g))
        old_nlp = -nl.detach().gather(1, actions.unsqueeze(1)).squeeze(1) + torch.randn(M, generator=g) * 0.2
        adv, old_v, ret = torch.randn(M, generator=g), torch.randn(M, 1, generator=g), torch.randn(M, 1, generator=g)
        rmask = (torch.rand(M, generator=g) < 0.8).float() if masked else None
        nlp = -nl.gather(1, actions.unsqueeze(1)).squeeze(1)
        a = O.actor_loss(old_nlp, nlp, adv, True, e_clip, smooth=False)
        c = O.critic_loss(old_v, value, e_clip, ret, True)
        losses, _ = O.apply_masks([a.unsqueeze(1), c, ent.unsqueeze(1)], rmask)
        (losses[0] + 0.5 * losses[1] * critic_coef - losses[2] * ent_coef).backward()
        # ---- the kernel's arithmetic ----
        with torch.no_grad():
            z = logits.detach() if amask is None else torch.where(amask, logits.detach(), torch.tensor(-1e8))
            lse = torch.logsumexp(z, dim=1, keepdim=True)
            lp = z - lse
            p = lp.exp()
            legal = torch.ones_like(p, dtype=torch.bool) if amask is None else amask
            H = -(torch.where(legal, p * lp, torch.zeros_like(p))).sum(1, keepdim=True)
            nlp_k = -lp.gather(1, actions.unsqueeze(1)).squeeze(1)
            w = (torch.ones(M) / M) if rmask is None else rmask / rmask.sum()
            ratio = torch.exp(old_nlp - nlp_k)
            clamped = ratio.clamp(1 - e_clip, 1 + e_clip)
            dcl = ((ratio >= 1 - e_clip) & (ratio <= 1 + e_clip)).float()
            t1, t2 = -(adv * ratio), -(adv * clamped)
            d1, d2 = adv * ratio, adv * dcl * ratio
            g_a = torch.where(t1 > t2, d1, torch.where(t1 < t2, d2, 0.5 * (d1 + d2)))
            onehot = torch.nn.functional.one_hot(actions, K).float()
            dz = w.unsqueeze(1) * (g_a.unsqueeze(1) * (p - onehot) + ent_coef * p * (lp + H))
            dz = torch.where(legal, dz, torch.zeros_like(dz))
            val = value.detach()
            delta = val - old_v
            vpc = old_v + delta.clamp(-e_clip, e_clip)
            e1, e2 = val - ret, vpc - ret
            l1, l2 = e1 * e1, e2 * e2
            g1, g2 = 2 * e1, torch.where((delta >= -e_clip) & (delta <= e_clip), 2 * e2, torch.zeros_like(e2))
            dc = torch.where(l1 > l2, g1, torch.where(l1 < l2, g2, 0.5 * (g1 + g2)))
            dv = w.unsqueeze(1) * 0.5 * critic_coef * dc
        torch.testing.assert_close(dz, logits.grad, rtol=1e-4, atol=1e-8)
        torch.testing.assert_close(dv, value.grad, rtol=1e-5, atol=1e-9)


def test_prepare_dataset_works_on_a_bare_agent_like_the_gpu_kernel_test_uses_it():
    """tests/test_kernels_gpu.py::test_prepare_batch_vs_oracle calls OracleAgent.prepare_dataset on an agent built with __new__ (no env, no
    central value): keep that entry point usable -- this is the CPU guard for a GPU-only test"""
    g = torch.Generator().manual_seed(7)
    B = 64
    ag = O.OracleAgent.__new__(O.OracleAgent)
    ag.cfg = dict(O.DEFAULT_CFG)
    ag.model = O.OracleModel(O.init_params(4, [8], 2), 4, [8], 2)
    batch = {'returns': torch.randn(B, 1, generator=g), 'values': torch.randn(B, 1, generator=g), 'neglogpacs': torch.zeros(B),
             'actions': torch.zeros(B, 2), 'obses': torch.zeros(B, 4), 'dones': torch.zeros(B), 'mus': torch.zeros(B, 2), 'sigmas': torch.ones(B, 2),
             'rnn_masks': (torch.rand(B, generator=g) < 0.8).float()}
    ag.prepare_dataset(batch)
    assert torch.isfinite(ag.dataset['advantages']).all() and int(ag.model.value_mean_std.count) > 1
