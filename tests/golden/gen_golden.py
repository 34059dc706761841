"""Generate golden vectors by running the REAL reference (Denys88/rl_games @ /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/gen_golden.py

Writes small ``tests/golden/*.pt`` fixtures (committed).  ``gymnasium`` and ``tensorboardX`` are not
installed here, so the import stubs under ``tests/golden/_stubs`` are put on ``sys.path`` -- the hot
path only uses gymnasium to describe shapes/dtypes (reference experience.py:385-398).

Action sampling: the reference samples with ``torch.distributions.Normal.sample`` ->
``torch.normal(loc, scale)``; this script swaps ``torch.normal`` for ``loc + scale * noise_tape[k]``
so the sampled noise is part of the fixture and the oracle / CUDA path can replay it.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '_stubs'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_num_threads(1)


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print('wrote', path, os.path.getsize(path), 'bytes')


# ------------------------------------------------------------------ GAE
def gen_gae():
    from rl_games.triton_kernels.gae_kernel import _pytorch_gae, compute_gae
    sys.path.insert(0, '/root/reference/tests')
    from test_triton_gae import make_inputs, reference_gae  # the reference's own KAT helpers
    cases = []
    for shape in [(8, 4, 1), (16, 6, 3), (8, 16, 1), (36, 64, 3), (32, 40, 1)]:
        for gamma, tau in [(0.99, 0.95), (1.0, 1.0)]:
            inp = make_inputs(*shape)
            out = _pytorch_gae(*inp, gamma, tau)
            assert torch.equal(out, compute_gae(*inp, gamma, tau))
            ref64 = reference_gae(*inp, gamma, tau) if shape[1] <= 16 else None
            cases.append({'shape': shape, 'gamma': gamma, 'tau': tau, 'inputs': inp, 'pytorch_gae': out,
                          'scalar_f64_as_f32': ref64})
    # edge cases from tests/test_triton_gae.py:102-111 (all-done / no-done)
    rewards, values, dones, lv, ld = make_inputs(10, 4, 1)
    for fill in (0.0, 1.0):
        d, l = torch.full_like(dones, fill), torch.full_like(ld, fill)
        cases.append({'shape': (10, 4, 1), 'gamma': 0.99, 'tau': 0.95, 'inputs': (rewards, values, d, lv, l),
                      'pytorch_gae': _pytorch_gae(rewards, values, d, lv, l, 0.99, 0.95), 'scalar_f64_as_f32': None})
    save('gae.pt', cases)


# ------------------------------------------------------------------ RunningMeanStd / masks / losses
def gen_math():
    from rl_games.algos_torch.running_mean_std import RunningMeanStd
    from rl_games.algos_torch import torch_ext
    from rl_games.common import common_losses
    g = torch.Generator().manual_seed(1)
    out = {}
    # running mean std: three sequential train-mode updates, then eval normalise + denorm
    rms = RunningMeanStd((6,))
    rms.train()
    xs = [torch.randn(37, 6, generator=g) * 3 + 1.5, torch.randn(5, 6, generator=g) - 4, torch.randn(64, 6, generator=g) * 0.1]
    ys = [rms(x) for x in xs]
    rms.eval()
    xe = torch.randn(9, 6, generator=g) * 10
    out['rms'] = {'xs': xs, 'ys': ys, 'state': {k: v.clone() for k, v in rms.state_dict().items()},
                  'x_eval': xe, 'y_eval': rms(xe), 'y_denorm': rms(xe, denorm=True)}
    # masked update (value normaliser on valid rows: a2c_common.py:1605-1615)
    rmsm = RunningMeanStd((1,))
    rmsm.train()
    xv = torch.randn(50, 1, generator=g) * 2 + 3
    mask = (torch.rand(50, generator=g) < 0.7)
    rmsm(xv[mask])
    out['rms_valid_rows'] = {'x': xv, 'mask': mask, 'state': {k: v.clone() for k, v in rmsm.state_dict().items()}}
    # masked moments / normalisation
    v = torch.randn(41, generator=g) * 2 + 0.5
    m = (torch.rand(41, generator=g) < 0.6).float()
    mean, var = torch_ext.get_mean_var_with_masks(v, m)
    out['masked'] = {'v': v, 'm': m, 'mean': mean, 'var': var,
                     'norm_masked': torch_ext.normalization_with_masks(v, m),
                     'norm_unmasked': torch_ext.normalization_with_masks(v, None)}
    # degenerate masks (tests/test_ppo_masking.py:568-584)
    for name, mm in [('zero', torch.zeros(41)), ('one', torch.cat([torch.ones(1), torch.zeros(40)]))]:
        mean, var = torch_ext.get_mean_var_with_masks(v, mm)
        out['masked_' + name] = {'m': mm, 'mean': mean, 'var': var, 'norm': torch_ext.normalization_with_masks(v, mm)}
    # losses
    B, A = 33, 5
    old_nlp = torch.randn(B, generator=g) * 0.3 + 4
    nlp = old_nlp + torch.randn(B, generator=g) * 0.3
    adv = torch.randn(B, generator=g)
    out['actor'] = {'old': old_nlp, 'new': nlp, 'adv': adv,
                    'hard': common_losses.actor_loss(old_nlp, nlp, adv, True, 0.2),
                    'smooth': common_losses.smoothed_actor_loss(old_nlp, nlp, adv, True, 0.2)}
    vp, vv, rr = torch.randn(B, 1, generator=g), torch.randn(B, 1, generator=g), torch.randn(B, 1, generator=g)
    out['critic'] = {'old_values': vp, 'values': vv, 'returns': rr,
                     'clip': common_losses.critic_loss(None, vp, vv, 0.2, rr, True),
                     'noclip': common_losses.critic_loss(None, vp, vv, 0.2, rr, False)}
    mu0, mu1 = torch.randn(B, A, generator=g), torch.randn(B, A, generator=g)
    s0, s1 = torch.rand(B, A, generator=g) + 0.2, torch.rand(B, A, generator=g) + 0.2
    out['kl'] = {'mu0': mu0, 's0': s0, 'mu1': mu1, 's1': s1, 'kl': torch_ext.policy_kl(mu0, s0, mu1, s1),
                 'kl_rows': torch_ext.policy_kl(mu0, s0, mu1, s1, False)}
    yp, yy = torch.randn(B, 1, generator=g), torch.randn(B, 1, generator=g)
    out['diag'] = {'y_pred': yp, 'y': yy, 'ev': torch_ext.explained_variance(yp, yy),
                   'clip_frac': torch_ext.policy_clip_fraction(nlp, old_nlp, 0.2)}
    # AverageMeter
    am = torch_ext.AverageMeter(1, 10)
    seq = [torch.randn(k, 1, generator=g) for k in (3, 0, 12, 4)]
    means = []
    for s in seq:
        am.update(s)
        means.append((am.mean.clone(), am.current_size))
    out['meter'] = {'seq': seq, 'means': means}
    # schedulers (tests/test_perf_fixes.py:204-222)
    from rl_games.common import schedulers
    sch = schedulers.AdaptiveScheduler(0.008)
    lr, lrs = 3e-4, []
    kls = [0.001, 0.02, 0.005, 0.017, 0.0039, 0.1, 0.1, 0.1, 0.0, 0.0]
    for k in kls:
        lr, _ = sch.update(lr, 0.0, 0, 0, k)
        lrs.append(lr)
    out['adaptive'] = {'kls': kls, 'lrs': lrs}
    save('math.pt', out)


# ------------------------------------------------------------------ EMA advantage normaliser
def gen_rms_adv():
    from rl_games.algos_torch.moving_mean_std import GeneralizedMovingStats
    g = torch.Generator().manual_seed(2)
    gms = GeneralizedMovingStats((1,), decay=0.5)
    gms.train()
    seq = []
    for k, (n, masked) in enumerate([(64, False), (37, True), (16, 'none_valid'), (128, False), (5, True)]):
        x = torch.randn(n, generator=g) * (1 + k) + 0.3 * k
        if masked == 'none_valid':
            mask = torch.zeros(n)
        elif masked:
            mask = (torch.rand(n, generator=g) < 0.5).float()
        else:
            mask = None
        y = gms(x, mask=mask) if mask is not None else gms(x)
        seq.append({'x': x, 'mask': mask, 'y': y, 'step': gms.step.clone(), 'mean': gms.mean.clone(), 'sqrs': gms.sqrs.clone()})
    gms.eval()
    xe = torch.randn(9, generator=g) * 30
    save('rms_adv.pt', {'decay': 0.5, 'seq': seq, 'x_eval': xe, 'y_eval': gms(xe), 'y_denorm': gms(xe, denorm=True)})


# ------------------------------------------------------------------ full agent epochs
class TapeVecEnv:
    """Tensor env fed from tapes (mirrors oracle.ppo_oracle.TapeEnv)."""

    def __init__(self, obs_tape, done_tape, timeout_tape, autoreset_mode='same_step'):
        self.obs_tape, self.done_tape, self.timeout_tape = obs_tape, done_tape, timeout_tape
        self.i = 0
        self.autoreset_mode = autoreset_mode

    def reset(self):
        self.i = 0
        return self.obs_tape[0].clone()

    def step(self, actions):
        rew = -(actions * actions).sum(-1) * 0.1
        self.i += 1
        j = self.i % self.obs_tape.shape[0]
        return self.obs_tape[j].clone(), rew, self.done_tape[j].clone(), {'time_outs': self.timeout_tape[j].clone()}

    def get_env_info(self):
        import gymnasium as gym
        D, A = self.obs_tape.shape[-1], self.A
        lo, hi = getattr(self, 'act_bounds', (-1.0, 1.0))
        info = {'observation_space': gym.spaces.Box(-np.inf, np.inf, (D,), np.float32),
                'action_space': gym.spaces.Box(lo, hi, (A,), np.float32)}
        if self.autoreset_mode != 'same_step':
            info['autoreset_mode'] = self.autoreset_mode
        return info

    def get_env_state(self):
        return None

    def set_env_state(self, s):
        pass

    def set_train_info(self, *a, **kw):
        pass


def make_params(N, H, mb, units, overrides=None, rnn_units=0, rnn_before_mlp=True, space_over=None, network_over=None):
    network = {
        'name': 'actor_critic', 'separate': False,
        'space': {'continuous': {'mu_activation': 'None', 'sigma_activation': 'None',
                                 'mu_init': {'name': 'default'},
                                 'sigma_init': {'name': 'const_initializer', 'val': 0},
                                 'fixed_sigma': True}},
        'mlp': {'units': list(units), 'activation': 'elu', 'initializer': {'name': 'default'}},
    }
    network['space']['continuous'].update(space_over or {})          # e.g. min_sigma (configs/mjlab/ppo_lift_cube_yam.yaml)
    network.update(network_over or {})                               # e.g. separate: True (configs/ppo_continuous.yaml)
    if rnn_units:
        network['rnn'] = {'name': 'lstm', 'units': rnn_units, 'layers': 1, 'before_mlp': rnn_before_mlp}
    # hyper-parameters of configs/mujoco/ant_envpool.yaml:28-56
    config = {
        'name': 'golden', 'env_name': 'unused', 'reward_shaper': {'scale_value': 1.0},
        'device': 'cpu', 'multi_gpu': False, 'mixed_precision': False, 'torch_compile': False,
        'normalize_input': True, 'normalize_value': True, 'value_bootstrap': True, 'normalize_advantage': True,
        'gamma': 0.99, 'tau': 0.95, 'learning_rate': 3e-4, 'lr_schedule': 'adaptive', 'kl_threshold': 0.008,
        'grad_norm': 1.0, 'entropy_coef': 0.0, 'truncate_grads': True, 'e_clip': 0.2, 'clip_value': True,
        'use_smooth_clamp': True, 'bound_loss_type': 'regularisation', 'bounds_loss_coef': 0.0,
        'max_epochs': 100, 'num_actors': N, 'horizon_length': H, 'minibatch_size': mb, 'mini_epochs': 4,
        'critic_coef': 2, 'save_frequency': 0, 'save_best_after': 10_000, 'print_stats': False,
        'train_dir': '/tmp/golden_runs',
    }
    config.update(overrides or {})
    return {'algo': {'name': 'a2c_continuous'}, 'model': {'name': 'continuous_a2c_logstd'},
            'network': network, 'config': config}


def gen_agent(name, N=8, H=8, D=6, A=3, units=(16, 8), mb=32, epochs=2, overrides=None, autoreset='same_step',
              seed=3, rnn_units=0, rnn_before_mlp=True, act_bounds=(-1.0, 1.0), train_loop=False, space_over=None, network_over=None):
    from rl_games.torch_runner import Runner
    from oracle.ppo_oracle import make_tapes
    torch.manual_seed(seed)
    np.random.seed(seed)
    T = H * epochs + 1
    obs_tape, done_tape, tout_tape = make_tapes(T, N, D, seed=seed)
    env = TapeVecEnv(obs_tape, done_tape, tout_tape, autoreset)
    env.A = A
    env.act_bounds = act_bounds
    params = make_params(N, H, mb, units, overrides, rnn_units, rnn_before_mlp, space_over, network_over)
    params['config']['env_info'] = env.get_env_info()
    shaper_cfg = dict(params['config']['reward_shaper'])        # Runner.load_config replaces the dict by a DefaultRewardsShaper object
    runner = Runner()
    runner.load({'params': params})
    runner.params['config']['vec_env'] = env
    agent = runner.algo_factory.create(runner.algo_name, base_name='golden', params=runner.params)
    init_state = {k: v.clone() for k, v in agent.model.state_dict().items()}
    # perturb sigma / biases a little so nothing is exactly zero
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for k, p in agent.model.named_parameters():
            if k.endswith('bias') or k.endswith('sigma'):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    init_state = {k: v.clone() for k, v in agent.model.state_dict().items()}

    noise = torch.randn(epochs, H + 1, N, A, generator=g)   # [epoch, step (H rollout + 1 last-value fwd)]
    counter = {'k': 0}
    orig_normal = torch.normal

    def fake_normal(loc, scale, *a, **kw):
        k = counter['k']
        counter['k'] += 1
        e, n = divmod(k, H + 1)
        return loc + scale * noise[e, n]
    torch.normal = fake_normal
    try:
        agent.init_tensors()
        agent.obs = agent.env_reset()
        epochs_out = []
        loop_out = {}

        def snapshot(res):
            step_time, play_time, update_time, total, a_losses, c_losses, b_losses, entropies, kls, last_lr, lr_mul = res
            ds = agent.dataset.values_dict
            epochs_out.append({
                'a_losses': torch.stack([x.detach() for x in a_losses]), 'c_losses': torch.stack([x.detach() for x in c_losses]),
                'entropies': torch.stack([x.detach() for x in entropies]), 'kls': torch.stack([x.detach() for x in kls]),
                'b_losses': torch.stack([x.detach() for x in b_losses]) if len(b_losses) else None,
                'last_lr': agent.last_lr,
                'state': {k: v.clone() for k, v in agent.model.state_dict().items()},
                'dataset': {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ds.items() if k != 'rnn_states'},
                'mb_rewards': agent.experience_buffer.tensor_dict['rewards'].clone(),
                'mb_values': agent.experience_buffer.tensor_dict['values'].clone(),
                'mb_dones': agent.experience_buffer.tensor_dict['dones'].clone(),
                'game_rewards_mean': agent.game_rewards.mean.clone(), 'game_rewards_size': agent.game_rewards.current_size,
                'game_lengths_mean': agent.game_lengths.mean.clone(),
                'adam_exp_avg': [agent.optimizer.state[p]['exp_avg'].clone() for p in agent.model.parameters()],
                'adam_exp_avg_sq': [agent.optimizer.state[p]['exp_avg_sq'].clone() for p in agent.model.parameters()],
                'frame_before': agent.frame, 'epoch_num': agent.epoch_num,
            })
        if train_loop:
            # the reference's own outer loop (a2c_common.py:1662-1782): frame / epoch accounting, stop conditions, checkpoints
            orig_epoch = agent.train_epoch

            def wrapped():
                res = orig_epoch()
                snapshot(res)
                return res
            agent.train_epoch = wrapped
            scalars = []
            agent.writer = type('W', (), {'add_scalar': lambda self, tag, v, step=None: scalars.append((tag, float(v), step)),
                                          'flush': lambda self: None, 'close': lambda self: None})()
            agent.algo_observer.writer = agent.writer
            ret = agent.train()
            loop_out = {'return': (float(ret[0]), int(ret[1])), 'frame': int(agent.frame), 'epoch_num': int(agent.epoch_num),
                        'last_mean_rewards': float(agent.last_mean_rewards), 'mean_rewards': float(agent.mean_rewards),
                        'saved': sorted(os.listdir(agent.nn_dir)), 'scalars': scalars}
        else:
            for ep in range(epochs):
                agent.epoch_num += 1
                snapshot(agent.train_epoch())
                agent.dataset.update_values_dict(None)
        assert counter['k'] == epochs * (H + 1), counter
    finally:
        torch.normal = orig_normal
    save(name, {'N': N, 'H': H, 'D': D, 'A': A, 'units': list(units), 'mb': mb, 'epochs': epochs,
                'config': {k: v for k, v in params['config'].items() if isinstance(v, (int, float, str, bool, type(None)))},
                'reward_shaper': shaper_cfg, 'space_over': dict(space_over or {}), 'network_over': dict(network_over or {}),
                'act_bounds': tuple(act_bounds), 'autoreset': autoreset, 'rnn_units': rnn_units, 'rnn_before_mlp': rnn_before_mlp, 'obs_tape': obs_tape, 'done_tape': done_tape, 'timeout_tape': tout_tape,
                'noise': noise, 'init_state': init_state, 'epochs_out': epochs_out, 'train_loop': loop_out,
                'param_order': [k for k, _ in agent.model.named_parameters()]})


# ------------------------------------------------------------------ central value / asymmetric critic
class CVTapeVecEnv(TapeVecEnv):
    """TapeVecEnv whose observations are {'obs': actor view, 'states': privileged critic view} (state tape = [obs, extra features])"""

    def __init__(self, obs_tape, state_tape, done_tape, timeout_tape, autoreset_mode='same_step'):
        super().__init__(obs_tape, done_tape, timeout_tape, autoreset_mode)
        self.state_tape = state_tape

    def reset(self):
        self.i = 0
        return {'obs': self.obs_tape[0].clone(), 'states': self.state_tape[0].clone()}

    def step(self, actions):
        o, rew, done, info = super().step(actions)
        j = self.i % self.obs_tape.shape[0]
        return {'obs': o, 'states': self.state_tape[j].clone()}, rew, done, info

    def get_env_info(self):
        import gymnasium as gym
        info = super().get_env_info()
        info['state_space'] = gym.spaces.Box(-np.inf, np.inf, (self.state_tape.shape[-1],), np.float32)
        return info


def gen_agent_cv(name='agent_cv.pt', N=8, H=8, D=6, S=10, A=3, units=(16, 8), cv_units=(24, 12), mb=32, cv_mb=16, epochs=2, seed=8,
                 autoreset='next_step'):
    from rl_games.torch_runner import Runner
    from oracle.ppo_oracle import make_tapes
    torch.manual_seed(seed)
    np.random.seed(seed)
    T = H * epochs + 1
    obs_tape, done_tape, tout_tape = make_tapes(T, N, D, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    state_tape = torch.cat([obs_tape, torch.randn(T, N, S - D, generator=g) * 2.0 - 0.5], dim=-1)
    env = CVTapeVecEnv(obs_tape, state_tape, done_tape, tout_tape, autoreset)
    env.A = A
    cv_cfg = {'minibatch_size': cv_mb, 'mini_epochs': 3, 'learning_rate': 5e-4, 'clip_value': True, 'normalize_input': True,
              'truncate_grads': True, 'grad_norm': 1.0,
              'network': {'name': 'actor_critic', 'central_value': True,
                          'mlp': {'units': list(cv_units), 'activation': 'elu', 'initializer': {'name': 'default'}}}}
    params = make_params(N, H, mb, units, {'central_value_config': cv_cfg})
    params['config']['env_info'] = env.get_env_info()
    runner = Runner()
    runner.load({'params': params})
    runner.params['config']['vec_env'] = env
    agent = runner.algo_factory.create(runner.algo_name, base_name='golden', params=runner.params)
    with torch.no_grad():
        for mdl in (agent.model, agent.central_value_net.model):
            for k, p in mdl.named_parameters():
                if k.endswith('bias') or k.endswith('sigma'):
                    p.add_(torch.randn(p.shape, generator=g) * 0.05)
    init_state = {k: v.clone() for k, v in agent.model.state_dict().items()}
    cv_init_state = {k: v.clone() for k, v in agent.central_value_net.model.state_dict().items()}
    noise = torch.randn(epochs, H + 1, N, A, generator=g)
    counter = {'k': 0}
    orig_normal = torch.normal

    def fake_normal(loc, scale, *a, **kw):
        k = counter['k']
        counter['k'] += 1
        e, n = divmod(k, H + 1)
        return loc + scale * noise[e, n]
    torch.normal = fake_normal
    try:
        agent.init_tensors()
        agent.obs = agent.env_reset()
        epochs_out = []
        cv_scalars = []
        agent.central_value_net.writter = type('W', (), {'add_scalar': lambda self, tag, v, step=None: cv_scalars.append((tag, float(v), step))})()
        for ep in range(epochs):
            agent.epoch_num += 1
            res = agent.train_epoch()
            step_time, play_time, update_time, total, a_losses, c_losses, b_losses, entropies, kls, last_lr, lr_mul = res
            ds = agent.dataset.values_dict
            cvn = agent.central_value_net
            epochs_out.append({
                'a_losses': torch.stack([x.detach() for x in a_losses]), 'c_losses': torch.stack([x.detach() for x in c_losses]),
                'entropies': torch.stack([x.detach() for x in entropies]), 'kls': torch.stack([x.detach() for x in kls]),
                'last_lr': agent.last_lr, 'state': {k: v.clone() for k, v in agent.model.state_dict().items()},
                'cv_state': {k: v.clone() for k, v in cvn.model.state_dict().items()}, 'cv_lr': cvn.lr,
                'cv_adam_exp_avg': [cvn.optimizer.state[p]['exp_avg'].clone() for p in cvn.model.parameters()],
                'dataset': {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ds.items() if k != 'rnn_states'},
                'mb_values': agent.experience_buffer.tensor_dict['values'].clone(),
                'mb_rewards': agent.experience_buffer.tensor_dict['rewards'].clone(),
            })
            agent.dataset.update_values_dict(None)
        # the normal draws: H rollout steps per epoch only -- get_values goes through the critic when a central value exists
        assert counter['k'] == epochs * H, counter
    finally:
        torch.normal = orig_normal
    save(name, {'N': N, 'H': H, 'D': D, 'S': S, 'A': A, 'units': list(units), 'cv_units': list(cv_units), 'mb': mb, 'epochs': epochs,
                'config': {k: v for k, v in params['config'].items() if isinstance(v, (int, float, str, bool, type(None)))},
                'cv_config': {k: v for k, v in cv_cfg.items() if isinstance(v, (int, float, str, bool, type(None)))},
                'autoreset': autoreset, 'obs_tape': obs_tape, 'state_tape': state_tape, 'done_tape': done_tape, 'timeout_tape': tout_tape,
                'noise': noise, 'init_state': init_state, 'cv_init_state': cv_init_state, 'epochs_out': epochs_out, 'cv_scalars': cv_scalars,
                'cv_param_order': [k for k, _ in agent.central_value_net.model.named_parameters()]})


# ------------------------------------------------------------------ checkpoint wire format
def gen_checkpoint(name='ref_checkpoint.pt', N=8, H=8, D=6, A=3, units=(16, 8), mb=32, seed=9):
    """A checkpoint dict exactly as the reference writes it (A2CBase.get_full_state_weights, a2c_common.py:825-850, after one
    train_epoch so that the Adam state exists) -- the input of the host-side interop test."""
    from rl_games.torch_runner import Runner
    from oracle.ppo_oracle import make_tapes
    torch.manual_seed(seed)
    np.random.seed(seed)
    obs_tape, done_tape, tout_tape = make_tapes(H + 1, N, D, seed=seed)
    env = TapeVecEnv(obs_tape, done_tape, tout_tape)
    env.A = A
    params = make_params(N, H, mb, units)
    params['config']['env_info'] = env.get_env_info()
    runner = Runner()
    runner.load({'params': params})
    runner.params['config']['vec_env'] = env
    agent = runner.algo_factory.create(runner.algo_name, base_name='golden', params=runner.params)
    agent.init_tensors()
    agent.obs = agent.env_reset()
    agent.epoch_num += 1
    agent.train_epoch()
    ck = agent.get_full_state_weights()
    ck = {k: v for k, v in ck.items() if k in ('model', 'epoch', 'frame', 'optimizer', 'last_mean_rewards')}
    save(name, {'D': D, 'A': A, 'units': list(units), 'checkpoint': ck, 'param_order': [k for k, _ in agent.model.named_parameters()]})


def gen_resume(name='agent_resume.pt', N=8, H=8, D=6, A=3, units=(16, 8), mb=32, seed=19):
    """Resume from a checkpoint: reference agent A trains one epoch and exports get_full_state_weights(); a FRESH reference agent B
    (different initial weights) imports it with set_full_state_weights(), resets the env and trains one more epoch.  The fixture holds
    the checkpoint and B's epoch: whoever loads that checkpoint into a fresh trainer must continue exactly like B (weights, Adam
    moments and step count, normaliser statistics, adaptive LR, epoch / frame counters)."""
    from rl_games.torch_runner import Runner
    from oracle.ppo_oracle import make_tapes
    obs_tape, done_tape, tout_tape = make_tapes(H + 1, N, D, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    noise = torch.randn(2, H + 1, N, A, generator=g)
    counter = {'k': 0, 'e': 0}
    orig_normal = torch.normal

    def fake_normal(loc, scale, *a, **kw):
        k = counter['k']
        counter['k'] += 1
        return loc + scale * noise[counter['e'], k]

    def make(seed_):
        torch.manual_seed(seed_)
        np.random.seed(seed_)
        env = TapeVecEnv(obs_tape, done_tape, tout_tape)
        env.A = A
        params = make_params(N, H, mb, units, {'weight_decay': 0.01})
        params['config']['env_info'] = env.get_env_info()
        runner = Runner()
        runner.load({'params': params})
        runner.params['config']['vec_env'] = env
        ag = runner.algo_factory.create(runner.algo_name, base_name='golden', params=runner.params)
        ag.init_tensors()
        return ag, params
    torch.normal = fake_normal
    try:
        a, params = make(seed)
        a.obs = a.env_reset()
        a.epoch_num += 1
        a.train_epoch()
        a.frame += a.batch_size                       # what train() does after the epoch (a2c_common.py:1686)
        ck = a.get_full_state_weights()
        ck = {k: v for k, v in ck.items() if k in ('model', 'epoch', 'frame', 'optimizer', 'last_mean_rewards')}
        ck = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ck.items()}
        import copy
        ck = copy.deepcopy(ck)
        b, _ = make(seed + 1)
        init_b = {k: v.clone() for k, v in b.model.state_dict().items()}
        b.set_full_state_weights(copy.deepcopy(ck))
        b.obs = b.env_reset()
        counter['k'], counter['e'] = 0, 1
        b.epoch_num += 1
        res = b.train_epoch()
        out = {'a_losses': torch.stack([x.detach() for x in res[4]]), 'c_losses': torch.stack([x.detach() for x in res[5]]),
               'kls': torch.stack([x.detach() for x in res[8]]), 'last_lr': b.last_lr, 'epoch_num': b.epoch_num, 'frame': b.frame,
               'state': {k: v.clone() for k, v in b.model.state_dict().items()},
               'adam_exp_avg': [b.optimizer.state[p]['exp_avg'].clone() for p in b.model.parameters()],
               'adam_step': float(b.optimizer.state[next(iter(b.model.parameters()))]['step']),
               'mb_values': b.experience_buffer.tensor_dict['values'].clone()}
    finally:
        torch.normal = orig_normal
    save(name, {'N': N, 'H': H, 'D': D, 'A': A, 'units': list(units), 'mb': mb, 'autoreset': 'same_step',
                'config': {k: v for k, v in params['config'].items() if isinstance(v, (int, float, str, bool, type(None)))},
                'obs_tape': obs_tape, 'done_tape': done_tape, 'timeout_tape': tout_tape, 'noise': noise, 'checkpoint': ck,
                'init_state': init_b, 'resumed_epoch': out, 'lr_at_checkpoint': a.last_lr})


# ------------------------------------------------------------------ discrete PPO (configs/ppo_cartpole.yaml shape)
class DiscreteTapeVecEnv:
    """mirrors oracle.ppo_discrete_oracle.DiscreteTapeEnv behind the reference's tensor-env contract"""

    def __init__(self, obs_tape, done_tape, timeout_tape, K, mask_tape=None, autoreset_mode='same_step'):
        self.obs_tape, self.done_tape, self.timeout_tape, self.K, self.mask_tape = obs_tape, done_tape, timeout_tape, K, mask_tape
        self.i = 0
        self.autoreset_mode = autoreset_mode

    def reset(self):
        self.i = 0
        return self.obs_tape[0].clone()

    def get_action_masks(self):
        return self.mask_tape[self.i % self.mask_tape.shape[0]].numpy()

    def step(self, actions):
        j0 = self.i % self.obs_tape.shape[0]
        if isinstance(self.K, (list, tuple)):
            off, rew = 0, 0.0
            for j, k in enumerate(self.K):
                rew = rew + (actions[:, j].long() == self.obs_tape[j0][:, off:off + k].argmax(dim=-1)).float() / len(self.K)
                off += k
        else:
            target = self.obs_tape[j0][:, :self.K].argmax(dim=-1)
            rew = (actions.long() == target).float()
        self.i += 1
        j = self.i % self.obs_tape.shape[0]
        return self.obs_tape[j].clone(), rew, self.done_tape[j].clone(), {'time_outs': self.timeout_tape[j].clone()}

    def get_env_info(self):
        import gymnasium as gym
        multi = isinstance(self.K, (list, tuple))
        info = {'observation_space': gym.spaces.Box(-np.inf, np.inf, (self.obs_tape.shape[-1],), np.float32),
                'action_space': gym.spaces.Tuple([gym.spaces.Discrete(k) for k in self.K]) if multi else gym.spaces.Discrete(self.K)}
        if self.autoreset_mode != 'same_step':
            info['autoreset_mode'] = self.autoreset_mode
        return info

    def get_env_state(self):
        return None

    def set_env_state(self, s):
        pass

    def set_train_info(self, *a, **kw):
        pass


def gen_agent_discrete(name, N=16, H=8, D=4, K=2, units=(32, 32), mb=64, epochs=2, overrides=None, separate=True,
                       use_action_masks=False, autoreset='same_step', seed=11, train_loop=False):
    from rl_games.torch_runner import Runner
    from oracle.ppo_oracle import make_tapes
    from oracle.ppo_discrete_oracle import sample_inverse_cdf
    torch.manual_seed(seed)
    np.random.seed(seed)
    T = H * epochs + 1
    obs_tape, done_tape, tout_tape = make_tapes(T, N, D, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    mask_tape = None
    multi = isinstance(K, (list, tuple))
    if use_action_masks:
        mask_tape = torch.rand(T, N, sum(K) if multi else K, generator=g) < 0.6
        off = 0
        for k in (K if multi else [K]):                            # at least one legal action per row and head
            forced = torch.randint(0, k, (T, N), generator=g) + off
            mask_tape.scatter_(2, forced.unsqueeze(-1), True)
            off += k
    env = DiscreteTapeVecEnv(obs_tape, done_tape, tout_tape, K, mask_tape, autoreset)
    network = {'name': 'actor_critic', 'separate': separate, 'space': {'multi_discrete' if multi else 'discrete': None},
               'mlp': {'units': list(units), 'activation': 'relu', 'initializer': {'name': 'default'}, 'regularizer': {'name': 'None'}}}
    # hyper-parameters of configs/ppo_cartpole.yaml:28-52
    config = {'name': 'golden_discrete', 'env_name': 'unused', 'reward_shaper': {'scale_value': 0.1}, 'normalize_advantage': True,
              'gamma': 0.99, 'tau': 0.9, 'learning_rate': 2e-4, 'grad_norm': 1.0, 'entropy_coef': 0.01, 'truncate_grads': True,
              'e_clip': 0.2, 'clip_value': True, 'num_actors': N, 'horizon_length': H, 'minibatch_size': mb, 'mini_epochs': 4,
              'critic_coef': 1, 'lr_schedule': None, 'kl_threshold': 0.008, 'normalize_input': False, 'normalize_value': False,
              'device': 'cpu', 'multi_gpu': False, 'mixed_precision': False, 'torch_compile': False, 'max_epochs': 100,
              'save_frequency': 0, 'save_best_after': 10_000, 'print_stats': False, 'train_dir': '/tmp/golden_runs',
              'use_action_masks': use_action_masks}
    config.update(overrides or {})
    params = {'algo': {'name': 'a2c_discrete'}, 'model': {'name': 'multi_discrete_a2c' if multi else 'discrete_a2c'}, 'network': network,
              'config': config}
    params['config']['env_info'] = env.get_env_info()
    runner = Runner()
    runner.load({'params': params})
    runner.params['config']['vec_env'] = env
    agent = runner.algo_factory.create(runner.algo_name, base_name='golden', params=runner.params)
    with torch.no_grad():
        for k, p in agent.model.named_parameters():
            if k.endswith('bias'):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    init_state = {k: v.clone() for k, v in agent.model.state_dict().items()}
    nh = len(K) if multi else 1
    # [epoch, step (H rollout + the get_values forward, which samples too), head]: one multinomial call per head, in head order
    u = torch.rand(epochs, H + 1, nh, N, generator=g)
    counter = {'k': 0}
    orig_multinomial = torch.multinomial

    def fake_multinomial(probs_2d, num_samples, replacement=False, **kw):
        assert num_samples == 1
        k = counter['k']
        counter['k'] += 1
        e, r = divmod(k, (H + 1) * nh)
        n, j = divmod(r, nh)
        return sample_inverse_cdf(probs_2d, u[e, n, j]).unsqueeze(-1)
    torch.multinomial = fake_multinomial
    try:
        agent.init_tensors()
        agent.obs = agent.env_reset()
        epochs_out = []
        loop_out = {}

        def snapshot(res):
            step_time, play_time, update_time, total, a_losses, c_losses, entropies, kls, last_lr, lr_mul = res
            ds = agent.dataset.values_dict
            epochs_out.append({
                'a_losses': torch.stack([x.detach() for x in a_losses]), 'c_losses': torch.stack([x.detach() for x in c_losses]),
                'entropies': torch.stack([x.detach() for x in entropies]), 'kls': torch.stack([x.detach() for x in kls]),
                'last_lr': agent.last_lr, 'state': {k: v.clone() for k, v in agent.model.state_dict().items()},
                'dataset': {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ds.items() if k != 'rnn_states'},
                'mb_rewards': agent.experience_buffer.tensor_dict['rewards'].clone(),
                'mb_values': agent.experience_buffer.tensor_dict['values'].clone(),
                'mb_actions': agent.experience_buffer.tensor_dict['actions'].clone(),
                'game_rewards_mean': agent.game_rewards.mean.clone(), 'game_rewards_size': agent.game_rewards.current_size,
                'adam_exp_avg': [agent.optimizer.state[p]['exp_avg'].clone() for p in agent.model.parameters()],
                'frame_before': agent.frame, 'epoch_num': agent.epoch_num,
            })
        if train_loop:      # DiscreteA2CBase.train (a2c_common.py:1361-1470): the reference's own outer loop
            orig_epoch = agent.train_epoch

            def wrapped():
                res = orig_epoch()
                snapshot(res)
                return res
            agent.train_epoch = wrapped
            scalars = []
            agent.writer = type('W', (), {'add_scalar': lambda self, tag, v, step=None: scalars.append((tag, float(v), step)),
                                          'flush': lambda self: None, 'close': lambda self: None})()
            agent.algo_observer.writer = agent.writer
            ret = agent.train()
            loop_out = {'return': (float(ret[0]), int(ret[1])), 'frame': int(agent.frame), 'epoch_num': int(agent.epoch_num),
                        'last_mean_rewards': float(agent.last_mean_rewards), 'mean_rewards': float(agent.mean_rewards),
                        'saved': sorted(os.listdir(agent.nn_dir)), 'scalars': scalars}
        else:
            for ep in range(epochs):
                agent.epoch_num += 1
                snapshot(agent.train_epoch())
                agent.dataset.update_values_dict(None)
        assert counter['k'] == epochs * (H + 1) * nh, counter
    finally:
        torch.multinomial = orig_multinomial
    if not multi:
        u = u[:, :, 0]
    save(name, {'N': N, 'H': H, 'D': D, 'K': K, 'units': list(units), 'mb': mb, 'epochs': epochs, 'separate': separate,
                'use_action_masks': use_action_masks, 'autoreset': autoreset,
                'config': {k: v for k, v in params['config'].items() if isinstance(v, (int, float, str, bool, type(None)))},
                'obs_tape': obs_tape, 'done_tape': done_tape, 'timeout_tape': tout_tape, 'mask_tape': mask_tape, 'u': u,
                'init_state': init_state, 'epochs_out': epochs_out, 'train_loop': loop_out,
                'param_order': [k for k, _ in agent.model.named_parameters()]})


if __name__ == '__main__':
    which = set(sys.argv[1:]) or {'gae', 'math', 'continuous', 'discrete', 'rmsadv', 'checkpoint', 'cv', 'tcshape', 'lstm_after', 'sched', 'misc', 'train', 'resume', 'lstm_masked', 'train_discrete', 'minsigma', 'separate'}      # e.g. `gen_golden.py discrete` regenerates only that group
    if 'gae' in which:
        gen_gae()
    if 'math' in which:
        gen_math()
    if 'continuous' in which:
        gen_agent('agent_base.pt')
        gen_agent('agent_masked.pt', autoreset='next_step', seed=4)
        gen_agent('agent_hardclip.pt', seed=5, overrides={
            'use_smooth_clamp': False, 'bound_loss_type': 'bound', 'bounds_loss_coef': 0.001, 'entropy_coef': 0.003,
            'clip_value': False, 'truncate_grads': False, 'value_bootstrap': False, 'mini_epochs': 2,
            'weight_decay': 0.01, 'lr_schedule': None})
        gen_agent('agent_lstm.pt', seed=6, rnn_units=8, overrides={'seq_length': 4})
    if 'separate' in which:
        # separate actor / critic trunks of a continuous policy (network_builder.py:494-512; configs/ppo_continuous.yaml, ppo_lunar.yaml ...):
        # masked rows, global-norm clip over both trunks, weight decay, an entropy bonus and a bound loss on top
        gen_agent('agent_separate.pt', seed=25, autoreset='next_step', network_over={'separate': True}, overrides={
            'weight_decay': 0.01, 'entropy_coef': 0.002, 'bounds_loss_coef': 0.001, 'bound_loss_type': 'bound', 'mini_epochs': 3})
    if 'minsigma' in which:
        # sigma floor of the 'exp' parametrisation (models.py:296-300; configs/mjlab/ppo_lift_cube_yam.yaml: min_sigma 0.15 with the hard
        # clip, a bound loss, an entropy bonus -- whose gradient reaches only sigma -- and unclipped actions), masked rows on top
        gen_agent('agent_minsigma.pt', seed=24, autoreset='next_step', space_over={'min_sigma': 0.15}, overrides={
            'use_smooth_clamp': False, 'bounds_loss_coef': 0.0001, 'entropy_coef': 0.005, 'clip_actions': False, 'clip_value': False,
            'learning_rate': 1e-3, 'kl_threshold': 0.01, 'critic_coef': 1, 'mini_epochs': 3})
    if 'checkpoint' in which:
        gen_checkpoint()
    if 'lstm_after' in which:
        # the placement most shipped configs use: MLP -> LSTM -> heads (before_mlp: False is the reference default)
        gen_agent('agent_lstm_after.pt', seed=14, rnn_units=12, rnn_before_mlp=False, overrides={'seq_length': 4})
    if 'misc' in which:
        # config keys no other fixture moves: linear LR + entropy schedule, all three normalisers off, the full reward shaper
        # (shift, scale, clip), unclipped actions into a non-unit action box, different gamma / tau / e_clip / critic_coef, a short meter
        gen_agent('agent_misc.pt', seed=16, epochs=3, act_bounds=(-2.0, 0.5), overrides={
            'lr_schedule': 'linear', 'max_epochs': 10, 'schedule_entropy': True, 'entropy_coef': 0.01, 'normalize_input': False,
            'normalize_value': False, 'normalize_advantage': False, 'clip_actions': False, 'games_to_track': 5, 'e_clip': 0.1,
            'critic_coef': 1, 'tau': 0.9, 'gamma': 0.95, 'mini_epochs': 2,
            'reward_shaper': {'scale_value': 0.5, 'shift_value': 0.2, 'min_val': -0.3, 'max_val': 0.25}})
        # clipped + rescaled actions into the non-unit box, vanilla policy gradient (ppo: False), masked rows, minibatch_size_per_env
        gen_agent('agent_rescale.pt', seed=17, autoreset='next_step', act_bounds=(-2.0, 0.5), overrides={
            'ppo': False, 'clip_actions': True, 'minibatch_size_per_env': 4, 'bounds_loss_coef': 0.001, 'bound_loss_type': 'bound'})
    if 'lstm_masked' in which:
        # LSTM policy on a next_step-autoreset env (envpool-style): filler reset rows are masked, the state absorbed on them is
        # re-zeroed in the rollout, and the train-time reset also fires entering the first real row (a2c_common.py:1097-1106, :1180-1191)
        gen_agent('agent_lstm_masked.pt', seed=20, rnn_units=8, autoreset='next_step', overrides={'seq_length': 4})
        gen_agent('agent_lstm_after_masked.pt', seed=21, rnn_units=12, rnn_before_mlp=False, autoreset='next_step', overrides={'seq_length': 4})
    if 'train_discrete' in which:
        import shutil
        shutil.rmtree('/tmp/golden_runs', ignore_errors=True)
        gen_agent_discrete('agent_discrete_trainloop.pt', separate=False, seed=22, epochs=3, autoreset='next_step', train_loop=True, overrides={
            'normalize_input': True, 'normalize_value': True, 'lr_schedule': 'adaptive', 'kl_threshold': 0.02, 'max_epochs': 3,
            'save_frequency': 2, 'save_best_after': 1, 'games_to_track': 10})
    if 'train' in which:
        # the outer loop: stops on max_frames, linear schedule driven by FRAMES, periodic + best + final checkpoints
        import shutil
        shutil.rmtree('/tmp/golden_runs', ignore_errors=True)
        gen_agent('agent_trainloop.pt', seed=18, epochs=3, train_loop=True, overrides={
            'lr_schedule': 'linear', 'max_epochs': -1, 'max_frames': 3 * 64, 'save_frequency': 2, 'save_best_after': 1, 'games_to_track': 10})
        # adaptive schedule: 'info/last_lr' logs the lr the LAST minibatch ran on (train_actor_critic's return value), not the value
        # after the epoch's final scheduler step; stops on max_epochs
        gen_agent('agent_trainloop_adaptive.pt', seed=23, epochs=2, train_loop=True, overrides={'max_epochs': 2, 'save_best_after': 1})
    if 'resume' in which:
        gen_resume()
    if 'sched' in which:
        # schedule_type 'standard' (what the shipped mjlab configs use): one adaptive-KL scheduler step per mini-epoch on the mean KL;
        # 4 minibatches per mini-epoch, a learning rate high enough that the schedule moves both ways
        gen_agent('agent_sched_standard.pt', N=16, H=8, mb=32, seed=15, epochs=3,
                  overrides={'schedule_type': 'standard', 'learning_rate': 3e-3, 'kl_threshold': 0.0015})
    if 'tcshape' in which:
        # three hidden layers and an observation width that is a multiple of 4: the shape class of the tcgen05 path's host logic
        # (per-minibatch obs moments precomputed once per epoch, merged in the optimiser tail); masked autoreset on top
        gen_agent('agent_tcshape.pt', N=8, H=8, D=8, A=3, units=(16, 12, 8), mb=32, seed=10, autoreset='next_step')
    if 'cv' in which:
        gen_agent_cv()
    if 'rmsadv' in which:
        gen_rms_adv()
        gen_agent('agent_rmsadv.pt', autoreset='next_step', seed=7, overrides={'normalize_rms_advantage': True, 'adv_rms_momentum': 0.5})
    if 'discrete' in which:
        gen_agent_discrete('agent_discrete.pt')                                   # configs/ppo_cartpole.yaml shape: separate MLP [32,32], 2 actions
        gen_agent_discrete('agent_multidiscrete.pt', K=[3, 4], D=8, units=(16, 8), separate=True, use_action_masks=True, seed=13,
                           overrides={'normalize_input': True, 'normalize_value': True, 'value_bootstrap': True})
        gen_agent_discrete('agent_discrete_masked.pt', K=5, D=7, units=(16, 8), separate=False, use_action_masks=True, autoreset='next_step',
                           seed=12, overrides={'normalize_input': True, 'normalize_value': True, 'lr_schedule': 'adaptive',
                                               'kl_threshold': 0.002, 'value_bootstrap': True})
