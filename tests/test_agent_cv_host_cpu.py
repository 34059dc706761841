"""Host logic of the central-value agent (rl_games_b200.agent_cv.A2CAgentCV, not yet run on hardware) on CPU against the reference's
golden run with central_value_config (tests/golden/agent_cv.pt), with torch stand-ins for every kernel (tests/_torch_ops.py)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_agent_host_cpu import _CudaLookingStr, _Event, _Stream  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


class _Env:
    def __init__(self, g):
        self.g, self.i = g, 0

    def _o(self, j):
        return {'obs': self.g['obs_tape'][j].clone(), 'states': self.g['state_tape'][j].clone()}

    def reset(self):
        self.i = 0
        return self._o(0)

    def step(self, actions):
        g = self.g
        rew = -(actions * actions).sum(-1) * 0.1
        self.i += 1
        j = self.i % g['obs_tape'].shape[0]
        return self._o(j), rew, g['done_tape'][j].clone(), {'time_outs': g['timeout_tape'][j].clone()}

    def get_env_info(self):
        from rl_games_b200.common import Box
        info = {'observation_space': Box(-np.inf, np.inf, (self.g['D'],)), 'action_space': Box(-1.0, 1.0, (self.g['A'],)),
                'state_space': Box(-np.inf, np.inf, (self.g['S'],))}
        if self.g['autoreset'] != 'same_step':
            info['autoreset_mode'] = self.g['autoreset']
        return info

    def get_env_state(self):
        return None

    def set_env_state(self, s):
        pass


def _build_cv(monkeypatch, tmp_path, g, over=None):
    import _torch_ops
    from rl_games_b200.agent_cv import A2CAgentCV
    from rl_games_b200.runner import Runner
    _torch_ops.install_continuous(monkeypatch)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    monkeypatch.setattr(torch.cuda, 'Event', _Event)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a: _Stream())
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self: self)
    cfgk = g['config']
    env = _Env(g)
    cv_cfg = dict(g['cv_config'])
    cv_cfg['network'] = {'name': 'actor_critic', 'central_value': True,
                         'mlp': {'units': g['cv_units'], 'activation': 'elu', 'initializer': {'name': 'default'}}}
    config = {k: v for k, v in cfgk.items() if k not in ('device', 'torch_compile')}
    config.update({'device': _CudaLookingStr('cpu'), 'env_info': env.get_env_info(), 'vec_env': env, 'reward_shaper': {'scale_value': 1.0},
                   'mixed_precision': False, 'b200_cuda_graph': False, 'train_dir': str(tmp_path), 'lr_schedule': cfgk.get('lr_schedule', None),
                   'central_value_config': cv_cfg})
    config.update(over or {})
    network = {'name': 'actor_critic', 'separate': False,
               'space': {'continuous': {'mu_activation': 'None', 'sigma_activation': 'None', 'mu_init': {'name': 'default'},
                                        'sigma_init': {'name': 'const_initializer', 'val': 0}, 'fixed_sigma': True}},
               'mlp': {'units': g['units'], 'activation': 'elu', 'initializer': {'name': 'default'}}}
    r = Runner()
    r.load({'params': {'seed': 7, 'algo': {'name': 'a2c_continuous'}, 'model': {'name': 'continuous_a2c_logstd'}, 'network': network,
                       'config': config}})
    r.params['config']['vec_env'] = env
    agent = r.algo_factory.create(r.algo_name, base_name='parity', params=r.params)
    assert isinstance(agent, A2CAgentCV)
    agent.model.load_state_dict(g['init_state'], strict=False)
    cv = agent.central_value_net
    assert cv.param_names() == g['cv_param_order']
    cv.load_state_dict(g['cv_init_state'])
    agent.init_tensors()
    agent.obs = agent.env_reset()
    return agent


@pytest.mark.parametrize('host_kernels', [False, True], ids=['torch-stand-ins', 'kernel-thread-bodies-on-host'])
def test_central_value_agent_host_logic_matches_reference_golden(host_kernels, monkeypatch, tmp_path):
    """host_kernels: the critic's loss runs the value-loss kernel's OWN per-thread body (csrc/critic.cu, compiled for the host) over the
    agent's arena, and the scheduler step is the optimiser kernels' own function"""
    from oracle import ppo_oracle as O
    g = torch.load(os.path.join(GOLDEN, 'agent_cv.pt'), weights_only=False)
    agent = _build_cv(monkeypatch, tmp_path, g)
    if host_kernels:
        import _torch_ops
        _torch_ops.install_host_kernels(monkeypatch)
    cv = agent.central_value_net
    cv_scalars = []
    cv.writter = type('W', (), {'add_scalar': lambda self, tag, v, step=None: cv_scalars.append((tag, float(v), step))})()
    fl = O.swap_and_flatten01
    flat_noise = g['noise'].reshape(-1, g['N'], g['A'])          # H draws per epoch (the last-value forward goes through the critic)
    for ep, ref in enumerate(g['epochs_out']):
        agent.epoch_num += 1
        agent.train_epoch(noise=flat_noise[ep * g['H']:(ep + 1) * g['H']])
        ds = ref['dataset']
        torch.testing.assert_close(agent.values.unsqueeze(2), ref['mb_values'], rtol=1e-4, atol=1e-5)          # critic values
        torch.testing.assert_close(agent.rewards.unsqueeze(2), ref['mb_rewards'], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(fl(agent.advs_n), ds['advantages'], rtol=2e-4, atol=2e-5)
        torch.testing.assert_close(fl(agent.returns_n.unsqueeze(2)), ds['returns'], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(fl(agent.old_values_n.unsqueeze(2)), ds['old_values'], rtol=1e-4, atol=1e-5)
        st = agent.last_stats
        torch.testing.assert_close(st[:, 0], ref['a_losses'], rtol=2e-3, atol=2e-6)
        torch.testing.assert_close(st[:, 1], ref['c_losses'], rtol=2e-3, atol=2e-6)
        assert agent.last_lr == pytest.approx(ref['last_lr'], rel=1e-12) and cv.lr == pytest.approx(ref['cv_lr'], rel=1e-12)
        sd = agent.model.state_dict()
        for k in O.param_names(len(g['units'])):
            torch.testing.assert_close(sd[k], ref['state'][k], rtol=1e-3, atol=2e-5, msg=lambda m: k + ': ' + m)
        csd = cv.state_dict()
        for k in g['cv_param_order']:
            torch.testing.assert_close(csd[k], ref['cv_state'][k], rtol=1e-3, atol=2e-5, msg=lambda m: 'cv ' + k + ': ' + m)
        assert int(csd['running_mean_std.count']) == int(ref['cv_state']['running_mean_std.count'])
        assert int(csd['value_mean_std.count']) == int(ref['cv_state']['value_mean_std.count'])
        torch.testing.assert_close(csd['value_mean_std.running_var'], ref['cv_state']['value_mean_std.running_var'].reshape(-1), rtol=1e-5, atol=1e-7)
        assert int(sd['value_mean_std.count']) == 1          # the actor model's own value normaliser never moves
    # the critic's own summary scalars (central_value.py:270-272), tag / step / value
    assert [(t, s_) for t, _, s_ in cv_scalars] == [(t, s_) for t, _, s_ in g['cv_scalars']]
    for (t, v, _), (_, rv, _) in zip(cv_scalars, g['cv_scalars']):
        assert v == pytest.approx(rv, rel=2e-3), t
    ck = agent.get_full_state_weights()
    assert [k for k in ck['assymetric_vf_nets'] if 'a2c_network' in k] == ['model.' + k for k in g['cv_param_order']]
    for i, mref in enumerate(g['epochs_out'][-1]['cv_adam_exp_avg']):
        torch.testing.assert_close(ck['assymetric_vf_optimizer']['state'][i]['exp_avg'].reshape(mref.shape), mref, rtol=1e-3, atol=1e-7)


def test_load_critic_only_restores_the_critic_and_leaves_the_actor(monkeypatch, tmp_path):
    """torch_runner.py:43-50 `load_critic_only` -> a2c_continuous.py:90-92 restore_central_value_function -> a2c_common.py:885-887: the
    critic's weights / normalisers come from the checkpoint, the actor and both optimisers are untouched; an agent without a central value
    refuses like the reference's Runner does"""
    from rl_games_b200 import runner as R
    g = torch.load(os.path.join(GOLDEN, 'agent_cv.pt'), weights_only=False)
    a = _build_cv(monkeypatch, tmp_path, g)
    flat_noise = g['noise'].reshape(-1, g['N'], g['A'])
    a.epoch_num += 1
    a.train_epoch(noise=flat_noise[:g['H']])
    a.save(str(tmp_path / 'ck'))
    b = _build_cv(monkeypatch, tmp_path, g)
    actor0, exp0 = b.model.flat.clone(), b.central_value_net.exp_avg.clone()
    assert not torch.equal(b.central_value_net.flat, a.central_value_net.flat)
    R._restore(b, {'train': True, 'checkpoint': str(tmp_path / 'ck.pth'), 'load_critic_only': True})
    for k, v in a.central_value_net.state_dict().items():
        assert torch.equal(b.central_value_net.state_dict()[k], v), k
    assert torch.equal(b.model.flat, actor0) and torch.equal(b.central_value_net.exp_avg, exp0) and b.epoch_num == 0
    R._restore(b, {'train': True, 'checkpoint': str(tmp_path / 'ck.pth')})          # the full restore still works on the same file
    assert torch.equal(b.model.flat, a.model.flat) and b.epoch_num == a.epoch_num
    plain = type('A', (), {'has_central_value': False})()
    with pytest.raises(ValueError, match='asymmetric actor critic'):
        R._restore(plain, {'train': True, 'checkpoint': 'x.pth', 'load_critic_only': True})
