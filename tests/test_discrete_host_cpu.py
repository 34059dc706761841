"""Host logic of rl_games_b200.agent_discrete.DiscreteA2CAgent on CPU: every op it calls is replaced by a torch stand-in
(tests/_torch_ops.py) that restates the kernel's contract, and the agent is run against the reference's golden discrete runs.
This checks arena addressing, flat-parameter offsets (incl. the split heads of a separate critic), call order, the per-mini-epoch
scheduler, meters and checkpoint keys -- NOT the CUDA kernels (tests/test_discrete_gpu.py does that on a GPU)."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


class _Env:
    def __init__(self, g):
        self.g, self.K, self.i = g, g['K'], 0

    def reset(self):
        self.i = 0
        return self.g['obs_tape'][0].clone()

    def get_action_masks(self):
        return self.g['mask_tape'][self.i % self.g['mask_tape'].shape[0]]

    def step(self, actions):
        g = self.g
        j0 = self.i % g['obs_tape'].shape[0]
        if isinstance(self.K, (list, tuple)):
            off, rew = 0, 0.0
            for j, k in enumerate(self.K):
                rew = rew + (actions[:, j].long() == g['obs_tape'][j0][:, off:off + k].argmax(dim=-1)).float() / len(self.K)
                off += k
        else:
            rew = (actions.long() == g['obs_tape'][j0][:, :self.K].argmax(dim=-1)).float()
        self.i += 1
        j = self.i % g['obs_tape'].shape[0]
        return g['obs_tape'][j].clone(), rew, g['done_tape'][j].clone(), {'time_outs': g['timeout_tape'][j].clone()}

    def get_env_info(self):
        from rl_games_b200.common import Box, Discrete, Tuple
        space = Tuple([Discrete(k) for k in self.K]) if isinstance(self.K, (list, tuple)) else Discrete(self.K)
        info = {'observation_space': Box(-np.inf, np.inf, (self.g['obs_tape'].shape[-1],)), 'action_space': space}
        if self.g['autoreset'] != 'same_step':
            info['autoreset_mode'] = self.g['autoreset']
        return info


def _build(monkeypatch, tmp_path, g, stand_ins=True, over=None):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _torch_ops
    from rl_games_b200 import agent_discrete
    from rl_games_b200.runner import Runner
    if stand_ins:
        _torch_ops.install(monkeypatch)
    monkeypatch.setattr(agent_discrete.DiscreteA2CAgent, '_require_cuda', lambda self: None)
    monkeypatch.setattr(agent_discrete.DiscreteA2CAgent, '_sync', staticmethod(lambda: None))
    cfgk = g['config']
    env = _Env(g)
    config = {k: v for k, v in cfgk.items() if k not in ('device', 'torch_compile')}
    config.update({'device': 'cpu', 'env_info': env.get_env_info(), 'vec_env': env, 'reward_shaper': {'scale_value': 0.1},
                   'train_dir': str(tmp_path), 'lr_schedule': cfgk.get('lr_schedule', None)})
    config.update(over or {})
    multi = isinstance(g['K'], (list, tuple))
    network = {'name': 'actor_critic', 'separate': g['separate'], 'space': {'multi_discrete' if multi else 'discrete': None},
               'mlp': {'units': g['units'], 'activation': 'relu', 'initializer': {'name': 'default'}}}
    r = Runner()
    r.load({'params': {'seed': 1, 'algo': {'name': 'a2c_discrete'}, 'model': {'name': 'multi_discrete_a2c' if multi else 'discrete_a2c'},
                       'network': network, 'config': config}})
    r.params['config']['vec_env'] = env
    agent = r.algo_factory.create(r.algo_name, base_name='parity', params=r.params)
    agent.model.load_state_dict(g['init_state'], strict=False)
    assert agent.model.param_names() == g['param_order']
    agent.init_tensors()
    agent.obs = agent.env_reset()
    return agent


@pytest.mark.parametrize('host_kernels', [False, True], ids=['torch-stand-ins', 'kernel-thread-bodies-on-host'])
@pytest.mark.parametrize('name', ['agent_discrete.pt', 'agent_discrete_masked.pt', 'agent_multidiscrete.pt'])
def test_discrete_agent_host_logic_matches_reference_golden(name, host_kernels, monkeypatch, tmp_path):
    """host_kernels: the two categorical ops run the kernels' OWN per-thread bodies (csrc/discrete.cu, __host__ __device__, compiled for the
    host) over the agent's arena instead of the torch stand-ins -- the kernels' indexing and arithmetic against the reference's golden runs"""
    g = torch.load(os.path.join(GOLDEN, name), weights_only=False)
    cfgk = g['config']
    agent = _build(monkeypatch, tmp_path, g)
    if host_kernels:
        import _torch_ops
        _torch_ops.install_host_kernels(monkeypatch)
    fl = lambda t: t.transpose(0, 1).reshape(-1, *t.shape[2:])    # noqa: E731
    for ep, ref in enumerate(g['epochs_out']):
        agent.epoch_num += 1
        res = agent.train_epoch(u=g['u'][ep])
        ds = ref['dataset']
        assert torch.equal(agent.actions, ref['mb_actions'])
        if g['use_action_masks']:
            assert torch.equal(fl(agent.action_masks).bool(), ds['action_masks'])
        if ds.get('rnn_masks') is not None:
            assert torch.equal(fl(agent.valid), ds['rnn_masks'])
        torch.testing.assert_close(agent.rewards.unsqueeze(2), ref['mb_rewards'], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(agent.values.unsqueeze(2), ref['mb_values'], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(fl(agent.advs_n), ds['advantages'], rtol=2e-4, atol=2e-5)
        torch.testing.assert_close(fl(agent.returns_n).unsqueeze(1), ds['returns'], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(fl(agent.neglogpacs), ds['old_logp_actions'], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(torch.stack(res[4]), ref['a_losses'], rtol=2e-3, atol=2e-6)
        torch.testing.assert_close(torch.stack(res[5]), ref['c_losses'], rtol=2e-3, atol=2e-6)
        torch.testing.assert_close(torch.stack(res[6]), ref['entropies'], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(torch.stack(res[7]), ref['kls'], rtol=5e-3, atol=1e-8)
        assert agent.last_lr == pytest.approx(ref['last_lr'], rel=1e-12)
        sd = agent.model.state_dict()
        for k in g['param_order']:
            torch.testing.assert_close(sd[k], ref['state'][k], rtol=1e-3, atol=2e-5, msg=lambda m: k + ': ' + m)
        if cfgk.get('normalize_input'):
            assert int(sd['running_mean_std.count']) == int(ref['state']['running_mean_std.count'])
            torch.testing.assert_close(sd['value_mean_std.running_var'], ref['state']['value_mean_std.running_var'].reshape(-1), rtol=1e-5, atol=1e-7)
        assert agent.game_rewards.current_size == ref['game_rewards_size']
        torch.testing.assert_close(agent.game_rewards.mean, ref['game_rewards_mean'].reshape(-1), rtol=1e-4, atol=1e-5)
    # checkpoint keys / Adam state in the reference's order
    ck = agent.get_full_state_weights()
    assert [k for k in ck['model'] if k.startswith('a2c_network')] == g['param_order']
    ref_m = g['epochs_out'][-1]['adam_exp_avg']
    for i, m in enumerate(ref_m):
        torch.testing.assert_close(ck['optimizer']['state'][i]['exp_avg'].reshape(m.shape), m, rtol=1e-3, atol=1e-7)


def test_discrete_train_loop_and_checkpoint_roundtrip(monkeypatch, tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _torch_ops
    from rl_games_b200 import agent_discrete
    from rl_games_b200.runner import Runner
    _torch_ops.install(monkeypatch)
    monkeypatch.setattr(agent_discrete.DiscreteA2CAgent, '_require_cuda', lambda self: None)
    monkeypatch.setattr(agent_discrete.DiscreteA2CAgent, '_sync', staticmethod(lambda: None))
    g = torch.load(os.path.join(GOLDEN, 'agent_discrete_masked.pt'), weights_only=False)

    def build():
        env = _Env(g)
        config = {k: v for k, v in g['config'].items() if k not in ('device', 'torch_compile')}
        config.update({'device': 'cpu', 'env_info': env.get_env_info(), 'vec_env': env, 'reward_shaper': {'scale_value': 0.1},
                       'train_dir': str(tmp_path), 'lr_schedule': g['config'].get('lr_schedule', None),
                       'max_epochs': 2, 'print_stats': False, 'name': 'dloop'})
        network = {'name': 'actor_critic', 'separate': g['separate'], 'space': {'discrete': None},
                   'mlp': {'units': g['units'], 'activation': 'relu', 'initializer': {'name': 'default'}}}
        r = Runner()
        r.load({'params': {'seed': 1, 'algo': {'name': 'a2c_discrete'}, 'model': {'name': 'discrete_a2c'}, 'network': network, 'config': config}})
        r.params['config']['vec_env'] = env
        return r.algo_factory.create(r.algo_name, base_name='parity', params=r.params)
    a = build()
    last, epochs = a.train()
    assert epochs == 2 and a.frame == 2 * g['N'] * g['H'] and torch.isfinite(a.model.flat).all()
    fn = str(tmp_path / 'dck')
    a.save(fn)
    b = build()
    b.restore(fn + '.pth')
    assert torch.equal(b.model.flat, a.model.flat) and torch.equal(b.model.exp_avg_sq, a.model.exp_avg_sq)
    # reference semantics (a2c_common.py:852-866): the optimizer returns with its lr, last_lr is not part of a restore
    assert b.epoch_num == a.epoch_num and float(b.opt_state[0]) == a.last_lr and b.last_lr == g['config']['learning_rate']


def test_discrete_train_loop_matches_the_reference_outer_loop(monkeypatch, tmp_path):
    """agent.train() against DiscreteA2CBase.train of the reference (a2c_common.py:1361-1470) on the same tapes and uniform draws:
    frame / epoch accounting, per-mini-epoch adaptive schedule, every summary scalar (tag, step, value), periodic / best / final
    checkpoint names, return value"""
    import re
    g = torch.load(os.path.join(GOLDEN, 'agent_discrete_trainloop.pt'), weights_only=False)
    ref = g['train_loop']
    agent = _build(monkeypatch, tmp_path, g)
    orig, seen = agent.train_epoch, []

    def with_draws():
        seen.append((agent.frame, agent.epoch_num))
        return orig(u=g['u'][agent.epoch_num - 1])
    agent.train_epoch = with_draws
    scalars = []
    agent.writer = type('W', (), {'add_scalar': lambda self, tag, v, step=None: scalars.append((tag, float(v), step))})()
    agent.algo_observer.writer = agent.writer
    ret = agent.train()
    assert (float(ret[0]), int(ret[1])) == pytest.approx(ref['return'], rel=1e-5)
    assert agent.frame == ref['frame'] and agent.epoch_num == ref['epoch_num']
    assert seen == [(e['frame_before'], e['epoch_num']) for e in g['epochs_out']]
    assert agent.last_lr == pytest.approx(g['epochs_out'][-1]['last_lr'], rel=1e-12)
    assert [(t, s_) for t, _, s_ in scalars if not t.endswith('/time')] == [(t, s_) for t, _, s_ in ref['scalars'] if not t.endswith('/time')]
    assert [t for t, _, _ in scalars] == [t for t, _, _ in ref['scalars']]
    for (t, v, _), (_, rv, _) in zip(scalars, ref['scalars']):
        if not t.startswith('performance/'):
            assert v == pytest.approx(rv, rel=5e-3, abs=2e-6), t
    sd = agent.model.state_dict()
    for k in g['param_order']:
        torch.testing.assert_close(sd[k], g['epochs_out'][-1]['state'][k], rtol=1e-3, atol=2e-5, msg=lambda m: k + ': ' + m)
    num = re.compile(r'_rew__?(-?[0-9.]+)')
    got, want = sorted(os.listdir(agent.nn_dir)), sorted(ref['saved'])
    assert [num.sub('_rew_#', n) for n in got] == [num.sub('_rew_#', n) for n in want] and len(got) == 3
    for a, b in zip(got, want):
        ma, mb_ = num.search(a), num.search(b)
        if ma:
            assert float(ma.group(1).rstrip('.')) == pytest.approx(float(mb_.group(1).rstrip('.')), rel=1e-5)
            assert len(ma.group(1)) <= len(mb_.group(1)) + 1
