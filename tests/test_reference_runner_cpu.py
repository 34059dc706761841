"""INTEGRATION.md section 1 as a test: the B200 agent registered into the REFERENCE's own `rl_games.torch_runner.Runner`
(torch_runner.py:98-354), a stock YAML of the reference loaded unchanged (configs/mujoco/ant_envpool.yaml), `runner.run(...)`.

No GPU: every C-ABI call is answered by the header-checking recorder of test_abi_calls_cpu (argument count / ctypes types against
include/b200rl.h, returns success), so what is tested is the boundary -- the reference Runner's config injection (DefaultRewardsShaper
object, features.observer, seed handling), its run_train sequence (_restore, _override_sigma, torch.compile of agent.model: the stock
YAML leaves torch_compile at its default True) and the agent's train() loop with the reference's own DefaultAlgoObserver.
The reference comes from /root/reference (build container) or the vendored oracle/_ref; absent both the test is skipped."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import test_abi_calls_cpu as ABI  # noqa: E402
import test_agent_host_cpu as H  # noqa: E402

ROOT = os.path.dirname(HERE)


def _reference_root():
    for r in ('/root/reference', os.path.join(ROOT, 'oracle', '_ref')):
        if os.path.isdir(os.path.join(r, 'rl_games')):
            return r
    return None


REF = _reference_root()
pytestmark = pytest.mark.skipif(REF is None, reason='reference rl_games not present (neither /root/reference nor oracle/_ref)')


class _GymEnv:
    """what config['env_info'] / config['vec_env'] injection needs (a2c_common.py:236-241); spaces are the stub's gymnasium Boxes,
    i.e. what a reference user would pass"""

    def __init__(self, N, D, A):
        self.N, self.D, self.A = N, D, A

    def reset(self):
        return torch.zeros(self.N, self.D)

    def step(self, actions):
        assert tuple(actions.shape) == (self.N, self.A)
        z8 = torch.zeros(self.N, dtype=torch.uint8)
        return torch.zeros(self.N, self.D), torch.zeros(self.N), z8, {'time_outs': z8}

    def get_env_info(self):
        import gymnasium as gym
        import numpy as np
        return {'observation_space': gym.spaces.Box(-np.inf, np.inf, (self.D,), np.float32),
                'action_space': gym.spaces.Box(-1.0, 1.0, (self.A,), np.float32)}

    def set_train_info(self, *a, **kw):
        pass

    def get_env_state(self):
        return None

    def set_env_state(self, s):
        pass


@pytest.mark.parametrize('obs_dim', [27, 105])      # Ant without / with contact forces: resident-W1 and wide-observation tcgen05 kernels
def test_stock_yaml_trains_through_the_reference_runner(obs_dim, monkeypatch, tmp_path, capsys):
    import yaml
    for p in (os.path.join(HERE, 'golden', '_stubs'), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from rl_games.torch_runner import Runner
    from rl_games.common.algo_observer import DefaultAlgoObserver
    from rl_games_b200.agent import A2CAgent as B200A2CAgent

    rec = ABI._patch(monkeypatch)
    cfg = yaml.safe_load(open(os.path.join(REF, 'rl_games', 'configs', 'mujoco', 'ant_envpool.yaml')))
    c = cfg['params']['config']
    assert c['mixed_precision'] is True and 'torch_compile' not in c          # the YAML itself is untouched below except for:
    env = _GymEnv(c['num_actors'], obs_dim, 8)
    c.update({'env_info': env.get_env_info(), 'vec_env': env, 'device': H._CudaLookingStr('cpu'), 'max_epochs': 2,
              'train_dir': str(tmp_path), 'b200_cuda_graph': False})

    runner = Runner()
    runner.algo_factory.register_builder('a2c_continuous', lambda **kwargs: B200A2CAgent(**kwargs))
    runner.load(cfg)
    runner.params['config']['vec_env'] = env          # Runner.load deep-copies the config
    made = []
    orig_create = runner.algo_factory.create

    def create(name, **kw):
        a = orig_create(name, **kw)
        made.append(a)
        return a
    monkeypatch.setattr(runner.algo_factory, 'create', create)
    runner.run({'train': True, 'play': False, 'checkpoint': None, 'sigma': None})

    agent, = made
    assert isinstance(agent, B200A2CAgent) and isinstance(agent.algo_observer, DefaultAlgoObserver)
    assert type(agent.rewards_shaper).__module__.startswith('rl_games.')          # the reference's DefaultRewardsShaper object, accepted as is
    assert hasattr(agent.model, 'load_optimizer_state_dict')                       # torch.compile(agent.model) was ignored, not installed
    assert 'torch.compile of agent.model ignored' in capsys.readouterr().out
    assert agent.epoch_num == 2 and agent.frame == 2 * c['num_actors'] * c['horizon_length']
    assert agent.use_tc and agent.tc_wide == (obs_dim > 64)
    n_upd = 2 * c['mini_epochs'] * (c['num_actors'] * c['horizon_length'] // c['minibatch_size'])
    assert rec.calls['b200rl_tc_mlp_fwd_train'] == n_upd and rec.calls['b200rl_tc_mlp_bwd'] == n_upd
    assert rec.calls['b200rl_tc_mlp_fwd_rollout'] == 2 * (c['horizon_length'] + 1)
    assert rec.calls['b200rl_gae_fused_f32'] == 2
    ckpts = os.listdir(agent.nn_dir)
    assert any(f.startswith('last_Ant-v5_envpool_ep_2') and f.endswith('.pth') for f in ckpts), ckpts
    sd = torch.load(os.path.join(agent.nn_dir, ckpts[0]), weights_only=False)
    assert {'model', 'optimizer', 'epoch', 'frame', 'last_mean_rewards'} <= set(sd) and 'a2c_network.mu.weight' in sd['model']
