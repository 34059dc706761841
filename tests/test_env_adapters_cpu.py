"""Env-side adapters on CPU: the manager-based tensor-env adapter against a fake backend with mjlab's API, the
device-resident episode-metric observer against the reference's concatenate-and-mean rule, and the config keys that select them
(envs/mjlab_vecenv.py, common/algo_observer.py:95-156, torch_runner.py:153-165)."""
import types

import pytest
import torch

from rl_games_b200 import env_adapters as EA
from rl_games_b200.common import IsaacAlgoObserver, DefaultAlgoObserver, vecenv_config, configurations


class _FakeManagerEnv:
    """reset() -> (obs_dict, extras); step(a) -> (obs_dict, reward, terminated, truncated, extras); extras dict reused across steps"""

    def __init__(self, n=6, d=5, s=9, a=3, critic=True, key='actor'):
        self.n, self.d, self.s, self.key, self.critic = n, d, s, key, critic
        self.action_space = types.SimpleNamespace(shape=(n, a))
        self.extras = {}
        self.t = 0
        self.g = torch.Generator().manual_seed(0)
        self.log_buf = torch.zeros(2)
        self.closed = False

    def _obs(self):
        o = {self.key: torch.randn(self.n, self.d, generator=self.g)}
        if self.critic:
            o['critic'] = torch.randn(self.n, self.s, generator=self.g)
        return o

    def reset(self):
        self.t = 0
        return self._obs(), self.extras

    def step(self, actions):
        self.t += 1
        term = torch.rand(self.n, generator=self.g) < 0.2
        trunc = (torch.rand(self.n, generator=self.g) < 0.2) & ~term
        self.log_buf.fill_(float(self.t))           # the env reuses the tensor it reports
        self.extras['log'] = {'len': self.log_buf, 'plain': 2.0} if self.t % 2 == 1 else {}
        return self._obs(), torch.randn(self.n, generator=self.g), term, trunc, self.extras

    def close(self):
        self.closed = True


@pytest.mark.parametrize('critic,key', [(True, 'actor'), (False, 'policy')])
def test_manager_based_adapter_follows_reference_wrapper(critic, key):
    env = EA.ManagerBasedEnvAdapter(_FakeManagerEnv(critic=critic, key=key))
    info = env.get_env_info()
    assert info['observation_space'].shape == (5,) and info['action_space'].shape == (3,)
    assert (info['action_space'].low == -1).all() and (info['action_space'].high == 1).all()
    assert ('state_space' in info) == critic and info.get('use_global_observations', False) == critic
    if critic:
        assert info['state_space'].shape == (9,)
    assert env.get_number_of_agents() == 1 and env.num_envs == 6
    o = env.reset()
    assert (set(o) == {'obs', 'states'}) if critic else (o.shape == (6, 5))
    kept = []
    for t in range(1, 5):
        obs, rew, done, infos = env.step(torch.zeros(6, 3))
        assert done.dtype == torch.bool and done.shape == (6,) and rew.shape == (6,)
        assert infos['time_outs'].dtype == torch.bool and not (infos['time_outs'] & ~done).any()      # truncated is a subset of done
        if t % 2 == 1:      # reset burst: metrics copied, not aliased
            assert infos['episode']['plain'] == 2.0
            kept.append((t, infos['episode']['len']))
        else:               # an empty log must drop the stale entry (mjlab_vecenv.py:108-113)
            assert 'episode' not in infos
    for t, v in kept:
        assert float(v[0]) == float(t)          # cloned at the time: later in-place writes by the env do not leak in
    env.close()
    assert env.env.closed


def test_adapter_rejects_unknown_observation_groups_and_mjlab_import_is_loud():
    with pytest.raises(KeyError, match="'actor' or 'policy'"):
        EA.ManagerBasedEnvAdapter(_FakeManagerEnv(key='proprio'))
    assert 'MJLAB' in vecenv_config
    with pytest.raises(ImportError, match='mjlab'):
        vecenv_config['MJLAB']('mjlab_go1_velocity', 4, task_name='Mjlab-Velocity-Flat-Unitree-Go1')


def test_mjlab_config_helpers():
    assert EA.resolve_sim_device('cuda', local_rank='3') == 'cuda:3'
    assert EA.resolve_sim_device('cuda:1', local_rank='3') == 'cuda:1'
    assert EA.resolve_sim_device('cpu', local_rank='3') == 'cpu'
    stages = [{'step': 0}, {'step': 10}, {'step': 20}]
    cfg = types.SimpleNamespace(curriculum={'command_vel': types.SimpleNamespace(params={'velocity_stages': stages})})
    EA.apply_velocity_stage_steps(cfg, [0, 60000, 120000])
    assert [s['step'] for s in stages] == [0, 60000, 120000]
    with pytest.raises(ValueError, match='velocity_stage_steps has 2 entries, task schedule has 3 stages'):
        EA.apply_velocity_stage_steps(cfg, [0, 1])


class _Writer:
    def __init__(self):
        self.rows = []

    def add_scalar(self, tag, v, step):
        self.rows.append((tag, float(v), step))


def _reference_episode_means(bursts):
    """common/algo_observer.py:128-146: per key (union over bursts), concatenate every burst's values, take the mean"""
    out = {}
    for key in sorted(set().union(*bursts)):
        vals = []
        for b in bursts:
            if key in b:
                v = b[key] if isinstance(b[key], torch.Tensor) else torch.Tensor([b[key]])
                vals.append(v.reshape(-1) if v.dim() else v.unsqueeze(0))
        out[key] = torch.cat(vals).mean().item()
    return out


def test_isaac_observer_matches_reference_concat_mean_without_done_indices():
    g = torch.Generator().manual_seed(3)
    bursts = [{'a': torch.randn(4, generator=g), 'b': torch.randn((), generator=g)}, {'a': torch.randn(7, generator=g)},
              {'b': 1.5, 'c': torch.randn(2, generator=g)}]
    obs = IsaacAlgoObserver()
    assert obs.wants_done_indices is False and obs.wants_infos is True
    w = _Writer()
    obs.after_init(types.SimpleNamespace(writer=w, ppo_device='cpu', games_to_track=100))
    with pytest.raises(ValueError, match="expected 'infos' as dict"):
        obs.process_infos([{}], None)
    for b in bursts:
        obs.process_infos({'episode': {k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()}, 'time_outs': torch.zeros(3),
                           'curriculum_level': 2.0}, None)
        obs.process_infos({'time_outs': torch.zeros(3)}, None)       # steps between bursts carry no episode dict
    obs.after_print_stats(1000, 7, 12.5)
    got = {t: v for t, v, s in w.rows if t.startswith('Episode/')}
    ref = _reference_episode_means(bursts)
    assert set(got) == {'Episode/' + k for k in ref}
    for k, v in ref.items():
        assert got['Episode/' + k] == pytest.approx(v, rel=1e-5, abs=1e-6)
    assert all(s == 7 for t, v, s in w.rows if t.startswith('Episode/'))
    # the last infos had no scalar except none -> direct_info reflects the LAST call only (reference resets it every call)
    assert not [r for r in w.rows if r[0].startswith('curriculum_level')]
    n = len(w.rows)
    obs.process_infos({'curriculum_level': 3.0, 'lvl_t': torch.tensor(4.0)}, None)
    obs.after_print_stats(2000, 8, 20.0)        # bursts were cleared at the previous print
    new = w.rows[n:]
    assert ('curriculum_level/frame', 3.0, 2000) in new and ('curriculum_level/iter', 3.0, 8) in new and ('lvl_t/time', 4.0, 20.0) in new
    assert not [r for r in new if r[0].startswith('Episode/')]


def test_runner_selects_observer_and_registers_vecenv_type_from_config():
    from rl_games_b200.runner import Runner
    base = {'seed': 3, 'algo': {'name': 'a2c_continuous'}, 'model': {'name': 'continuous_a2c_logstd'}, 'network': {},
            'config': {'name': 'x', 'env_name': 'mjlab_some_task', 'vecenv_type': 'MJLAB', 'algo_observer': 'isaac',
                       'reward_shaper': {'scale_value': 1.0}}}
    r = Runner()
    r.load({'params': base})
    assert isinstance(r.algo_observer, IsaacAlgoObserver) and r.params['config']['features']['observer'] is r.algo_observer
    assert configurations['mjlab_some_task'] == {'vecenv_type': 'MJLAB'}
    injected = DefaultAlgoObserver()
    r2 = Runner(injected)
    r2.load({'params': base})
    assert r2.algo_observer is injected         # object injection wins over the config key (torch_runner.py:160-164)
