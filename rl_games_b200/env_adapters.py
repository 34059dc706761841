"""Tensor-env adapters: the step on the ENV side of the hot path.

``ManagerBasedEnvAdapter`` puts an Isaac-Lab-style manager-based environment (mjlab's ``ManagerBasedRlEnv`` and plugins with the
same API: ``reset() -> (groups, extras)``, ``step(a) -> (groups, reward, terminated, truncated, extras)``, torch tensors on the
simulation device) behind the IVecEnv contract the agent consumes.  Behaviour follows the reference's wrapper
(envs/mjlab_vecenv.py:16-170): policy observations come from the group called 'actor' (mjlab) or 'policy' (Isaac-Lab plugins), a
'critic' group becomes the central-value ``states`` (+ ``use_global_observations``), ``time_outs`` are the truncations, and the
env's per-burst episode metrics are copied -- not aliased -- into ``infos['episode']``.  What differs is on the ingestion side and
invisible to the maths: dones stay a bool tensor (the post-step kernel reads bool / uint8 / fp32 flags directly, so the reference's
``.float()`` cast launch is not needed) and nothing here synchronises the device.

``MjlabVecEnv`` (vecenv type 'MJLAB', common/vecenv.py:412-415) builds the mjlab backend from ``env_config``; mjlab / warp are not part
of this image, so that constructor only runs where they are installed -- the adapter logic is tested against a fake backend
(tests/test_env_adapters_cpu.py).
"""
import os

import numpy as np
import torch

from .common import IVecEnv, Box, register_vecenv

_POLICY_GROUPS = ('actor', 'policy')        # mjlab's own tasks / Isaac-Lab-style task plugins
_PRIVILEGED_GROUP = 'critic'


class ManagerBasedEnvAdapter(IVecEnv):
    def __init__(self, env, device=None):
        self.env, self.device = env, device
        groups, _ = env.reset()
        self.policy_group = next((k for k in _POLICY_GROUPS if k in groups), None)
        if self.policy_group is None:
            raise KeyError(f"manager-based env returned observation groups {sorted(groups)}; expected 'actor' or 'policy'")
        first = groups[self.policy_group]
        self.num_envs, self.obs_dim = int(first.shape[0]), int(first.shape[-1])
        self.privileged = _PRIVILEGED_GROUP in groups
        self.observation_space = Box(-np.inf, np.inf, (self.obs_dim,))
        self.action_space = Box(-1.0, 1.0, (int(env.action_space.shape[-1]),))
        self.state_space = Box(-np.inf, np.inf, (int(groups[_PRIVILEGED_GROUP].shape[-1]),)) if self.privileged else None

    # ---- observation groups -> what the agent reads
    def _view(self, groups):
        policy = groups[self.policy_group]
        return {'obs': policy, 'states': groups[_PRIVILEGED_GROUP]} if self.privileged else policy

    @staticmethod
    def _refresh_episode_metrics(extras):
        """The env reuses its extras dict (and possibly the metric tensors) across steps and reports an empty 'log' between reset bursts
        (mjlab_vecenv.py:100-113): rebuild the entry on every step, cloning tensors so that later in-place writes cannot leak in."""
        burst = extras.get('log') or None
        if burst is None:
            extras.pop('episode', None)
            return
        extras['episode'] = {name: (value.clone() if torch.is_tensor(value) else value) for name, value in burst.items()}

    # ---- IVecEnv
    def step(self, actions):
        groups, reward, terminated, truncated, extras = self.env.step(actions)
        extras['time_outs'] = truncated
        self._refresh_episode_metrics(extras)
        return self._view(groups), reward, torch.logical_or(terminated, truncated), extras

    def reset(self):
        # like the reference adapter (envs/mjlab_vecenv.py:128-132): every reset() resets the env again, the constructor's probing reset
        # included, so the env's RNG stream is the reference's for the same seed
        groups, _ = self.env.reset()
        return self._view(groups)

    def get_number_of_agents(self):
        return 1

    def get_env_info(self):
        spaces = {'observation_space': self.observation_space, 'action_space': self.action_space}
        if self.privileged:
            spaces.update(state_space=self.state_space, use_global_observations=True)
        return spaces

    def seed(self, seed):
        """the backend was seeded at construction (env_config seed)"""

    def close(self):
        self.env.close()


def resolve_sim_device(device, local_rank=None):
    """Under torchrun a bare 'cuda' would put every rank's simulation on cuda:0 (mjlab_vecenv.py:31-36)."""
    rank = os.getenv('LOCAL_RANK') if local_rank is None else local_rank
    return f'cuda:{rank}' if (device == 'cuda' and rank is not None) else device


def apply_velocity_stage_steps(cfg, stage_steps):
    """Optional override of a step-scheduled command curriculum: one switch point (in env steps) per stage (mjlab_vecenv.py:48-60)."""
    stages = cfg.curriculum['command_vel'].params['velocity_stages']
    if len(stages) != len(stage_steps):
        raise ValueError(f"velocity_stage_steps has {len(stage_steps)} entries, task schedule has {len(stages)} stages")
    for i, at in enumerate(stage_steps):
        stages[i]['step'] = int(at)


class MjlabVecEnv(ManagerBasedEnvAdapter):
    """``env_config: {task_name, device, seed, velocity_stage_steps}`` (envs/mjlab_vecenv.py:23-71)"""
    OPTIONS = ('task_name', 'device', 'seed', 'velocity_stage_steps')

    def __init__(self, config_name, num_actors, **env_config):
        try:
            import warp
            from mjlab.envs.manager_based_rl_env import ManagerBasedRlEnv
            from mjlab.tasks.registry import load_env_cfg
        except ImportError as e:
            raise ImportError("vecenv_type MJLAB needs the 'mjlab' and 'warp' packages (not part of this image); any env with "
                              "the manager-based API can be wrapped directly with ManagerBasedEnvAdapter(env)") from e
        warp.init()
        unknown = sorted(set(env_config) - set(self.OPTIONS))
        if unknown:     # a silently dropped key means the run trains under other settings than the config claims
            print(f'WARNING: mjlab wrapper ignoring unknown env_config keys: {unknown}')
        device = resolve_sim_device(env_config.get('device', 'cuda'))
        cfg = load_env_cfg(env_config.get('task_name', config_name))
        cfg.scene.num_envs = num_actors
        if env_config.get('seed') is not None:          # the runner's seed, already rank-offset: reproducible randomisation
            cfg.seed = int(env_config['seed'])
        if env_config.get('velocity_stage_steps') is not None:
            apply_velocity_stage_steps(cfg, env_config['velocity_stage_steps'])
        super().__init__(ManagerBasedRlEnv(cfg, device=device), device=device)


register_vecenv('MJLAB', lambda config_name, num_actors, **kwargs: MjlabVecEnv(config_name, num_actors, **kwargs))
