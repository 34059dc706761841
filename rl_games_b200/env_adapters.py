"""Tensor-env adapters: the step on the ENV side of the hot path.

``ManagerBasedEnvAdapter`` puts an Isaac-Lab-style manager-based environment (mjlab's ``ManagerBasedRlEnv`` and
plugins with the same API: ``reset() -> (obs_dict, extras)``, ``step(a) -> (obs_dict, reward, terminated, truncated,
extras)``, all torch tensors on the simulation device) behind the reference's IVecEnv contract, the way
envs/mjlab_vecenv.py:16-170 does -- same observation-group rules ('actor' or 'policy'; 'critic' => central-value
``states`` + ``use_global_observations``), same ``time_outs = truncated``, same per-burst copy of the env's episode log
into ``infos['episode']``.  Differences, all on the ingestion side and invisible to the maths:

* dones stay a bool tensor (``terminated | truncated``): the post-step kernel reads bool / uint8 / fp32 dones and
  time-outs directly (ops.post_step), so the reference's ``done.float()`` cast launch is not needed;
* nothing here synchronises the device; episode-log tensors are cloned on the stream they were produced on.

``MjlabVecEnv`` builds the mjlab backend exactly as the reference wrapper does (vecenv type 'MJLAB',
common/vecenv.py:412-415); mjlab / warp are not part of this image, so that constructor is exercised only where they are
installed -- the adapter logic itself is tested against a fake backend (tests/test_env_adapters_cpu.py).
"""
import os

import numpy as np
import torch

from .common import IVecEnv, Box, register_vecenv


class ManagerBasedEnvAdapter(IVecEnv):
    def __init__(self, env, device=None):
        self.env = env
        self.device = device
        obs, _ = env.reset()
        # mjlab's own tasks call the policy group 'actor'; Isaac-Lab-style plugins call it 'policy' (mjlab_vecenv.py:73-76)
        self.actor_key = 'actor' if 'actor' in obs else 'policy'
        if self.actor_key not in obs:
            raise KeyError(f"manager-based env returned observation groups {sorted(obs)}; expected 'actor' or 'policy'")
        self.num_envs = obs[self.actor_key].shape[0]
        self.obs_dim = obs[self.actor_key].shape[-1]
        self.has_critic_obs = 'critic' in obs
        self.observation_space = Box(-np.inf, np.inf, (self.obs_dim,))
        self.action_space = Box(-1.0, 1.0, (env.action_space.shape[-1],))
        if self.has_critic_obs:
            self.state_space = Box(-np.inf, np.inf, (obs['critic'].shape[-1],))
        self._first_obs = obs

    def _pack(self, obs_dict):
        if self.has_critic_obs:     # asymmetric actor-critic: privileged observations feed the central value net
            return {'obs': obs_dict[self.actor_key], 'states': obs_dict['critic']}
        return obs_dict[self.actor_key]

    @staticmethod
    def _extract_episode_log(info):
        """mjlab_vecenv.py:100-113: the env reuses its extras dict (and may reuse the value tensors) and emits an empty 'log'
        between reset bursts -- refresh every step, clone tensors so nothing aliases env internals."""
        log = info.get('log')
        if log:
            info['episode'] = {k: v.clone() if torch.is_tensor(v) else v for k, v in log.items()}
        else:
            info.pop('episode', None)

    def step(self, actions):
        obs_dict, reward, terminated, truncated, info = self.env.step(actions)
        info['time_outs'] = truncated
        self._extract_episode_log(info)
        return self._pack(obs_dict), reward, terminated | truncated, info

    def reset(self):
        if self._first_obs is not None:     # the constructor already reset the backend once to learn the shapes
            obs, self._first_obs = self._first_obs, None
            return self._pack(obs)
        obs_dict, _ = self.env.reset()
        return self._pack(obs_dict)

    def get_number_of_agents(self):
        return 1

    def get_env_info(self):
        info = {'action_space': self.action_space, 'observation_space': self.observation_space}
        if self.has_critic_obs:
            info['state_space'] = self.state_space
            info['use_global_observations'] = True
        return info

    def seed(self, seed):
        pass

    def close(self):
        self.env.close()


def apply_velocity_stage_steps(cfg, stage_steps):
    """mjlab_vecenv.py:48-60: optional override of a step-scheduled command curriculum (switch points in env steps)."""
    stages = cfg.curriculum['command_vel'].params['velocity_stages']
    if len(stage_steps) != len(stages):
        raise ValueError(f"velocity_stage_steps has {len(stage_steps)} entries, task schedule has {len(stages)} stages")
    for stage, step in zip(stages, stage_steps):
        stage['step'] = int(step)


def resolve_sim_device(device, local_rank=None):
    """mjlab_vecenv.py:31-36: under torchrun a bare 'cuda' would put every rank's simulation on cuda:0."""
    local_rank = os.getenv('LOCAL_RANK') if local_rank is None else local_rank
    if device == 'cuda' and local_rank is not None:
        return f'cuda:{local_rank}'
    return device


class MjlabVecEnv(ManagerBasedEnvAdapter):
    """vecenv type 'MJLAB' (envs/mjlab_vecenv.py:23-71): ``env_config: {task_name, device, seed, velocity_stage_steps}``"""

    def __init__(self, config_name, num_actors, **kwargs):
        try:
            import warp as wp
            from mjlab.tasks.registry import load_env_cfg
            from mjlab.envs.manager_based_rl_env import ManagerBasedRlEnv
        except ImportError as e:
            raise ImportError("vecenv_type MJLAB needs the 'mjlab' and 'warp' packages (not part of this image); any env with "
                              "the manager-based API can be wrapped directly with ManagerBasedEnvAdapter(env)") from e
        wp.init()
        task_name = kwargs.pop('task_name', config_name)
        device = resolve_sim_device(kwargs.pop('device', 'cuda'))
        cfg = load_env_cfg(task_name)
        cfg.scene.num_envs = num_actors
        seed = kwargs.pop('seed', None)
        if seed is not None:        # the runner's seed (already rank-offset), so randomisation is reproducible
            cfg.seed = int(seed)
        stage_steps = kwargs.pop('velocity_stage_steps', None)
        if stage_steps is not None:
            apply_velocity_stage_steps(cfg, stage_steps)
        if kwargs:
            print(f"WARNING: mjlab wrapper ignoring unknown env_config keys: {sorted(kwargs)}")
        super().__init__(ManagerBasedRlEnv(cfg, device=device), device=device)


register_vecenv('MJLAB', lambda config_name, num_actors, **kwargs: MjlabVecEnv(config_name, num_actors, **kwargs))
