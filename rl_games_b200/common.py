"""Host-side mirrors of the small reference interfaces the PPO path plugs into.

These keep the reference's names, argument meaning and error behaviour so a config / env / observer
written for rl_games drops in unchanged when rl_games itself is not importable (e.g. on a box without
gymnasium): ObjectFactory (common/object_factory.py), IVecEnv (common/ivecenv.py), the vecenv / env
registries (common/vecenv.py:368-391, env_configurations.py:358-366), AlgoObserver hooks
(common/algo_observer.py:6-26), DefaultRewardsShaper (common/tr_helpers.py:16-42) and the LR schedulers
(common/schedulers.py).  When rl_games IS importable, INTEGRATION.md shows how to register the B200
agent into its own Runner instead.
"""
import math

import torch


class ObjectFactory:
    """common/object_factory.py:1-39"""

    def __init__(self):
        self._builders = {}

    def register_builder(self, name, builder):
        self._builders[name] = builder

    def set_builders(self, builders):
        self._builders = builders

    def create(self, name, **kwargs):
        builder = self._builders.get(name)
        if not builder:
            raise ValueError(name)
        return builder(**kwargs)


class IVecEnv:
    """common/ivecenv.py:1-36"""

    def step(self, actions):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def has_action_masks(self):
        return False

    def get_number_of_agents(self):
        return 1

    def get_env_info(self):
        pass

    def seed(self, seed):
        pass

    def set_train_info(self, env_frames, *args, **kwargs):
        pass

    def get_env_state(self):
        return None

    def set_env_state(self, env_state):
        pass


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (only .shape/.dtype/.low/.high are read on the path,
    experience.py:385-398, a2c_common.py:1496-1497).  Real gymnasium spaces work too (duck-typed by class name)."""

    def __init__(self, low, high, shape, dtype='float32'):
        import numpy as np
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()


class Discrete:
    """Minimal stand-in for gymnasium.spaces.Discrete (only .n is read: a2c_common.py:1213-1216); duck-typed by class name."""

    def __init__(self, n):
        self.n = int(n)
        self.shape = ()


class Tuple:
    """Minimal stand-in for gymnasium.spaces.Tuple (multi-discrete action spaces: a2c_common.py:1217-1220 reads [a.n for a in space])"""

    def __init__(self, spaces):
        self.spaces = list(spaces)

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)


# ---- registries: common/vecenv.py:368-391 and common/env_configurations.py:358-366 ----
vecenv_config = {}
configurations = {}


def register_vecenv(config_name, func):
    vecenv_config[config_name] = func


def register_env(name, config):
    configurations[name] = config


def create_vec_env(config_name, num_actors, **kwargs):
    """common/vecenv.py:379-391.  A name that is not registered HERE is looked up in the reference's own registries when rl_games is
    importable (the agent running under the reference's Runner): `env_name: envpool` / `gymnasium` / `ray` /
    anything the user registered through `rl_games.common.vecenv.register` + `env_configurations.register` keeps working without an
    injected `vec_env`."""
    if config_name not in configurations:
        try:
            from rl_games.common import vecenv as reference_vecenv
        except ImportError:
            raise KeyError(f"env '{config_name}' is not registered (rl_games_b200.common.register_env; registered: {sorted(configurations)}) "
                           f"and rl_games is not importable to look it up there") from None
        return reference_vecenv.create_vec_env(config_name, num_actors, **kwargs)
    config = configurations[config_name]
    vec_env_name = config['vecenv_type']
    merged = {**config.get('default_env_config', {}), **kwargs}
    return vecenv_config[vec_env_name](config_name, num_actors, **merged)


class AlgoObserver:
    """common/algo_observer.py:6-26"""

    def before_init(self, base_name, config, experiment_name):
        pass

    def after_init(self, algo):
        pass

    def process_infos(self, infos, done_indices):
        pass

    def after_steps(self):
        pass

    def after_clear_stats(self):
        pass

    def after_print_stats(self, frame, epoch_num, total_time):
        pass


class DefaultAlgoObserver(AlgoObserver):
    """The reference's default observer only consumes list-of-dict infos ('scores'/'battle_won'); tensor
    envs hand over a dict without those keys, so nothing needs the done indices on the host."""
    wants_done_indices = False

    def after_init(self, algo):
        self.algo = algo
        self.writer = algo.writer


class IsaacAlgoObserver(AlgoObserver):
    """common/algo_observer.py:95-156 (``algo_observer: isaac``): logs the env's per-episode metrics (``infos['episode']``, a
    dict of tensors emitted on reset bursts) as ``Episode/<key>`` = mean over everything seen since the last print, plus scalar
    entries of ``infos`` as ``<key>/frame|iter|time``.

    The reference keeps every burst's dict and concatenates them at print time; here each key owns a device-resident
    (sum, count) pair updated with two stream-ordered adds per burst -- same mean, O(1) memory, and no host
    synchronisation until ``after_print_stats`` reads one scalar per key, once per epoch.  ``done_indices`` is never used
    by this observer, so it does not ask the agent for them (``wants_done_indices = False``: no per-step ``nonzero`` sync);
    it does need the per-step infos (``wants_infos``), which keeps the env step out of the captured rollout graph."""
    wants_done_indices = False
    wants_infos = True

    def after_init(self, algo):
        self.algo = algo
        self.writer = algo.writer
        self.ep_sums = {}
        self.direct_info = {}

    def process_infos(self, infos, done_indices=None):
        if not isinstance(infos, dict):
            raise ValueError(f"{self.__class__.__name__} expected 'infos' as dict. Received: {type(infos)}")
        ep = infos.get('episode')
        if ep:
            dev = getattr(self.algo, 'ppo_device', None)
            for k, v in ep.items():
                v = v if isinstance(v, torch.Tensor) else torch.tensor([float(v)])
                v = v.detach().reshape(-1).to(device=dev, dtype=torch.float32)
                acc = self.ep_sums.get(k)
                if acc is None:
                    acc = self.ep_sums[k] = torch.zeros(2, dtype=torch.float32, device=v.device)
                acc[0] += v.sum()
                acc[1] += v.numel()
        if len(infos) > 0:      # direct logging from the env: scalars only
            self.direct_info = {k: v for k, v in infos.items()
                                if isinstance(v, (float, int)) or (isinstance(v, torch.Tensor) and v.dim() == 0)}

    def after_print_stats(self, frame, epoch_num, total_time):
        if self.ep_sums:
            keys = sorted(self.ep_sums)
            host = torch.stack([self.ep_sums[k] for k in keys]).cpu()       # one D2H for all keys
            for k, (s, n) in zip(keys, host.tolist()):
                if n > 0:
                    self.writer.add_scalar('Episode/' + k, s / n, epoch_num)
            self.ep_sums.clear()
        for k, v in self.direct_info.items():
            self.writer.add_scalar(f'{k}/frame', v, frame)
            self.writer.add_scalar(f'{k}/iter', v, epoch_num)
            self.writer.add_scalar(f'{k}/time', v, total_time)


class DefaultRewardsShaper:
    """common/tr_helpers.py:16-42 (parameters only; the arithmetic runs in b200rl_post_step_f32)."""

    def __init__(self, scale_value=1, shift_value=0, min_val=-math.inf, max_val=math.inf, log_val=False, is_torch=True):
        self.scale_value = scale_value
        self.shift_value = shift_value
        self.min_val = min_val
        self.max_val = max_val
        self.log_val = log_val
        self.is_torch = is_torch

    def __call__(self, reward):
        import torch
        reward = (reward + self.shift_value) * self.scale_value
        reward = torch.clamp(reward, self.min_val, self.max_val)
        return torch.log(reward) if self.log_val else reward


# ---- schedulers: common/schedulers.py:1-58 ----
class RLScheduler:
    def update(self, current_lr, entropy_coef, epoch, frames, **kwargs):
        pass


class IdentityScheduler(RLScheduler):
    def update(self, current_lr, entropy_coef, epoch, frames, kl_dist, **kwargs):
        return current_lr, entropy_coef


class AdaptiveScheduler(RLScheduler):
    def __init__(self, kl_threshold=0.008, min_lr=1e-6, max_lr=1e-2, lr_multiplier=1.5):
        self.min_lr = min_lr
        self.max_lr = max_lr
        self.kl_threshold = kl_threshold
        self.lr_multiplier = lr_multiplier

    def update(self, current_lr, entropy_coef, epoch, frames, kl_dist, **kwargs):
        lr = current_lr
        if kl_dist > (2.0 * self.kl_threshold):
            lr = max(current_lr / self.lr_multiplier, self.min_lr)
        if kl_dist < (0.5 * self.kl_threshold):
            lr = min(current_lr * self.lr_multiplier, self.max_lr)
        return lr, entropy_coef


class LinearScheduler(RLScheduler):
    def __init__(self, start_lr, min_lr=1e-6, max_steps=1000000, use_epochs=True, apply_to_entropy=False, **kwargs):
        self.start_lr = start_lr
        self.min_lr = min_lr
        self.max_steps = max_steps
        self.use_epochs = use_epochs
        self.apply_to_entropy = apply_to_entropy
        if apply_to_entropy:
            self.start_entropy_coef = kwargs.pop('start_entropy_coef', 0.01)
            self.min_entropy_coef = kwargs.pop('min_entropy_coef', 0.0001)

    def update(self, current_lr, entropy_coef, epoch, frames, kl_dist, **kwargs):
        steps = epoch if self.use_epochs else frames
        mul = max(0, self.max_steps - steps) / self.max_steps
        lr = self.min_lr + (self.start_lr - self.min_lr) * mul
        if self.apply_to_entropy:
            entropy_coef = self.min_entropy_coef + (self.start_entropy_coef - self.min_entropy_coef) * mul
        return lr, entropy_coef


class NullWriter:
    """Stand-in for tensorboardX.SummaryWriter when it is not installed (a2c_common.py:455-462)."""

    def add_scalar(self, *a, **kw):
        pass

    def add_scalars(self, *a, **kw):
        pass

    def flush(self):
        pass

    def close(self):
        pass


def make_summary_writer(path):
    try:
        from tensorboardX import SummaryWriter
        return SummaryWriter(path)
    except Exception:
        try:
            from torch.utils.tensorboard import SummaryWriter
            return SummaryWriter(path)
        except Exception:
            return NullWriter()
