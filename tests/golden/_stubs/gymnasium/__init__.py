"""Import stub for `gymnasium` (not installed in this image, no network).

TEST INFRASTRUCTURE ONLY: lets /root/reference import in the build container so
tests/golden/gen_golden.py can drive the reference's own functions. The hot
path uses gymnasium only to describe shapes/dtypes (reference
rl_games/common/experience.py:385-398) -- no arithmetic lives here.
Never imported by product code.
"""
import types
import numpy as np
from . import spaces, vector, wrappers  # noqa: F401


class Env:
    metadata = {}
    observation_space = None
    action_space = None

    def reset(self, **kw):
        raise NotImplementedError

    def step(self, a):
        raise NotImplementedError


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.observation_space = getattr(env, 'observation_space', None)
        self.action_space = getattr(env, 'action_space', None)

    def __getattr__(self, name):
        return getattr(self.env, name)


class ObservationWrapper(Wrapper):
    pass


class RewardWrapper(Wrapper):
    pass


class ActionWrapper(Wrapper):
    pass


def register(*a, **kw):
    pass


def register_envs(*a, **kw):
    pass


def make(*a, **kw):
    raise RuntimeError('gymnasium stub: make() unavailable')


envs = types.SimpleNamespace(registry={})
