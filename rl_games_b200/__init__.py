"""rl_games_b200 -- the PPO rollout -> GAE -> minibatch-update hot path of Denys88/rl_games as
hand-written sm_100a CUDA behind the reference's plugin surface (Runner / A2CAgent / compute_gae)."""
__version__ = '0.1.0'


def register(runner):
    """rl_games_b200.runner.register: the B200 agents behind `a2c_continuous` / `a2c_discrete` in the given runner's algo_factory
    (imported lazily: the package itself stays importable without torch / the built library)"""
    from .runner import register as _register
    return _register(runner)
