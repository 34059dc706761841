"""rl_games_b200 -- the PPO rollout -> GAE -> minibatch-update hot path of Denys88/rl_games as
hand-written sm_100a CUDA behind the reference's plugin surface (Runner / A2CAgent / compute_gae)."""
__version__ = '0.1.0'
