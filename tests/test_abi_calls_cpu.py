"""Every C-ABI call the agents make, checked against the prototypes of include/b200rl.h without a GPU: `ops.lib` is replaced by a
recorder that verifies argument COUNT and ctypes convertibility of every call against the header-derived prototypes (the same
argtypes the real binding installs) and returns success without computing anything.  Catches binding drift (a wrapper passing one
argument too few or a tensor where a scalar belongs) on the paths no CPU stand-in exercises, because the stand-ins replace the
wrappers themselves.  Pure host-side queries (geometry, sizes, tables) are delegated to the real library."""
import ctypes
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_agent_host_cpu as H  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')
HOST_ONLY = {'b200rl_tc_supported', 'b200rl_tc_pack_bytes', 'b200rl_tc_tile_bytes', 'b200rl_tc_xtile_bytes', 'b200rl_tc_pack_table',
             'b200rl_loss_partial_stride', 'b200rl_built_arch', 'b200rl_set_pdl', 'b200rl_gae_set_tma'}


class _Recorder:
    def __init__(self, real):
        self.real, self.calls = real, {}

    def __getattr__(self, name):
        if not name.startswith('b200rl_'):
            raise AttributeError(name)
        if name in HOST_ONLY:
            return getattr(self.real, name)
        sig = self.real.protos[name]          # KeyError: the wrapper calls something the header does not declare

        def call(*args):
            assert len(args) == len(sig), f'{name}: {len(args)} arguments passed, the header declares {len(sig)}'
            for a, (an, t) in zip(args, sig):
                assert not isinstance(a, (torch.Tensor, bool)) or t is not ctypes.c_void_p or not isinstance(a, torch.Tensor), \
                    f'{name}: tensor passed for {an}'
                try:
                    t.from_param(a)
                except (TypeError, ctypes.ArgumentError) as e:
                    raise AssertionError(f'{name}: argument {an} = {a!r} is not a {t.__name__}') from e
            self.calls[name] = self.calls.get(name, 0) + 1
            return 0
        return call


def _patch(monkeypatch):
    from rl_games_b200 import ops, _lib
    rec = _Recorder(_lib.lib)
    monkeypatch.setattr(ops, 'lib', rec)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    monkeypatch.setattr(torch.cuda, 'Event', H._Event)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a: H._Stream())
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self: self)
    return rec


class _ZeroEnv:
    def __init__(self, N, D, A, autoreset='same_step'):
        self.N, self.D, self.A, self.autoreset = N, D, A, autoreset

    def reset(self):
        return torch.zeros(self.N, self.D)

    def step(self, actions):
        return torch.zeros(self.N, self.D), torch.zeros(self.N), torch.zeros(self.N, dtype=torch.uint8), {'time_outs': torch.zeros(self.N, dtype=torch.uint8)}

    def get_env_info(self):
        from rl_games_b200.common import Box
        info = {'observation_space': Box(-1, 1, (self.D,)), 'action_space': Box(-1.0, 1.0, (self.A,))}
        if self.autoreset != 'same_step':
            info['autoreset_mode'] = self.autoreset
        return info


def _agent(tmp_path, N, H_, D, A, units, mb, over=None, rnn=None, autoreset='same_step'):
    from rl_games_b200.runner import Runner
    env = _ZeroEnv(N, D, A, autoreset)
    config = {'name': 'abi', 'env_name': 'unused', 'reward_shaper': {'scale_value': 1.0}, 'device': H._CudaLookingStr('cpu'),
              'normalize_input': True, 'normalize_value': True, 'normalize_advantage': True, 'value_bootstrap': True, 'gamma': 0.99,
              'tau': 0.95, 'learning_rate': 3e-4, 'lr_schedule': 'adaptive', 'kl_threshold': 0.008, 'grad_norm': 1.0, 'truncate_grads': True,
              'entropy_coef': 0.0, 'e_clip': 0.2, 'clip_value': True, 'num_actors': N, 'horizon_length': H_, 'minibatch_size': mb,
              'mini_epochs': 2, 'critic_coef': 2, 'bounds_loss_coef': 0.0001, 'train_dir': str(tmp_path), 'b200_cuda_graph': False,
              'mixed_precision': False, 'env_info': env.get_env_info(), 'vec_env': env, 'print_stats': False}
    config.update(over or {})
    network = {'name': 'actor_critic', 'separate': False,
               'space': {'continuous': {'mu_activation': 'None', 'sigma_activation': 'None', 'mu_init': {'name': 'default'},
                                        'sigma_init': {'name': 'const_initializer', 'val': 0}, 'fixed_sigma': True}},
               'mlp': {'units': list(units), 'activation': 'elu', 'initializer': {'name': 'default'}}}
    if rnn:
        network['rnn'] = rnn
    r = Runner()
    r.load({'params': {'seed': 7, 'algo': {'name': 'a2c_continuous'}, 'model': {'name': 'continuous_a2c_logstd'}, 'network': network,
                       'config': config}})
    r.params['config']['vec_env'] = env
    a = r.algo_factory.create(r.algo_name, base_name='abi', params=r.params)
    a.init_tensors()
    a._repack()
    a.obs = a.env_reset()
    return a


CASES = {
    'fp32': dict(N=8, H_=8, D=6, A=3, units=(16, 8), mb=32),
    'fp32 masked, rms advantage, standard schedule': dict(N=8, H_=8, D=8, A=3, units=(16, 12, 8), mb=32, autoreset='next_step',
                                                           over={'normalize_rms_advantage': True, 'schedule_type': 'standard'}),
    'lstm before the mlp': dict(N=8, H_=8, D=6, A=3, units=(16, 8), mb=32, over={'seq_length': 4},
                                rnn={'name': 'lstm', 'units': 8, 'layers': 1, 'before_mlp': True}),
    'lstm after the mlp': dict(N=8, H_=8, D=6, A=3, units=(16, 8), mb=32, over={'seq_length': 4},
                               rnn={'name': 'lstm', 'units': 12, 'layers': 1, 'before_mlp': False}),
    'lstm before the mlp, next_step autoreset': dict(N=8, H_=8, D=6, A=3, units=(16, 8), mb=32, autoreset='next_step',
                                                     over={'seq_length': 4},
                                                     rnn={'name': 'lstm', 'units': 8, 'layers': 1, 'before_mlp': True}),
    'lstm after the mlp on the tensor cores (layer-wise GEMMs)': dict(N=8, H_=8, D=6, A=3, units=(16, 8), mb=32, over={'seq_length': 4, 'mixed_precision': True},
                                                                  rnn={'name': 'lstm', 'units': 12, 'layers': 1, 'before_mlp': False}),
    'mlp [512,256,128] on the tensor cores (layer-wise GEMMs)': dict(N=16, H_=8, D=20, A=17, units=(512, 256, 128), mb=64, over={'mixed_precision': True}),
    'tcgen05': dict(N=256, H_=4, D=60, A=8, units=(256, 128, 64), mb=512, over={'mixed_precision': True}),
    'tcgen05, pipelined wgrad + masked': dict(N=256, H_=4, D=60, A=8, units=(256, 128, 64), mb=512, autoreset='next_step',
                                              over={'mixed_precision': True, 'b200_pipelined_wgrad': True}),
    'tcgen05 wide observations': dict(N=256, H_=4, D=256, A=8, units=(256, 128, 64), mb=512,
                                      over={'mixed_precision': True}),
}


@pytest.mark.parametrize('case', list(CASES))
def test_continuous_agent_calls_match_the_header(case, monkeypatch, tmp_path):
    rec = _patch(monkeypatch)
    a = _agent(tmp_path, **CASES[case])
    for _ in range(2):
        a.epoch_num += 1
        a.train_epoch()                     # device RNG path (no noise tape): the Philox arguments are exercised too
    a.get_full_state_weights()
    assert rec.calls, 'no kernel call was made'
    if 'layer-wise' in case:
        assert a.gemm_tc and not a.use_tc
        for n in ('b200rl_linear_fwd_tc', 'b200rl_linear_bwd_data_tc', 'b200rl_linear_bwd_weight_tc'):
            assert n in rec.calls, (n, sorted(rec.calls))
        assert 'b200rl_linear_fwd_f32' not in rec.calls
    want = {'tcgen05': ['b200rl_tc_mlp_fwd_train', 'b200rl_tc_mlp_bwd', 'b200rl_tc_mlp_fwd_rollout', 'b200rl_reduce_adam_f32'],
            'lstm': ['b200rl_lstm_cell_fwd_f32', 'b200rl_lstm_cell_bwd_f32'], 'fp32': ['b200rl_ppo_head_loss_f32', 'b200rl_gae_fused_f32']}
    for key, names in want.items():
        if case.startswith(key):
            for n in names:
                assert n in rec.calls, (n, sorted(rec.calls))


def test_central_value_and_discrete_agent_calls_match_the_header(monkeypatch, tmp_path):
    rec = _patch(monkeypatch)
    import test_agent_cv_host_cpu as HC
    import _torch_ops
    g = torch.load(os.path.join(GOLDEN, 'agent_cv.pt'), weights_only=False)

    class _P:       # _build_cv installs the torch stand-ins first; undo that so the real wrappers (and the recorder) are what runs
        def setattr(self, target, name, value):
            from rl_games_b200 import ops
            if target is not ops:
                monkeypatch.setattr(target, name, value)
    a = HC._build_cv(_P(), tmp_path, g)
    a.epoch_num += 1
    a.train_epoch()
    assert 'b200rl_value_loss_f32' in rec.calls
    del _torch_ops
    # discrete / multi-discrete
    import test_discrete_host_cpu as HD
    for name in ('agent_discrete.pt', 'agent_discrete_masked.pt', 'agent_multidiscrete.pt'):
        gd = torch.load(os.path.join(GOLDEN, name), weights_only=False)
        d = HD._build(monkeypatch, tmp_path, gd, stand_ins=False)
        d.epoch_num += 1
        d.train_epoch()                     # device RNG path (no uniform tape)
    assert 'b200rl_categorical_sample_f32' in rec.calls and 'b200rl_categorical_loss_f32' in rec.calls


def test_synthetic_env_and_gae_dropin_calls_match_the_header(monkeypatch):
    rec = _patch(monkeypatch)
    from rl_games_b200 import ops
    from rl_games_b200.envs import SyntheticGPUEnv
    monkeypatch.setattr(ops, '_need_cuda', lambda *ts: None)
    env = SyntheticGPUEnv('b200_synthetic', 16, obs_dim=6, act_dim=3, device='cpu')
    env.reset()
    env.begin_rollout()
    env.step(torch.zeros(16, 3))
    env.end_rollout()
    r = torch.zeros(4, 16, 1)
    ops.compute_gae(r, r, torch.zeros(4, 16), torch.zeros(16, 1), torch.zeros(16), 0.99, 0.95)
    ops.compute_gae(r, r, torch.zeros(4, 16, dtype=torch.uint8), torch.zeros(16, 1), torch.zeros(16, dtype=torch.bool), 0.99, 0.95,
                    returns_out=torch.zeros(4, 16, 1))
    for n in ('b200rl_synth_env_step', 'b200rl_bump_u64', 'b200rl_gae_f32'):
        assert n in rec.calls, (n, sorted(rec.calls))
