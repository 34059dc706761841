"""The bodies of the newest `-m gpu` tests (tests/test_zz_env_adapters_gpu.py, the min_sigma golden run of tests/test_agent_gpu.py) executed on the CPU with the torch stand-ins of every C-ABI
op (tests/_torch_ops.py), golden assertions ACTIVE: the test code, the adapter / observer host path and the expected values are checked
against the reference's golden runs before a GPU box is spent on them.  Says nothing about the kernels (the GPU run does)."""
import sys
import os

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def cpu_twin(monkeypatch):
    import _torch_ops
    from test_agent_host_cpu import _CudaLookingStr, _Event, _Stream
    import tests.test_agent_gpu as G
    import tests.test_zz_env_adapters_gpu as T
    _torch_ops.install_continuous(monkeypatch)
    monkeypatch.setitem(_torch_ops._USE_HOST_SCHED, 'on', True)
    monkeypatch.setattr(torch.cuda, 'set_device', lambda d: None)
    monkeypatch.setattr(torch.cuda, 'Event', _Event)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a: _Stream())
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self: self)
    dev = _CudaLookingStr('cpu')
    monkeypatch.setattr(G, 'DEV', dev)
    monkeypatch.setattr(T, 'DEV', dev)
    return T


def test_adapter_and_observer_gpu_test_body_holds_on_the_stand_ins(cpu_twin):
    cpu_twin.test_manager_based_adapter_and_isaac_observer_reproduce_the_reference_golden_run(False)


def test_critic_group_gpu_test_body_holds_on_the_stand_ins(cpu_twin):
    cpu_twin.test_critic_group_feeds_the_central_value_net_like_the_reference_golden_run(False)


def test_min_sigma_golden_gpu_test_body_holds_on_the_stand_ins(cpu_twin):
    import tests.test_agent_gpu as G
    G.test_agent_matches_reference_golden_more_config_keys('agent_minsigma.pt', False)
    G.test_agent_matches_reference_golden_more_config_keys('agent_misc.pt', False)        # the unchanged path through the same helper


def test_layerwise_min_sigma_gpu_test_body_holds_on_the_stand_ins(cpu_twin):
    import tests.test_agent_gpu as G
    G.test_layerwise_tensor_core_path_tracks_fp32_agent('mlp_128_64_32_min_sigma')


def test_separate_trunks_gpu_test_body_holds_on_the_stand_ins(cpu_twin):
    cpu_twin.test_separate_actor_critic_trunks_match_the_reference_golden_run(False)
