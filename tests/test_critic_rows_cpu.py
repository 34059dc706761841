"""Row arithmetic of the central-value loss kernel (csrc/critic.cu::value_loss_row, __host__ __device__) on the CPU vs autograd."""
import ctypes

import pytest
import torch

from oracle import ppo_oracle as O
from tests import _hooks


@pytest.mark.parametrize('clip_value', [True, False])
def test_value_loss_rows_match_autograd(clip_value):
    lib = _hooks.load()
    fn = lib.b200rl_hosttest_value_loss_rows
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    g = torch.Generator().manual_seed(1)
    M = 1000
    v = torch.randn(M, generator=g).requires_grad_(True)
    old_v, ret = torch.randn(M, generator=g), torch.randn(M, generator=g)
    mask = (torch.rand(M, generator=g) < 0.7).float()
    w = (mask / mask.sum()).contiguous()
    c = O.critic_loss(old_v.unsqueeze(1), v.unsqueeze(1), 0.2, ret.unsqueeze(1), clip_value).squeeze(1)
    loss = (c * w).sum()
    loss.backward()
    dv = torch.zeros(M)
    s = ctypes.c_double(0.0)
    p = lambda t: ctypes.c_void_p(t.data_ptr())    # noqa: E731
    assert fn(p(v.detach().contiguous()), p(old_v), p(ret), p(w), M, 0.2, int(clip_value), p(dv), ctypes.byref(s)) == 0
    assert s.value == pytest.approx(float(loss), rel=1e-5)
    torch.testing.assert_close(dv, v.grad, rtol=1e-5, atol=1e-9)
