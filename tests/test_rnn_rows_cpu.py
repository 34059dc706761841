"""Element function of b200rl_rnn_train_dones_u8 (csrc/rnn.cu) on the CPU through its host test entry point, against the reference's
expression (a2c_common.py:1180-1191: rnn_dones[1:] = max(rnn_dones[1:], (mb_valid == 0)[:-1]))."""
import ctypes

import torch

from tests import _hooks


def test_rnn_train_dones_rows_match_reference_expression():
    lib = _hooks.load()
    g = torch.Generator().manual_seed(0)
    for H, N in ((1, 7), (8, 5), (16, 33)):
        dones = (torch.rand(H, N, generator=g) < 0.3).to(torch.uint8)
        valid = (torch.rand(H, N, generator=g) < 0.7).float()
        out = torch.full((H, N), 9, dtype=torch.uint8)
        rc = lib.b200rl_hosttest_rnn_train_dones(ctypes.c_void_p(dones.data_ptr()), ctypes.c_void_p(valid.data_ptr()),
                                                 ctypes.c_void_p(out.data_ptr()), H, N)
        assert rc == 0
        ref = dones.clone()
        ref[1:] = torch.maximum(ref[1:], (valid == 0.0)[:-1].to(torch.uint8))
        assert torch.equal(out, ref)
