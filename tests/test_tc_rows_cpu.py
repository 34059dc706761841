"""Index arithmetic of the wide-observation tcgen05 kernels that can run on the CPU: the X-tile staging (csrc/mlp_tc.cu stage_x_cols,
__host__ __device__) executed thread by thread through its host test entry point, against the operand layout the MMAs expect -- the
INTERLEAVE bf16 tile [128 rows x 256 columns]: 16-byte chunk (row r, columns 8g..8g+7) at byte g*2048 + (r//8)*128 + (r%8)*16."""
import ctypes

import pytest
import torch

from tests import _hooks


def _decode(tile_u8, C=256):
    r = torch.arange(128).view(128, 1)
    c = torch.arange(C).view(1, C)
    off = ((c // 8) * 2048 + (r // 8) * 128 + (r % 8) * 16 + (c % 8) * 2) // 2
    return tile_u8.view(torch.bfloat16)[off.reshape(-1)].view(128, C).float()


@pytest.mark.parametrize('n_threads', [512, 256])
@pytest.mark.parametrize('D,rows_valid,norm', [(256, 128, True), (105, 128, True), (65, 77, True), (200, 1, False), (256, 128, False),
                                               (72, 128, True)])
def test_stage_x_cols_builds_the_operand_tile(D, rows_valid, norm, n_threads):
    lib = _hooks.load()
    g = torch.Generator().manual_seed(D * 1000 + rows_valid)
    row0 = 3
    obs = (torch.randn(row0 + 128, D, generator=g) * 3 + 0.5).contiguous()
    nm = (torch.randn(D, generator=g) * 0.3).contiguous() if norm else None
    ns = (torch.rand(D, generator=g) + 0.7).contiguous() if norm else None
    tile = torch.full((128 * 256 * 2,), 0xAB, dtype=torch.uint8)         # poison: every chunk must be written
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())   # noqa: E731
    rc = lib.b200rl_hosttest_stage_x_cols(p(obs), ctypes.c_int64(row0), rows_valid, D, p(nm), p(ns), n_threads, p(tile))
    assert rc == 0
    got = _decode(tile)
    x = obs[row0:row0 + 128].clone()
    if norm:
        x = torch.clamp((x - nm) * (1.0 / ns), -5.0, 5.0)
    want = torch.zeros(128, 256)
    want[:rows_valid, :D] = x[:rows_valid].to(torch.bfloat16).float()
    assert torch.equal(got, want)


def test_packed_weight_layout_of_the_wide_net():
    """one packed bf16 copy of all weights; W1 is [256 x 256] (observation columns >= D zero), W2 [128 x 256], W3 [64 x 128], heads [16 x 64]
    (rows > A zero); a weight tile [R x C] keeps chunk (r, cg) at cg * (R / 8) * 128 + (r // 8) * 128 + (r % 8) * 16"""
    from rl_games_b200 import ops
    lib = _hooks.load()
    D, A = 105, 8
    assert ops.tc_kind(D, [256, 128, 64], A) == 2 and ops.tc_kind(60, [256, 128, 64], A) == 1 and ops.tc_kind(300, [256, 128, 64], A) == 0
    g = torch.Generator().manual_seed(1)
    W = [torch.randn(256, D, generator=g), torch.randn(128, 256, generator=g), torch.randn(64, 128, generator=g), torch.randn(A + 1, 64, generator=g)]
    nbytes = ops.tc_pack_bytes(D, [256, 128, 64], A)
    assert nbytes == 2 * (256 * 256 + 128 * 256 + 64 * 128 + 16 * 64)
    buf = torch.full((nbytes,), 0xAB, dtype=torch.uint8)
    rc = lib.b200rl_hosttest_pack_weights_wide(*[ctypes.c_void_p(w.data_ptr()) for w in W], D, A, ctypes.c_void_p(buf.data_ptr()))
    assert rc == 0
    off = 0
    for w, (R, C) in zip(W, ((256, 256), (128, 256), (64, 128), (16, 64))):
        r = torch.arange(R).view(R, 1)
        c = torch.arange(C).view(1, C)
        idx = (off + (c // 8) * (R // 8) * 128 + (r // 8) * 128 + (r % 8) * 16 + (c % 8) * 2) // 2
        got = buf.view(torch.bfloat16)[idx.reshape(-1)].view(R, C).float()
        want = torch.zeros(R, C)
        want[:w.shape[0], :w.shape[1]] = w.to(torch.bfloat16).float()
        assert torch.equal(got, want)
        off += R * C * 2
