"""Layer-wise tensor-core GEMMs (csrc/gemm_tc.cu: b200rl_linear_{fwd,bwd_data,bwd_weight}_tc) against their arithmetic contract restated
in torch: operands rounded to bf16 (after the optional observation normalisation), exact products, fp64 accumulation, fp32 bias /
activation / outputs.  Differences left: fp32 accumulation order inside the tensor core -> rtol 2e-4 of the output scale.  Shapes: the
c4 LSTM gate GEMM (348 + 256 -> 1024), the Humanoid MLP layers, the 18-wide head, ragged everything, chunked arena rows, row splits."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda'
ACT = {0: lambda x: x, 1: F.elu, 2: torch.relu, 3: torch.tanh}
DACT = {0: lambda a: torch.ones_like(a), 1: lambda a: torch.where(a > 0, torch.ones_like(a), a + 1.0), 2: lambda a: (a > 0).float(),
        3: lambda a: 1.0 - a * a}


def bf(x):
    return x.to(torch.bfloat16).double()


def close(got, ref, scale_rtol=2e-4):
    tol = scale_rtol * float(ref.abs().max()) + 1e-6
    err = float((got.double() - ref).abs().max())
    assert err <= tol, (err, tol)


@pytest.mark.parametrize('M,K,N,act,accumulate', [(1000, 348, 1024, 0, False), (1000, 256, 1024, 0, True), (300, 60, 18, 1, False),
                                                  (4096, 256, 512, 1, False), (129, 17, 65, 3, False), (64, 512, 256, 2, False)])
def test_linear_fwd_tc(M, K, N, act, accumulate):
    from rl_games_b200 import ops
    g = torch.Generator().manual_seed(M + K + N)
    X = torch.randn(M, K, generator=g).to(DEV); W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV); b = (torch.randn(N, generator=g) * 0.1).to(DEV)
    Y0 = torch.randn(M, N, generator=g).to(DEV)
    Y = Y0.clone()
    ops.linear_fwd_tc(X, W, b, Y, act, accumulate=accumulate)
    torch.cuda.synchronize()
    ref = bf(X) @ bf(W).t() + b.double() + (Y0.double() if accumulate else 0.0)
    close(Y, ACT[act](ref))
    # weights taken from a bf16 twin of the parameter arena (what the agent does): W is a view at a non-zero offset of a flat arena
    off = 8      # keeps the bf16 rows 16-byte aligned when K % 8 == 0; other K exercise the 8-byte / scalar copies
    flat = torch.zeros(off + N * K, device=DEV); flat[off:] = W.reshape(-1)
    fb = torch.empty(flat.numel(), dtype=torch.bfloat16, device=DEV)
    ops.cast_bf16(flat, fb)
    Wv = flat[off:].view(N, K)
    Y2 = Y0.clone()
    ops.linear_fwd_tc(X, Wv, b, Y2, act, accumulate=accumulate, bf16_arena=(flat, fb))
    torch.cuda.synchronize()
    assert torch.equal(Y2, Y)          # same bf16 operand values either way -> bit-identical


def test_linear_fwd_tc_chunked_rows_with_normalisation():
    """first layer on arena rows: minibatch = H chunks of epm rows at stride N (time-major arena), obs normalised and clamped on the fly"""
    from rl_games_b200 import ops
    g = torch.Generator().manual_seed(5)
    H, Nenv, epm, D, U = 4, 640, 256, 348, 1024
    obs = (torch.randn(H, Nenv, D, generator=g) * 2 + 0.3).to(DEV)
    nm = (torch.randn(D, generator=g) * 0.3).to(DEV); ns = (torch.rand(D, generator=g) + 0.7).to(DEV)
    W = (torch.randn(U, D, generator=g) / D ** 0.5).to(DEV); b = torch.zeros(U, device=DEV)
    e0 = 128
    Y = torch.empty(H * epm, U, device=DEV)
    ops.linear_fwd_tc(obs[0, e0:], W, b, Y, 0, rows_per_chunk=epm, chunk_stride=Nenv, x_ld=D, norm_mean=nm, norm_std=ns, M=H * epm)
    torch.cuda.synchronize()
    x = torch.cat([obs[t, e0:e0 + epm] for t in range(H)])
    xn = torch.clamp((x - nm) / ns, -5.0, 5.0)
    close(Y, bf(xn) @ bf(W).t())


@pytest.mark.parametrize('M,K,N,act_prev', [(777, 348, 1024, 0), (2048, 512, 256, 1), (130, 20, 17, 3), (512, 256, 1024, 2)])
def test_linear_bwd_data_tc(M, K, N, act_prev):
    from rl_games_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    dY = torch.randn(M, N, generator=g).to(DEV); W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    A_prev = ACT[act_prev](torch.randn(M, K, generator=g)).to(DEV)
    dX = torch.empty(M, K, device=DEV)
    ops.linear_bwd_data_tc(dY, W, A_prev if act_prev else None, dX, act_prev)
    torch.cuda.synchronize()
    ref = bf(dY) @ bf(W)
    if act_prev:
        ref = ref * DACT[act_prev](A_prev).double()
    close(dX, ref)
    flat = torch.zeros(4 + N * K, device=DEV); flat[4:] = W.reshape(-1)
    fb = torch.empty(flat.numel(), dtype=torch.bfloat16, device=DEV)
    ops.cast_bf16(flat, fb)
    dX2 = torch.empty(M, K, device=DEV)
    ops.linear_bwd_data_tc(dY, flat[4:].view(N, K), A_prev if act_prev else None, dX2, act_prev, bf16_arena=(flat, fb))
    torch.cuda.synchronize()
    assert torch.equal(dX2, dX)


@pytest.mark.parametrize('M,K,N,splits', [(5000, 348, 1024, 7), (4096, 256, 512, 16), (1000, 60, 18, 3), (130, 17, 65, 1), (8192, 512, 256, 64)])
def test_linear_bwd_weight_tc(M, K, N, splits):
    from rl_games_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    dY = (torch.randn(M, N, generator=g) * 0.1).to(DEV); X = torch.randn(M, K, generator=g).to(DEV)
    stride = N * K + N
    part = torch.full((splits, stride), float('nan'), device=DEV)
    ops.linear_bwd_weight_tc(dY, X, part, part[:, N * K:], K, N, splits, split_stride=stride)
    out = torch.empty(stride, device=DEV)
    ops.reduce_splits(part, out, stride, splits, split_stride=stride)
    torch.cuda.synchronize()
    assert torch.isfinite(part).all()
    close(out[:N * K].view(N, K), bf(dY).t() @ bf(X))
    close(out[N * K:], bf(dY).sum(0))


def test_linear_bwd_weight_tc_chunked_rows_with_normalisation():
    from rl_games_b200 import ops
    g = torch.Generator().manual_seed(9)
    H, Nenv, epm, D, U, splits = 4, 640, 256, 60, 64, 5
    obs = (torch.randn(H, Nenv, D, generator=g) * 2 + 0.3).to(DEV)
    nm = (torch.randn(D, generator=g) * 0.3).to(DEV); ns = (torch.rand(D, generator=g) + 0.7).to(DEV)
    M, e0 = H * epm, 128
    dY = (torch.randn(M, U, generator=g) * 0.1).to(DEV)
    stride = U * D + U
    part = torch.full((splits, stride), float('nan'), device=DEV)
    ops.linear_bwd_weight_tc(dY, obs[0, e0:], part, part[:, U * D:], D, U, splits, rows_per_chunk=epm, chunk_stride=Nenv, x_ld=D,
                             norm_mean=nm, norm_std=ns, M=M, split_stride=stride)
    out = torch.empty(stride, device=DEV)
    ops.reduce_splits(part, out, stride, splits, split_stride=stride)
    torch.cuda.synchronize()
    x = torch.cat([obs[t, e0:e0 + epm] for t in range(H)])
    xn = torch.clamp((x - nm) / ns, -5.0, 5.0)
    close(out[:U * D].view(U, D), bf(dY).t() @ bf(xn))
