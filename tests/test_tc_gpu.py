"""tcgen05 building blocks (UMMA descriptors, TMEM load, mbarrier commit) against torch on the same bf16 inputs.
Tolerance: fp32 accumulation of exact bf16 products -> rtol 1e-5 vs a float64 reference of the bf16 values."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('N,K', [(16, 16), (64, 64), (256, 64), (128, 256), (64, 128), (16, 64), (256, 128)])
@pytest.mark.parametrize('a_mn,b_mn', [(False, False), (True, False), (False, True), (True, True)])
def test_tc_gemm(N, K, a_mn, b_mn):
    from rl_games_b200 import ops
    g = torch.Generator().manual_seed(N * 7 + K)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)
    ref = (A.double() @ B.double().t()).float()
    Ad = (A.t().contiguous() if a_mn else A).to(DEV)
    Bd = (B.t().contiguous() if b_mn else B).to(DEV)
    D = ops.tc_gemm_test(Ad, Bd, N, K, a_mn, b_mn)
    torch.cuda.synchronize()
    torch.testing.assert_close(D.cpu(), ref, rtol=1e-4, atol=1e-3)
