"""Target for `ncu --set full -k regex:gemm_tc_kernel`: one launch of each layer-wise tcgen05 GEMM (fwd / dgrad / wgrad, csrc/gemm_tc.cu) at
three c4 shapes, called exactly as the agent calls them (fp32 activations, bf16 weight twin, row splits).  9 launches, no timing here."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rl_games_b200 import ops  # noqa: E402


def main():
    dev = 'cuda'
    # (label, M, K, N, row splits): LSTM W_ih step (8192 sequences), MLP layers 1 / 2 of the update minibatch (32768 rows)
    for label, M, K, N, splits in (('lstm_ih', 8192, 348, 1024, 16), ('mlp1', 32768, 256, 512, 64), ('mlp2', 32768, 512, 256, 64)):
        X = torch.randn(M, K, device=dev); dY = torch.randn(M, N, device=dev) * 0.1
        flat = torch.randn(N * K + 8, device=dev) / K ** 0.5
        fb = torch.empty(flat.numel(), dtype=torch.bfloat16, device=dev)
        ops.cast_bf16(flat, fb)
        W = flat[8:].view(N, K); b = torch.zeros(N, device=dev)
        Y = torch.empty(M, N, device=dev); dX = torch.empty(M, K, device=dev)
        stride = N * K + N
        part = torch.empty(splits, stride, device=dev)
        ops.linear_fwd_tc(X, W, b, Y, 1, bf16_arena=(flat, fb))
        ops.linear_bwd_data_tc(dY, W, X, dX, 1, bf16_arena=(flat, fb))
        ops.linear_bwd_weight_tc(dY, X, part, part[:, N * K:], K, N, splits, split_stride=stride)
        torch.cuda.synchronize()
        print(label, 'ok', float(Y.abs().mean()), float(dX.abs().mean()), float(part.abs().mean()))


if __name__ == '__main__':
    main()
