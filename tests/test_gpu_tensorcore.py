"""GPU: the tcgen05 (3xTF32) dense layer against fp64."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N", [(128, 32, 32), (1000, 35, 32), (4097, 64, 64), (300, 256, 256), (513, 100, 48),
                                   (77, 8, 2)])
def test_dense_tf32x3_matches_fp64(M, K, N):
    from asr_b200 import ops
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn((M, K), generator=g)
    w = torch.randn((K, N), generator=g) / K**0.5
    b = torch.randn(N, generator=g)
    packed = ops.pack_weights(w.cuda())
    out = ops.dense_tf32x3(a.cuda(), packed)
    ref = a.double() @ w.double()
    err = (out.cpu().double() - ref).abs().max().item()
    assert err <= 2e-5 * max(1.0, ref.abs().max().item()), err
    out = ops.dense_tf32x3(a.cuda(), packed, bias=b.cuda(), relu=True)
    ref = torch.relu(ref + b.double())
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
