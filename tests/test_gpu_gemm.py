"""GPU: the tcgen05 GEMM kernel alone (through the C-ABI test entry mcm_test_linear) against fp64
matmuls of the SAME rounded operands -- the kernel's only remaining error is fp32 accumulation."""
import pytest
import torch

from motioncraft_b200.engine import test_linear as tc_linear
from tests import common as C

pytestmark = pytest.mark.gpu

SHAPES = [
    (128, 64, 64), (256, 256, 128), (300, 322, 512), (1000, 512, 322), (777, 196, 196), (4096, 1024, 512),
    (130, 2440, 2048), (64, 40, 24), (1, 8, 8), (129, 16, 1000), (515, 600, 75), (50176, 512, 512),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("fmt", [0, 1])
def test_linear_matches_rounded_operand_matmul(M, N, K, fmt):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    got = tc_linear(A.cuda(), W.cuda(), b.cuda(), fmt)
    if fmt == 0:   # fp16 operands: products are exact in fp32, so only accumulation order differs
        ref = A.cuda().half().double() @ W.cuda().half().double().T + b.cuda().double()
        tol = 2e-6
    else:          # bf16 hi/lo split, 3 passes: ~2^-16 relative per operand
        ref = A.cuda().double() @ W.cuda().double().T + b.cuda().double()
        tol = 2e-5
    assert C.rel_l2(got, ref) < tol


def test_linear_without_bias_and_fp16_saturation():
    A = torch.full((128, 64), 1e6)      # beyond fp16 range: operand packing saturates instead of producing inf
    W = torch.ones(64, 64) / 64
    got = tc_linear(A.cuda(), W.cuda(), None, 0)
    assert torch.isfinite(got).all()
    assert torch.allclose(got, torch.full_like(got, 65504.0), rtol=1e-3)


def test_bad_arguments_return_error_not_crash():
    from motioncraft_b200._lib import McmError
    with pytest.raises(McmError):
        tc_linear(torch.zeros(0, 8).cuda(), torch.zeros(8, 8).cuda(), None, 0)
    with pytest.raises(McmError):
        tc_linear(torch.zeros(8, 8).cuda(), torch.zeros(8, 8).cuda(), None, 7)
