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


@pytest.mark.parametrize("env", [{"MCM_MAX_CLUSTER": "4"}, {"MCM_MAX_CLUSTER": "2", "MCM_PREFETCH": "1"},
                                 {"MCM_GENERIC_EPILOGUE": "1"}])
def test_kernel_variants_in_subprocess(env):
    """Cluster-multicast (B tile shared by 2 / 4 CTAs), producer L2 prefetch and the non-TMA fallback epilogue
    are selected by environment at library init, so each runs in a fresh process: GEMM shapes + a full denoise."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from motioncraft_b200.engine import test_linear, DenoiserEngine
from tests import common as C
g = torch.Generator().manual_seed(3)
for (M, N, K) in [(515, 600, 75), (1000, 512, 322), (4096, 1024, 512), (130, 2440, 2048)]:
    A = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    got = test_linear(A.cuda(), W.cuda(), b.cuda(), 0)
    ref = A.cuda().half().double() @ W.cuda().half().double().T + b.cuda().double()
    assert C.rel_l2(got, ref) < 2e-6, (M, N, K)
T, B = 60, 3
sd = C.base_state(T)
x, xf_out, xf_proj = C.inputs(B, T)
want = C.oracle_forward(sd, x, 321, xf_proj, xf_out, torch.float64)
eng = DenoiserEngine(C.hot(sd), seq_len=T, max_batch=B)
eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
assert C.rel_l2(eng.denoise(x.cuda(), 321), want) < 1e-3
print("variant ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
