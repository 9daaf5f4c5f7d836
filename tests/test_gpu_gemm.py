"""GPU: the tcgen05 3xTF32 GEMM (csrc/gemm_tc.cu) against torch fp64 on seeded inputs -- all three operand
layouts, ragged sizes (M, N, K not multiples of the 128 x {64,128} x 32 tile), unaligned leading dimensions
(the head layer's x is [N,5]), split-K reductions, bias and accumulate.  Tolerance: 1e-5 max-abs relative to
max|ref| (measured: <= 3e-6, dominated by the fp32 TMEM accumulator's rounding over long reductions) -- an order
below the 1e-4 parity bar; a plain TF32 product would give ~1e-3."""
import pytest
import torch

from util import max_rel

pytestmark = pytest.mark.gpu

TOL = 1e-5

SHAPES = [
    # mode, M, N, K
    (0, 128, 64, 32), (0, 128, 128, 64), (0, 300, 1024, 128), (0, 77, 512, 2304), (0, 50, 17, 256),
    (0, 1000, 64, 5), (0, 33, 24, 40), (0, 20000, 64, 64), (0, 1, 1, 1),
    (1, 128, 64, 64), (1, 300, 128, 1024), (1, 1250, 2304, 512), (1, 999, 5, 64), (1, 70, 40, 24),
    (2, 64, 64, 128), (2, 1024, 128, 5000), (2, 512, 2304, 1250), (2, 64, 5, 3000), (2, 17, 256, 50),
    (2, 64, 64, 100000),
]


def _operands(mode, M, N, K, g):
    if mode == 0:
        return torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    if mode == 1:
        return torch.randn(M, K, generator=g), torch.randn(K, N, generator=g)
    return torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)


def _ref(mode, a, b):
    a, b = a.double(), b.double()
    return a @ b.t() if mode == 0 else a @ b if mode == 1 else a.t() @ b


@pytest.mark.parametrize('mode,M,N,K', SHAPES)
def test_gemm_matches_fp64(mode, M, N, K):
    from yolat_vectorgraphicsrecognition_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + mode)
    a, b = _operands(mode, M, N, K, g)
    bias = torch.randn(N, generator=g)
    ref = _ref(mode, a, b)
    out = ops.gemm(mode, a.cuda(), b.cuda())
    assert max_rel(out, ref) < TOL, (mode, M, N, K, max_rel(out, ref))
    out2 = ops.gemm(mode, a.cuda(), b.cuda(), bias=bias.cuda())
    assert max_rel(out2, ref + bias.double()) < TOL
    base = torch.randn(M, N, generator=g)
    out3 = ops.gemm(mode, a.cuda(), b.cuda(), out=base.clone().cuda(), accumulate=True)
    assert max_rel(out3, ref + base.double()) < TOL


def test_gemm_strided_views():
    """Operands and output that are column slices of wider matrices (how torch.cat disappears on the path)."""
    from yolat_vectorgraphicsrecognition_b200 import ops
    g = torch.Generator().manual_seed(11)
    wide_a = torch.randn(500, 200, generator=g).cuda()
    wide_c = torch.zeros(500, 300).cuda()
    w = torch.randn(96, 70, generator=g).cuda()
    a = wide_a[:, 30:100]
    ops.gemm(0, a, w, out=wide_c[:, 100:196])
    ref = a.double().cpu() @ w.double().cpu().t()
    assert max_rel(wide_c[:, 100:196], ref) < TOL
    assert float(wide_c[:, :100].abs().max()) == 0.0 and float(wide_c[:, 196:].abs().max()) == 0.0


def test_gemm_large_magnitude_spread():
    """hi/lo splitting must hold for operands spanning many binades."""
    from yolat_vectorgraphicsrecognition_b200 import ops
    g = torch.Generator().manual_seed(13)
    a = torch.randn(256, 96, generator=g) * torch.logspace(-6, 6, 96)
    b = torch.randn(64, 96, generator=g) * torch.logspace(6, -6, 96)
    ref = a.double() @ b.double().t()
    assert max_rel(ops.gemm(0, a.cuda(), b.cuda()), ref) < TOL
