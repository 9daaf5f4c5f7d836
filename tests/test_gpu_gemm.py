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


def test_gemm_tma_edges():
    """The TMA-fed path at its edges: rows spilling over the last full wave of 148 tiles (tail split along K), outputs
    narrower than one 32-column store box, reductions that are not a multiple of the 32-wide k-block, column-slice
    operands (row pitch != row length) and an output slice (TMA stores must clip to the slice)."""
    from yolat_vectorgraphicsrecognition_b200 import ops
    g = torch.Generator().manual_seed(21)
    for mode, M, N, K in ((1, 20000, 128, 1024), (0, 19000, 128, 320), (0, 4100, 100, 72), (1, 777, 36, 200),
                          (2, 96, 40, 7001), (0, 300, 20, 44), (2, 128, 5, 20000), (2, 200, 8, 777), (2, 64, 1, 130),
                          (0, 20000, 64, 5), (0, 777, 192, 5), (0, 300, 20, 8), (0, 130, 256, 3)):
        a, b = _operands(mode, M, N, K, g)
        assert max_rel(ops.gemm(mode, a.cuda(), b.cuda()), _ref(mode, a, b)) < TOL, (mode, M, N, K)
    wide_a = torch.randn(3000, 256, generator=g).cuda()
    w = torch.randn(96, 128, generator=g).cuda()
    wide_c = torch.zeros(3000, 320).cuda()
    a = wide_a[:, 64:192]
    ops.gemm(0, a, w, out=wide_c[:, 128:224])
    assert max_rel(wide_c[:, 128:224], a.double().cpu() @ w.double().cpu().t()) < TOL
    assert float(wide_c[:, :128].abs().max()) == 0.0 and float(wide_c[:, 224:].abs().max()) == 0.0


_TMA_ENV_SCRIPT = r'''
import sys, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import test_gpu_gemm as t
for shape in t.SHAPES:
    t.test_gemm_matches_fp64(*shape)
t.test_gemm_strided_views(); t.test_gemm_large_magnitude_spread(); t.test_gemm_tma_edges()
print('RESULT ok')
'''


@pytest.mark.parametrize('env', [{'YOLAT_TC_TMA': '0'}, {'YOLAT_TC_TMA_STORE': '0'}, {'YOLAT_TC_TAIL': '0'},
                                 {'YOLAT_TC_SKINNY': '0'}])
def test_gemm_register_loader_paths_keep_parity(env):
    """The GEMM variants behind the switches (register loaders instead of TMA loads, per-thread stores instead of TMA
    stores, no tail split) must pass the same shapes."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([sys.executable, '-c', _TMA_ENV_SCRIPT % dict(root=root, tests=os.path.join(root, 'tests'))], env=e,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and 'RESULT ok' in p.stdout, (env, p.stderr[-2000:])
