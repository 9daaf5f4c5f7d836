"""CPU: the C-ABI library loads and exports every symbol include/yolat_b200.h declares; the size-query
entry points (pure host code) answer without a GPU; argument validation returns error codes."""
import ctypes
import os
import re

import pytest

from util import ROOT


def _declared():
    src = open(os.path.join(ROOT, 'include', 'yolat_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(yolat_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    from yolat_vectorgraphicsrecognition_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from yolat_vectorgraphicsrecognition_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), 'libyolat_b200.so does not export %s' % n
    assert sorted(_lib.exported_symbols()) == names, 'ctypes binding table and header disagree'


def test_host_side_queries(lib):
    assert lib.yolat_abi_version() == 1
    assert lib.yolat_status_string(0) == b'ok' and b'workspace' in lib.yolat_status_string(-2)
    N, E = 20000, 80000
    assert lib.yolat_graph_ints(N, E) >= 2 * (N + 1) + 4 * E + N
    assert lib.yolat_segments_ints(N, 1250) >= N + 1250
    tape = lib.yolat_gp2_tape_floats(N, E, 64, 64, 64)
    assert tape >= 2 * E * 64 + N * 64
    assert lib.yolat_gp2_fwd_ws_floats(N, E, 64, 64, 64) > 0 and lib.yolat_gp2_bwd_ws_floats(N, E, 64, 64, 64) > 0
    assert lib.yolat_gp2_fwd_ws_floats(N, E, 64, 64, 48) < 0          # unsupported channel count
    assert lib.yolat_mlp_tape_floats(1250, 2304, 512, 1 | 2 | 4) >= 1250 * 512
    assert lib.yolat_mlp_tape_floats(1250, 256, 17, 0) == 0
    assert lib.yolat_fusemax_tape_floats(N, 128, 1024, 1250) >= 1250 * 1152
    assert lib.yolat_launch_count() == 0


def test_null_arguments_are_rejected(lib):
    assert lib.yolat_graph_build(None, 2, 1, 10, 5, None, None) == -1
    assert lib.yolat_segments_build(None, 10, 2, None, None) == -1
    assert lib.yolat_mlp_fwd(None, 8, 4, 8, None, None, 8, None, 0, None, 8, None, 0, None, 0, None) == -1
    assert lib.yolat_softmax_xent_fwd(None, 17, 4, 17, None, None, None, None, 0, None) == -1


def test_proposals_host_side(lib):
    """Proposal enumeration (include/yolat_b200.h, last block): the workspace query is pure host code, the entry points
    validate their arguments before touching the device."""
    from yolat_vectorgraphicsrecognition_b200 import proposals as P
    s = P.ProposalIn()
    s.n_all, s.ncc, s.cc_total, s.E, s.A, s.Es, s.As, s.G = 5000, 300, 4000, 6000, 6, 500, 6, 40
    s.sampling_step, s.n_classes, s.normalize_bbox = 5, 17, 1
    small = lib.yolat_proposals_ws_bytes(ctypes.byref(s))
    assert small > 300 * 441 * 8 * 13                                  # at least the per-candidate statistics
    s.sampling_step = 10
    assert lib.yolat_proposals_ws_bytes(ctypes.byref(s)) > small
    s.sampling_step = 63                                               # more grid lines than the kernels hold
    assert lib.yolat_proposals_ws_bytes(ctypes.byref(s)) == -1
    s.sampling_step = 0
    assert lib.yolat_proposals_ws_bytes(ctypes.byref(s)) == -1
    s.sampling_step = 5
    assert lib.yolat_proposals_count(ctypes.byref(s), None, 0, None, None) == -1
    assert lib.yolat_proposals_fill(ctypes.byref(s), None, 0, None, None) == -1
    assert lib.yolat_proposals_count(None, None, 0, None, None) == -1
