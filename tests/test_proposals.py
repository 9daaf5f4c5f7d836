"""Proposal enumeration (reference Datasets/graph_dict3.py:309-789).

CPU (`-m "not gpu"`):
  * the oracle restatement against goldens produced by the UNMODIFIED reference (tests/golden/proposals_ref.pkl,
    oracle/make_golden_proposals.py), and against the live reference when /root/reference is present;
  * the kernels' logic: csrc/proposals.cu compiled for the host against tests/emu/cuda_emu.h (one thread per CTA) and run
    through the same C entry points, against the oracle -- same proposal order, bit-exact outputs, statistics 1e-12;
  * the error conditions the reference raises on.
GPU (`-m gpu`): the real kernels through `proposals.get_proposal` against the oracle and the goldens.
"""
import ctypes as C
import os
import pickle
import shutil
import subprocess

import numpy as np
import pytest

from oracle import proposals as OP
from yolat_vectorgraphicsrecognition_b200 import proposals as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'proposals_ref.pkl')
EXACT = ('pos', 'is_super', 'edge', 'edge_super', 'e_attr', 'e_attr_super', 'bbox_idx', 'labels', 'has_obj', 'bbox',
         'bbox_targets', 'roots', 'per_cc', 'is_control')
# stat_feats columns that are counts / extrema / box sides (exact) vs means / standard deviations (summation order)
STAT_EXACT = [0, 1, 2, 3, 4, 5, 6, 8, 9]
STAT_TOL = [7, 10, 11, 12]


def assert_same_canon(a, b, what):
    for k in EXACT:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape, '%s: %s shape %s vs %s' % (what, k, x.shape, y.shape)
        assert np.array_equal(x.astype(np.float64), y.astype(np.float64)), '%s: %s differs' % (what, k)
    sa, sb = a['stat_feats'], b['stat_feats']
    assert sa.shape == sb.shape
    assert np.array_equal(sa[:, STAT_EXACT], sb[:, STAT_EXACT]), '%s: exact statistics differ' % what
    np.testing.assert_allclose(sa[:, STAT_TOL], sb[:, STAT_TOL], rtol=1e-10, atol=1e-12, err_msg=what)


def assert_same_result(r, o, what):
    """Two 14-tuples with the SAME proposal order: everything exact but the four mean / std statistics."""
    names = ('pos', 'is_super', 'is_control', 'edge', 'edge_super', 'e_attr', 'e_attr_super', 'labels', 'bbox_idx', 'bbox',
             'bbox_targets', 'stat_feats', 'has_obj')
    for i, k in enumerate(names):
        x, y = np.asarray(r[i]), np.asarray(o[i])
        assert x.shape == y.shape, '%s: %s shape %s vs %s' % (what, k, x.shape, y.shape)
        if k == 'stat_feats':
            assert np.array_equal(x[:, STAT_EXACT], y[:, STAT_EXACT]), '%s: exact statistics differ' % what
            np.testing.assert_allclose(x[:, STAT_TOL], y[:, STAT_TOL], rtol=1e-10, atol=1e-12, err_msg=what)
        else:
            assert np.array_equal(x.astype(np.float64), y.astype(np.float64)), '%s: %s differs' % (what, k)
    assert r[1].dtype == o[1].dtype and r[3].dtype == np.int64 and r[8].dtype == np.int64
    assert isinstance(r[7], list) and isinstance(r[12], list)
    ra, oa = r[13], o[13]
    assert len(ra) == len(oa)
    for x, y in zip(ra, oa):
        assert x.value == y.value, what
        assert [c.value for c in x.children] == [c.value for c in y.children], what


def golden_cases():
    with open(GOLDEN, 'rb') as f:
        return pickle.load(f)


# ---------------------------------------------------------------- oracle vs the unmodified reference
def test_oracle_matches_reference_goldens():
    for case in golden_cases():
        gd, gt_bbox, gt_labels = OP.synth_graph_dict(case['seed'], **case['kw'])
        got = OP.canonical_order(OP.get_proposal(gd, gt_bbox, gt_labels, case['step'], case['n_classes'], True))
        assert_same_canon(got, case['canon'], 'oracle vs reference golden, seed %d' % case['seed'])


@pytest.mark.skipif(not os.path.isdir('/root/reference/Datasets'), reason='reference tree not on this machine')
def test_oracle_matches_live_reference():
    from oracle import make_golden_proposals as MG
    cls = MG.reference_class()
    for seed, normalize, extra in ((11, True, {}), (12, False, {}), (13, True, {}), (14, True, dict(jitter=0.8)),
                                   (15, True, dict(dup_points=0.3)), (16, False, dict(jitter=0.4, dup_points=0.2))):
        gd, gt_bbox, gt_labels = OP.synth_graph_dict(seed, n_cc=5, max_nodes=12, grid=5, **extra)
        ref = OP.canonical_order(MG.run_reference(cls, gd, gt_bbox, gt_labels, 5, 17, normalize))
        got = OP.canonical_order(OP.get_proposal(gd, gt_bbox, gt_labels, 5, 17, normalize))
        assert_same_canon(got, ref, 'oracle vs live reference, seed %d' % seed)


# ---------------------------------------------------------------- the kernels' logic, compiled for the host
def build_emu(so, threads=1, extra=()):
    src = os.path.join(ROOT, 'yolat_vectorgraphicsrecognition_b200', 'csrc', 'proposals.cu')
    subprocess.check_call(['g++', '-O2', '-g', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-pthread', '-x', 'c++',
                           '-DYOLAT_HOST_EMU', '-DEMU_THREADS=%d' % threads, '-I' + os.path.join(ROOT, 'tests', 'emu'),
                           src, '-o', so, '-Wno-unused-variable'] + list(extra))
    return so


def bind_emu(so):
    lib = C.CDLL(so)
    lib.yolat_proposals_ws_bytes.restype = C.c_int64
    lib.yolat_proposals_ws_bytes.argtypes = [C.POINTER(P.ProposalIn)]
    lib.yolat_proposals_count.argtypes = [C.POINTER(P.ProposalIn), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.yolat_proposals_fill.argtypes = [C.POINTER(P.ProposalIn), C.c_void_p, C.c_int64, C.POINTER(P.ProposalOut),
                                         C.c_void_p]
    return lib


@pytest.fixture(scope='module', params=[1, 8, 64], ids=['1thread', '8threads', '64threads'])
def emu(request, tmp_path_factory):
    """csrc/proposals.cu compiled for the host: one thread per CTA; 8 real threads per CTA (one group); 64 real threads
    per CTA (two "warps" of 32, each candidate handled by one of them) with CTA and group barriers."""
    if shutil.which('g++') is None:
        pytest.skip('g++ not available')
    so = str(tmp_path_factory.mktemp('emu') / ('libprop_emu%d.so' % request.param))
    return bind_emu(build_emu(so, request.param))


def run_emu(lib, gd, gt_bbox, gt_labels, step, n_classes, normalize=True):
    """The product's host logic (pack -> count -> sizes -> fill -> unpack) over host memory and the emulated kernels."""
    p = P.pack_graph_dict(gd, gt_bbox, gt_labels)
    s_in = P.fill_in_struct(p, lambda name: p[name].ctypes.data, step, n_classes, normalize)
    ws_bytes = lib.yolat_proposals_ws_bytes(C.byref(s_in))
    assert ws_bytes > 0
    ws = np.full(ws_bytes + 64, 0xCD, dtype=np.uint8)          # poisoned: nothing may rely on zeroed scratch
    totals = np.zeros(P.N_TOTALS, dtype=np.int64)
    assert lib.yolat_proposals_count(C.byref(s_in), ws.ctypes.data, ws_bytes, totals.ctypes.data, None) == 0
    assert (ws[ws_bytes:] == 0xCD).all(), 'workspace overrun'
    P.raise_for(totals)
    outs = {name: np.empty(shape, dtype=dt)
            for name, dt, shape in P.output_specs(totals, s_in.A, s_in.As, s_in.ncc)}
    s_out = P.ProposalOut()
    for name, a in outs.items():
        setattr(s_out, name, a.ctypes.data)
    assert lib.yolat_proposals_fill(C.byref(s_in), ws.ctypes.data, ws_bytes, C.byref(s_out), None) == 0
    assert (ws[ws_bytes:] == 0xCD).all(), 'workspace overrun'
    return P.unpack_outputs(outs, p['is_super_dtype'])


def test_emulated_kernels_match_oracle(emu):
    for case in golden_cases():
        gd, gt_bbox, gt_labels = OP.synth_graph_dict(case['seed'], **case['kw'])
        got = run_emu(emu, gd, gt_bbox, gt_labels, case['step'], case['n_classes'])
        want = OP.get_proposal(gd, gt_bbox, gt_labels, case['step'], case['n_classes'], True)
        assert_same_result(got, want, 'emulated kernels vs oracle, seed %d' % case['seed'])
        assert_same_canon(OP.canonical_order(got), case['canon'], 'emulated kernels vs reference golden')


def test_emulated_kernels_random_cases(emu):
    rng = np.random.RandomState(7)
    for seed in range(100, 124):
        kw = dict(n_cc=int(rng.randint(1, 8)), max_nodes=int(rng.randint(3, 30)), grid=int(rng.randint(2, 10)),
                  with_control=bool(rng.rand() < 0.7))
        step = int(rng.choice([1, 2, 3, 5, 8]))
        normalize = bool(rng.rand() < 0.5)
        gd, gt_bbox, gt_labels = OP.synth_graph_dict(seed, **kw)
        try:
            want = OP.get_proposal(gd, gt_bbox, gt_labels, step, 17, normalize)
        except Exception as e:                                 # same condition -> same exception type
            with pytest.raises(type(e)):
                run_emu(emu, gd, gt_bbox, gt_labels, step, 17, normalize)
            continue
        got = run_emu(emu, gd, gt_bbox, gt_labels, step, 17, normalize)
        assert_same_result(got, want, 'seed %d %r step %d' % (seed, kw, step))


def test_emulated_kernels_off_lattice_duplicates_and_scale(emu):
    """Arbitrary doubles (every coordinate value distinct), nodes sitting on each other, a larger sampling grid, and a
    floor-plan-sized image with one large component."""
    cases = [(200, dict(n_cc=5, max_nodes=16, grid=6, jitter=0.8), 5), (201, dict(n_cc=4, max_nodes=20, grid=5, dup_points=0.3), 5),
             (202, dict(n_cc=3, max_nodes=30, grid=8, jitter=0.5, dup_points=0.2), 10),
             (203, dict(n_cc=6, max_nodes=12, grid=4, jitter=1.0), 1), (204, dict(n_cc=2, max_nodes=60, grid=10, jitter=0.3), 16)]
    for seed, kw, step in cases:
        gd, gt_bbox, gt_labels = OP.synth_graph_dict(seed, **kw)
        want = OP.get_proposal(gd, gt_bbox, gt_labels, step, 17, True)
        assert_same_result(run_emu(emu, gd, gt_bbox, gt_labels, step, 17), want, 'seed %d %r step %d' % (seed, kw, step))
    gd, gt_bbox, gt_labels = floorplan_scale_image()
    assert_same_result(run_emu(emu, gd, gt_bbox, gt_labels, 5, 17), OP.get_proposal(gd, gt_bbox, gt_labels, 5, 17, True),
                       'floor-plan scale')


def floorplan_scale_image(n_cc=120):
    """Hundreds of small components plus one large one (80 nodes on a 12 x 12 lattice) in a corner of the image."""
    gd, gt_bbox, gt_labels = OP.synth_graph_dict(42, n_cc=n_cc, max_nodes=24, grid=9)
    big, gtb, gtl = OP.synth_graph_dict(43, n_cc=1, max_nodes=80, grid=12)
    n0 = gd['pos']['spatial'].shape[0]
    gd['pos']['spatial'] = np.concatenate([gd['pos']['spatial'], big['pos']['spatial'] * 0.05 + 0.94])
    gd['cc'].append([i + n0 for i in big['cc'][0]])
    for k in ('shape', 'super'):
        gd['edge'][k] = np.concatenate([gd['edge'][k], big['edge'][k] + n0])
        gd['edge_attr'][k] = np.concatenate([gd['edge_attr'][k], big['edge_attr'][k]])
    for k in ('is_super', 'is_control'):
        gd['attr'][k] = np.concatenate([gd['attr'][k], big['attr'][k]])
    return gd, np.concatenate([gt_bbox, gtb * 0.05 + 0.94]), np.concatenate([gt_labels, gtl])


def _tiny():
    """One square with a diagonal: 4 nodes, 5 edges."""
    gd = {'cc': [[0, 1, 2, 3]],
          'pos': {'spatial': np.array([[0.1, 0.1], [0.3, 0.1], [0.3, 0.4], [0.1, 0.4]])},
          'edge': {'shape': np.array([[0, 1], [1, 2], [2, 3], [3, 0], [0, 2]]), 'super': np.array([[0, 2]])},
          'edge_attr': {'shape': np.arange(30, dtype=np.float64).reshape(5, 6), 'super': np.zeros((1, 6))},
          'attr': {'is_super': np.zeros((4, 1)), 'is_control': np.zeros((4, 1))}}
    return gd, np.array([[0.1, 0.1, 0.3, 0.4]]), np.array([3])


def test_error_conditions_match_reference(emu):
    gd, gt, gl = _tiny()
    got = run_emu(emu, gd, gt, gl, 5, 17)
    assert_same_result(got, OP.get_proposal(gd, gt, gl, 5, 17, True), 'tiny')
    assert 3 in got[7]                                         # the whole square matches the ground-truth box
    # a component without y extent: np.arange(lo, lo, 0.0) on numpy scalars -> ValueError('cannot compute length')
    flat = dict(gd, pos={'spatial': np.array([[0.1, 0.1], [0.3, 0.1], [0.4, 0.1], [0.2, 0.1]])})
    for fn in (lambda: OP.get_proposal(flat, gt, gl, 5, 17, True), lambda: run_emu(emu, flat, gt, gl, 5, 17)):
        with pytest.raises(ValueError, match='cannot compute length'):
            fn()
    # no ground-truth box touches the component -> SystemExit
    far = np.array([[0.8, 0.8, 0.9, 0.9]])
    for fn in (lambda: OP.get_proposal(gd, far, gl, 5, 17, True), lambda: run_emu(emu, gd, far, gl, 5, 17)):
        with pytest.raises(SystemExit):
            fn()
    # an edge that names a control point -> KeyError
    ctrl = dict(gd, attr={'is_super': np.zeros((4, 1)), 'is_control': np.array([[0], [0], [0], [1]])}, cc=[[0, 1, 2]])
    for fn in (lambda: OP.get_proposal(ctrl, gt, gl, 5, 17, True), lambda: run_emu(emu, ctrl, gt, gl, 5, 17)):
        with pytest.raises(KeyError):
            fn()
    # a component whose sets all fail the filters (two nodes, one edge: no angle) -> ValueError from the empty argmax
    two = {'cc': [[0, 1]], 'pos': {'spatial': np.array([[0.1, 0.1], [0.3, 0.4]])},
           'edge': {'shape': np.array([[0, 1]]), 'super': np.array([[0, 1]])},
           'edge_attr': {'shape': np.zeros((1, 6)), 'super': np.zeros((1, 6))},
           'attr': {'is_super': np.zeros((2, 1)), 'is_control': np.zeros((2, 1))}}
    for fn in (lambda: OP.get_proposal(two, gt, gl, 5, 17, True), lambda: run_emu(emu, two, gt, gl, 5, 17)):
        with pytest.raises((ValueError, IndexError)):
            fn()


def test_emulated_kernels_are_race_free(tmp_path):
    """ThreadSanitizer over the 64-threads-per-CTA emulation (two groups of 32): every shared / global access of a CTA
    must be ordered by a CTA or group barrier (a missing __syncthreads() / __syncwarp() is reported as a data race)."""
    if shutil.which('g++') is None:
        pytest.skip('g++ not available')
    tsan = subprocess.run(['g++', '-print-file-name=libtsan.so'], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(tsan) or not os.path.exists(tsan):
        pytest.skip('libtsan not available')
    so = build_emu(str(tmp_path / 'libprop_tsan.so'), 64, ['-fsanitize=thread', '-O1'])
    env = dict(os.environ, LD_PRELOAD=tsan, TSAN_OPTIONS='report_signal_unsafe=0 exitcode=66')
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'emu', 'tsan_driver.py'), so, '0', '1'], env=env,
                       capture_output=True, text=True, timeout=900)
    assert 'ThreadSanitizer: data race' not in r.stderr, r.stderr[-4000:]
    if r.returncode != 0 and 'Error' not in r.stderr and 'seed' not in r.stdout:
        pytest.skip('ThreadSanitizer runtime did not start here: %s' % r.stderr[-300:])   # e.g. unsupported mmap layout
    assert 'ThreadSanitizer' not in r.stderr, r.stderr[-4000:]
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert r.stdout.count(' ok: ') == 2


def test_enumerator_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    gd, gt, gl = _tiny()
    with pytest.raises(Exception) as e:
        P.ProposalEnumerator(17)._get_proposal(gd, gt, gl, bbox_sampling_step=5)
    assert 'CPU' in str(e.value) or 'missing' in str(e.value)
    with pytest.raises(NotImplementedError):
        P.ProposalEnumerator(17, do_mixup=True)._get_proposal(gd, gt, gl, bbox_sampling_step=5)


@pytest.mark.skipif(not os.path.isdir('/root/reference/Datasets'), reason='reference tree not on this machine')
def test_overlay_binds_to_the_reference_dataset_class(monkeypatch):
    """the one-line overlay on the UNMODIFIED `SESYDFloorPlan` routes its `_get_proposal` call
    (graph_dict3.py:929) here -- reading n_classes / normalize_bbox / do_mixup from the Dataset instance."""
    import torch
    from oracle import make_golden_proposals as MG
    cls = MG.reference_class()
    monkeypatch.setattr(cls, '_get_proposal', P.ProposalEnumerator._get_proposal)
    ds = cls.__new__(cls)
    ds.n_classes, ds.normalize_bbox, ds.do_mixup = 17, True, True
    gd, gt, gl = _tiny()
    with pytest.raises(NotImplementedError):                   # mixup is random augmentation, not on this path
        ds._get_proposal(gd, gt, gl, bbox_sampling_step=5)
    ds.do_mixup = False
    seen = {}
    monkeypatch.setattr(P, 'get_proposal', lambda *a: seen.setdefault('args', a) and None)
    ds._get_proposal(gd, gt, gl, bbox_sampling_step=5)
    assert seen['args'][3:] == (5, 17, True) and seen['args'][0] is gd
    monkeypatch.undo()
    if not torch.cuda.is_available():
        monkeypatch.setattr(cls, '_get_proposal', P.ProposalEnumerator._get_proposal)
        with pytest.raises(Exception, match='CPU|missing'):    # no silent fallback to the python loops
            ds._get_proposal(gd, gt, gl, bbox_sampling_step=5)


# ---------------------------------------------------------------- the real kernels
@pytest.mark.gpu
def test_gpu_proposals_match_goldens_and_oracle():
    for case in golden_cases():
        gd, gt_bbox, gt_labels = OP.synth_graph_dict(case['seed'], **case['kw'])
        got = P.get_proposal(gd, gt_bbox, gt_labels, case['step'], case['n_classes'], True)
        assert_same_canon(OP.canonical_order(got), case['canon'], 'GPU vs reference golden, seed %d' % case['seed'])
        want = OP.get_proposal(gd, gt_bbox, gt_labels, case['step'], case['n_classes'], True)
        assert_same_result(got, want, 'GPU vs oracle, seed %d' % case['seed'])


@pytest.mark.gpu
def test_gpu_proposals_random_and_errors():
    rng = np.random.RandomState(7)
    for seed in range(100, 116):
        kw = dict(n_cc=int(rng.randint(1, 8)), max_nodes=int(rng.randint(3, 30)), grid=int(rng.randint(2, 10)),
                  with_control=bool(rng.rand() < 0.7))
        step = int(rng.choice([1, 2, 3, 5, 8]))
        normalize = bool(rng.rand() < 0.5)
        gd, gt_bbox, gt_labels = OP.synth_graph_dict(seed, **kw)
        try:
            want = OP.get_proposal(gd, gt_bbox, gt_labels, step, 17, normalize)
        except Exception as e:
            with pytest.raises(type(e)):
                P.get_proposal(gd, gt_bbox, gt_labels, step, 17, normalize)
            continue
        got = P.get_proposal(gd, gt_bbox, gt_labels, step, 17, normalize)
        assert_same_result(got, want, 'seed %d %r step %d' % (seed, kw, step))
    for seed, kw, step in [(200, dict(n_cc=5, max_nodes=16, grid=6, jitter=0.8), 5),
                           (202, dict(n_cc=3, max_nodes=30, grid=8, jitter=0.5, dup_points=0.2), 10),
                           (204, dict(n_cc=2, max_nodes=60, grid=10, jitter=0.3), 16)]:
        gd, gt_bbox, gt_labels = OP.synth_graph_dict(seed, **kw)
        assert_same_result(P.get_proposal(gd, gt_bbox, gt_labels, step, 17, True),
                           OP.get_proposal(gd, gt_bbox, gt_labels, step, 17, True), 'seed %d' % seed)
    gd, gt, gl = _tiny()
    flat = dict(gd, pos={'spatial': np.array([[0.1, 0.1], [0.3, 0.1], [0.4, 0.1], [0.2, 0.1]])})
    with pytest.raises(ValueError, match='cannot compute length'):
        P.get_proposal(flat, gt, gl, 5, 17)
    with pytest.raises(SystemExit):
        P.get_proposal(gd, np.array([[0.8, 0.8, 0.9, 0.9]]), gl, 5, 17)


@pytest.mark.gpu
def test_gpu_proposals_floorplan_scale():
    """A floor-plan-sized image (hundreds of components, one of them large): against the oracle, and run twice
    (bit-identical outputs: the atomics only order scratch lists that are sorted afterwards)."""
    gd, gt_bbox, gt_labels = floorplan_scale_image()
    a = P.get_proposal(gd, gt_bbox, gt_labels, 5, 17, True)
    b = P.get_proposal(gd, gt_bbox, gt_labels, 5, 17, True)
    for x, y in zip(a[:13], b[:13]):
        assert np.array_equal(np.asarray(x), np.asarray(y))
    assert_same_result(a, OP.get_proposal(gd, gt_bbox, gt_labels, 5, 17, True), 'floor-plan scale')
