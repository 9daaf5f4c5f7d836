"""TEST INFRASTRUCTURE: runs the host-emulated proposal kernels (tests/emu/cuda_emu.h, EMU_THREADS > 1) from a process
that has libtsan preloaded, on a few synthetic images, and checks them against the oracle.  Launched by
tests/test_proposals.py::test_emulated_kernels_are_race_free; ThreadSanitizer reports go to stderr."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]


def main(so, seeds):
    from oracle import proposals as OP
    import test_proposals as T
    lib = T.bind_emu(so)
    for seed in seeds:
        gd, gb, gl = OP.synth_graph_dict(seed, n_cc=4, max_nodes=14, grid=5)
        got = T.run_emu(lib, gd, gb, gl, 5, 17)
        T.assert_same_result(got, OP.get_proposal(gd, gb, gl, 5, 17, True), 'seed %d' % seed)
        print('seed %d ok: %d proposals' % (seed, len(got[7])))


if __name__ == '__main__':
    main(sys.argv[1], [int(a) for a in sys.argv[2:]])
