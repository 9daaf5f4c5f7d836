#!/usr/bin/env python
"""Discrete-event model of the K-EDGE tile pipeline (edge_fused.cu) for planning: given the per-tile busy times of the
roles measured with tools/ef_trace.py it predicts the steady-state period of the three MMA-issue modes, so a design
change can be sized before it is written.  Pure Python, no GPU.

    python tools/ef_pipeline_model.py                 # the measured round-1 numbers (F_AGG and F_STATS)
    python tools/ef_pipeline_model.py --gather 2624 --issue 1364 --drain 3060

Resources: two a1 operand stages, two TMEM accumulators.  Per tile t:
    gather(t)  needs a1 stage t&1 free (= MMA(t-2) complete)                    busy G per gather warp
    MMA(t)     needs gather(t) of ALL warps + accumulator t&1 free (= drain(t-2)) blocks its issuer for I cycles
    drain(t)   needs MMA(t) complete                                            busy D
Mode 0: epilogue warp 0 issues MMA(t+1) before it drains tile t (the drain team waits for it at its first barrier).
Mode 1: gather warp t % W issues MMA(t) after its own arrival, then starts gathering tile t+1 (so it arrives late there).
Mode 2: a control warp issues; gather and drain never block on the issue.
The model reproduces the measured ordering of the modes; its absolute periods are 20-35 % optimistic (mbarrier hand-off
latencies, the ring and the slower 96-register epilogue of mode 2 are not modelled).
"""
import argparse


def simulate(mode, G, I, D, tiles=200, warps=8):
    g_done = [[0.0] * warps for _ in range(tiles)]      # arrival of each gather warp for tile t
    m_start = [0.0] * tiles
    m_done = [0.0] * tiles
    d_done = [0.0] * tiles
    free_at = [0.0] * warps                             # when each gather warp can start its next tile
    e_free = 0.0                                        # epilogue team
    c_free = 0.0                                        # control warp
    for t in range(tiles):
        stage_free = m_done[t - 2] if t >= 2 else 0.0
        for w in range(warps):
            g_done[t][w] = max(free_at[w], stage_free) + G
            free_at[w] = g_done[t][w]
        a_full = max(g_done[t])
        acc_free = d_done[t - 2] if t >= 2 else 0.0
        if mode == 1:
            w = t % warps
            m_start[t] = max(a_full, acc_free)
            m_done[t] = m_start[t] + I
            free_at[w] = m_done[t]                      # the issuer gathers its next tile only after the issue returns
            d_start = max(m_done[t], e_free)
            d_done[t] = d_start + D
            e_free = d_done[t]
        elif mode == 2:
            m_start[t] = max(a_full, acc_free, c_free)
            m_done[t] = m_start[t] + I
            c_free = m_done[t]
            d_start = max(m_done[t], e_free)
            d_done[t] = d_start + D
            e_free = d_done[t]
        else:
            # the epilogue team issues MMA(t) when it is free (after draining tile t-2 ... t-1 in program order)
            m_start[t] = max(a_full, acc_free, e_free)
            m_done[t] = m_start[t] + I
            e_free = m_done[t]
            if t >= 1:                                   # then drains the previous tile
                d_start = max(m_done[t - 1], e_free)
                d_done[t - 1] = d_start + D
                e_free = d_done[t - 1]
    last = d_done[tiles - 2] if mode == 0 else d_done[tiles - 1]
    first = d_done[tiles // 2 - 2] if mode == 0 else d_done[tiles // 2 - 1]
    return (last - first) / (tiles - tiles // 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gather', type=float, default=None)
    ap.add_argument('--issue', type=float, default=None)
    ap.add_argument('--drain', type=float, default=None)
    ap.add_argument('--tiles-per-cta', type=float, default=67.6, help='10 000 tiles over 148 CTAs at E = 1 280 000')
    ap.add_argument('--ghz', type=float, default=1.9)
    args = ap.parse_args()
    cases = [('F_AGG  (measured: 237 / 180 / 170 us)', 2624.0, 1364.0, 3060.0),
             ('F_STATS (measured: 144 / 155 / 127 us)', 2624.0, 1364.0, 1300.0)]
    if args.gather is not None:
        cases = [('custom', args.gather, args.issue or 1364.0, args.drain or 3060.0)]
    for name, G, I, D in cases:
        print('%s   gather %.0f  issue %.0f  drain %.0f cycles per tile' % (name, G, I, D))
        for mode, label in ((0, 'epilogue warp 0 issues'), (1, 'rotating gather warp issues'), (2, 'control warp issues')):
            p = simulate(mode, G, I, D)
            print('   mode %d (%-28s): period %5.0f cycles  ->  %5.0f us per pass' %
                  (mode, label, p, p * args.tiles_per_cta / (args.ghz * 1e3)))


if __name__ == '__main__':
    main()
