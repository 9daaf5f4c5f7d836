#!/usr/bin/env python
"""bench.py -- Bezier graphs/s, fwd + loss + bwd of YOLaT's proposal classifier, on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a engine
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

One "step" = SparseCADGCN.forward + DetectionLoss + backward over one batch of synthetic
Floorplans-shaped graphs (BASELINE.json configs[1]: batch 4 x 5 000 nodes / 20 000 edges, in_channels 5,
n_blocks 2).  At N > 1 every rank runs the same batch shape on its own seed (weak scaling) and the step
ends with one NCCL all-reduce of the flat gradient buffer.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1] / [2] / [4]: name -> (description with {g} graphs per rank, default graphs per rank)
    'floorplans': ('floorplans-shaped synthetic: batch {g} x (5000 nodes, 20000 edges), 16-node proposals', 4),
    'diagrams': ('diagrams-shaped synthetic: batch {g} x (3000 nodes, 9000 edges), 3..12-node proposals, 22 classes', 4),
    'hierarchical': ('YOLaT++-shaped synthetic: batch {g} x three-level union graph (15000 nodes, 50000 edges)', 1),
}
L2_FLUSH_BYTES = 256 << 20
# whole step = 9.0 GF and 34 MB of algorithmic traffic per floorplans graph, fwd + bwd
STEP_GF_PER_GRAPH = 9.0
STEP_MB_PER_GRAPH = 34.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='floorplans', choices=sorted(WORKLOADS))
    ap.add_argument('--graphs', type=int, default=0, help='graphs per rank per step (0 = the workload default)')
    ap.add_argument('--global-batch', type=int, default=0,
                    help='BASELINE config 4: total graphs per step, split across the ranks (strong scaling)')
    ap.add_argument('--collective', choices=('auto', 'symm', 'symm-single', 'overlap', 'single'), default='auto',
                    help='N > 1: symm = in-place symmetric-memory all-reduce (multimem / two-shot), head bucket under the backward + '
                         'remainder after it; symm-single = one such call after the backward; '
                         'overlap = NCCL head-bucket all-reduce under the backward + the remainder after it; single = one '
                         'NCCL all-reduce after the backward; auto = symm when every rank can set it up, else overlap')
    ap.add_argument('--mode', default='graph', choices=['graph', 'eager'],
                    help='graph: the step is replayed as one CUDA graph (GraphedStep); eager: launched from Python')
    ap.add_argument('--cpu-steps', type=int, default=8, help='timed steps of the cpu_baseline leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--scatter-graphs', type=int, default=64,
                    help='graphs in the > L2 K-EDGE roofline measurement (0 = use the step workload)')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit('--global-batch must be a multiple of the world size')
        args.graphs = args.global_batch // world
    if args.graphs <= 0:
        args.graphs = WORKLOADS[args.workload][1]
    args.workload_text = WORKLOADS[args.workload][0].format(g=args.graphs)
    return args


def peaks():
    """(HBM GB/s, bf16 TFLOP/s burst, bf16 TFLOP/s sustained, source)"""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return (float(d['hbm_gbs']), float(d['bf16_tflops']), float(d.get('bf16_tflops_sustained', d['bf16_tflops'])),
                'measured (MEASURED_PEAKS.json)')
    return 6650.0, 1590.0, 1400.0, 'fallback (B200_PROFILING.md)'


def make_batch(workload, graphs, seed):
    from yolat_vectorgraphicsrecognition_b200 import synth
    make, kw = synth.CONFIGS[workload]
    return make(graphs=graphs, seed=seed), synth.make_opt(**kw)


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (the reference tree itself cannot travel to the box)
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(graphs, steps, warmup, workload='floorplans'):
    from oracle import restatement as R
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch, opt = make_batch(workload, graphs, 1)
    torch.manual_seed(0)
    state = R.clone_state(arch.SparseCADGCN(opt).state_dict(), torch.float32)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        R.run_step(state, opt, batch, training=True)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return dict(value=graphs / med, unit='graphs/s', cores=cores, kind='port',
                sample='%d timed fwd+bwd steps (median) of the full step workload (%d graphs), oracle/restatement.py, '
                       'torch %s CPU fp32, %d threads' % (steps, graphs, torch.__version__, cores)), med


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    base, med = cpu_reference_run(args.graphs, max(1, min(args.steps, 10)), max(1, min(args.warmup, 3)), args.workload)
    line = {
        'impl': 'reference', 'metric': 'graphs_per_sec_fwd_bwd', 'value': base['value'], 'unit': 'graphs/s',
        'n_gpus': args.gpus, 'steps': max(1, min(args.steps, 10)), 'warmup': max(1, min(args.warmup, 3)),
        'ms_per_step': med * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload_text, 'step': 'fwd+loss+bwd', 'graphs_per_step': args.graphs},
        'cpu_baseline': base,
        'e2e': {'value': base['value'], 'unit': 'graphs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------
class Clocks(object):
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], None, set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def edge_path_bytes(N, E, Cin, C):
    """Algorithmic bytes of the edge path alone: x + edge_index (int64 as delivered) + e_attr + out."""
    return 4 * N * Cin + 16 * E + 16 * E + 4 * N * C


def ncu_traffic():
    """dram bytes (read + write) per launch of the roofline kernel from the committed `ncu --set full` capture
    (profiles/edge_fused_traffic.json, written by profiles/extract_traffic.py); None when no capture exists."""
    p = os.path.join(ROOT, 'profiles', 'edge_fused_traffic.json')
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get('dram_bytes_per_launch')


def ncu_gemm():
    """Tensor-pipe activity of the three fusion products from the committed `ncu --set full` capture
    (profiles/r2_z_gemm_ncu.json): per operand layout, the mean over the full-size launches of the capture."""
    p = os.path.join(ROOT, 'profiles', 'r2_z_gemm_ncu.json')
    if not os.path.exists(p):
        return None
    with open(p) as f:
        rep = json.load(f)
    out = {'file': 'profiles/r2_z_gemm_ncu.json'}
    for mode, name in ((0, 'NT'), (1, 'NN'), (2, 'TN')):
        ks = [k for k in rep['kernels'] if 'k_tc_gemm_ws<%d,' % mode in k['Kernel Name'] and int(k['launch__grid_size']) >= 140]
        if ks:
            out[name] = {
                'tensor_subpipe_active_pct': sum(float(k['sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active']) for k in ks) / len(ks),
                'tf32_ops_pct_of_peak': sum(float(k['sm__ops_path_tensor_src_tf32_dst_fp32.sum.pct_of_peak_sustained_elapsed']) for k in ks) / len(ks),
                'us': sum(float(k['gpu__time_duration.sum']) for k in ks) / len(ks)}
    return out


def edge_bytes(N, E, Cin, Cn, C):
    """Algorithmic bytes of one GraphConv('attr_edge_gp2') forward: x + edge_index (int64 as
    delivered) + e_attr + out, plus the node branch x_node + x_node_out."""
    return 4 * N * Cin + 16 * E + 16 * E + 4 * N * C + 4 * N * Cn + 4 * N * C


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference_arm(args)

    import ctypes as C
    import torch.distributed as dist
    from yolat_vectorgraphicsrecognition_b200 import _lib, synth, dp, ops
    from yolat_vectorgraphicsrecognition_b200.graphed import GraphedStep
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    from yolat_vectorgraphicsrecognition_b200.graph import CSRGraph
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl b200 needs a CUDA device (no CPU fallback)')
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.lib()
    hbm, tf_burst, tf_sust, peak_src = peaks()

    seed = 1 if world == 1 else 1000 + rank
    host, opt = make_batch(args.workload, args.graphs, seed)
    host = host.pin_memory()
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt).to(dev).train()
    crit = arch.DetectionLoss(opt)
    params = [p for p in model.parameters()]
    resident = host.to(dev)
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    # N > 1: gradients live in one flat buffer written by the backward kernels (no gather copy before the collective).
    # The classifier head's 5.3 MB are all-reduced from an autograd hook while the rest of the backward runs, the
    # remaining 1.1 MB after it.  In the captured step the hook issues the collective on a forked side stream (a parallel
    # branch of the step graph); --collective single reduces the whole buffer once at the end of the backward.
    sync = None
    if world > 1:
        sync = dp.OverlappedGradSync(model, overlap=args.collective in ('overlap', 'auto', 'symm'), side_stream=args.mode == 'graph',
                                     symmetric='auto' if args.collective == 'auto' else args.collective.startswith('symm'))

    def eager_step(batch):
        for p in params:
            p.grad = None
        out = model(batch, None)
        loss = crit(out, batch)['loss']
        loss.backward()
        if sync is not None:
            sync.finish()
        return loss

    # The step is ~130 short kernels: eager launching is host-bound, so the product path replays it as one CUDA
    # graph (graphed.py); at N > 1 the two NCCL all-reduces are captured into the same graph.
    graphed = GraphedStep(model, crit, extra=sync.finish if sync is not None else None)

    def graph_step(batch):
        return graphed(batch)

    step = eager_step if args.mode == 'eager' else graph_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, batch, steps):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed (untimed) in between."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in ev:
            flush.zero_()
            a.record()
            fn(batch)
            b.record()
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing: value ------------------------------------------------------------
    l0 = lib.yolat_launch_count()
    eager_step(resident)
    launches = lib.yolat_launch_count() - l0           # kernels of this library per step (the graph replays the same)
    for _ in range(max(args.warmup, 3)):
        step(resident)
    barrier()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    t_wall0 = time.perf_counter()
    ms = timed(step, resident, args.steps)
    t_wall1 = time.perf_counter()
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    other = eager_step if step is graph_step else graph_step
    for _ in range(3):
        other(resident)
    ms_other = timed(other, resident, max(3, args.steps // 2))

    # ---- N > 1: how much of the collective is exposed (same step captured without it) ----------------
    collective = None
    if sync is not None and args.mode == 'graph':
        sync.enabled = False               # the same step without any collective
        plain = GraphedStep(model, crit)
        for _ in range(3):
            plain(resident)
        ms_plain = timed(plain, resident, args.steps)
        sync.enabled = True
        how = ('torch.ops.symm_mem.%s over the symmetric flat buffer, %s' % (sync._symm_op, 'head bucket on a side stream under the '
               'backward + remainder after it' if sync.overlap else 'one kernel after the backward')) if sync.symmetric \
            else ('ncclAllReduce, schedule: %s' % ('overlap' if sync.overlap else 'single'))
        collective = {'op': 'all-reduce (mean) of the flat fp32 gradient buffer (written in place by the backward kernels), '
                            'captured in the step graph; %s' % how,
                      'requested': args.collective, 'symmetric_memory': bool(sync.symmetric), 'symm_error': sync._symm_error,
                      'bytes': 4 * sync.numel, 'overlapped_bytes': sync.overlapped_bytes,
                      'exposed_bytes': sync.exposed_bytes, 'exposed_us': (ms - ms_plain) * 1e3,
                      'ms_per_step_without_collective': ms_plain, 'views_adopted': not sync.copy_mode}

    # ---- end to end through the public API with host buffers: e2e -----------------------------------
    e2e = None
    if not args.no_e2e:
        if step is graph_step:
            # the public API a training loop uses: a collate that writes ONE pinned buffer (batch.PackedBatch), one
            # host->device copy per step, the next step's copy staged on a copy stream while this step computes
            # (GraphedStep.prefetch), the loss read back every step (train.py:286)
            from yolat_vectorgraphicsrecognition_b200.batch import PackedBatch
            packed = PackedBatch.from_batch(make_batch(args.workload, args.graphs, seed)[0])

            def run(steps):
                graphed.prefetch(packed)
                for i in range(steps):
                    loss = step(packed)                  # consumes the staged copy, replays the step
                    if i + 1 < steps:
                        graphed.prefetch(packed)         # H2D of step i + 1 overlaps the kernels of step i
                    float(loss.detach())                 # D2H read of the step's result
            h2d, how = packed.nbytes, 'one pinned packed buffer per step, staged one step ahead on a copy stream'
        else:
            def run(steps):
                for _ in range(steps):
                    loss = step(host)                    # H2D copies of the pinned host tensors happen inside the step
                    float(loss.detach())
            h2d = sum(getattr(host, n).numel() * getattr(host, n).element_size()
                      for n in ('x', 'bbox_idx', 'edge', 'bbox', 'e_attr', 'labels'))
            how = 'six pinned tensors copied inside the step'
        run(3)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(args.steps)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {'value': args.graphs * world / (float(t.item()) * 1e-3), 'unit': 'graphs/s',
               'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4, 'ms_per_step': float(t.item()), 'h2d': how}

    def prof_read(ids):
        out = {}
        for name, kid in ids:
            n, t = C.c_int64(0), C.c_double(0.0)
            lib.yolat_prof_read(kid, C.byref(n), C.byref(t))
            out[name] = {'launches': int(n.value), 'ms': (t.value / n.value) if n.value else None}
        return out

    KEDGE_IDS = (('k_edge_fused<F_Z1> (pass A)', 0), ('k_edge_fused<F_STATS> (pass B)', 1), ('k_edge_fused<F_AGG> (pass C)', 2),
                 ('k_edge_bwd<D1>', 4), ('k_edge_bwd<D2T>', 5), ('k_edge_bwd<D2S>', 6))

    # ---- rooflines (rank 0) --------------------------------------------------------------------------
    roof = roof_mlp = step_roof = None
    if rank == 0:
        # (1) K-EDGE kernels exactly as the timed step launches them (training mode, autograd on: passes A/B/C forward,
        #     D1/D2T/D2S backward), at the > L2 scale-up point of SURVEY.md 8(d) (batch 64 of the floorplans generator)
        #     with L2 flushed, and at the step's own size from one profiled eager step.
        sg = args.scatter_graphs if args.scatter_graphs > 0 else args.graphs
        big = synth.floorplans_batch(graphs=sg, seed=7).to(dev)
        Nn, Ee = big.x.shape[0], big.edge.shape[0]
        conv = model.cls_net.backbone[0].body if len(model.cls_net.backbone) else model.cls_net.head
        Cin = conv.gconv.lin_r.weight.shape[1]
        xin = torch.randn(Nn, Cin, device=dev).requires_grad_(True)
        xnode = torch.randn(Nn, Cin, device=dev).requires_grad_(True)
        go, gn = torch.randn(Nn, 64, device=dev), torch.randn(Nn, 64, device=dev)
        graph = CSRGraph(big.edge.T, Nn)

        def conv_fwd_bwd(time_it):
            flush.zero_()
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            out, xn = conv(xin, graph, None, big.e_attr, x_node=xnode)
            b.record()
            flush.zero_()
            c0 = torch.cuda.Event(enable_timing=True)
            c0.record()
            torch.autograd.backward([out, xn], [go, gn])
            c.record()
            xin.grad = xnode.grad = None
            for p in conv.parameters():
                p.grad = None
            if time_it:
                torch.cuda.synchronize()
                return a.elapsed_time(b), c0.elapsed_time(c)
            return 0.0, 0.0

        for _ in range(3):
            conv_fwd_bwd(False)
        torch.cuda.synchronize()
        lib.yolat_prof_enable(1)     # CUDA events around the K-EDGE launches, on the launching stream
        reps, tf, tb = 10, 0.0, 0.0
        for _ in range(reps):
            f, b = conv_fwd_bwd(True)
            tf += f
            tb += b
        lib.yolat_prof_enable(0)
        per_kernel = prof_read(KEDGE_IDS)
        k_ms = per_kernel['k_edge_fused<F_AGG> (pass C)']['ms']
        fwd_ms = sum(per_kernel[k]['ms'] or 0.0 for k in list(per_kernel)[:3])
        bwd_ms = sum(per_kernel[k]['ms'] or 0.0 for k in list(per_kernel)[3:])
        nbytes = edge_path_bytes(Nn, Ee, Cin, 64)
        # SURVEY.md 8(d) K-EDGE backward: forward bytes + dOut read + dX write - out not re-read
        nbytes_bwd = nbytes + 4 * Nn * Cin
        ach = nbytes / (k_ms * 1e-3) / 1e9
        op_bytes = edge_bytes(Nn, Ee, Cin, Cin, 64)
        roof = {'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm,
                'traffic': ncu_traffic(),
                'kernel': 'ef::k_edge_fused<F_AGG> (K-EDGE pass C: gather + Lin1 + BN1/ReLU + tcgen05 Lin2 + BN2/ReLU + '
                          'segmented mean-scatter) of one block-layer GraphConv(64->64) forward, training mode with '
                          'autograd recording (no tape is written: the backward recomputes)',
                'in_timed_step': True,
                'in_timed_step_note': 'the timed fwd+bwd step launches exactly these instantiations (passes A/B/C forward, '
                                      'D1/D2T/D2S backward, twice each: head + block layer); measured here through the '
                                      'same autograd path at the > L2 shape, and at the step size under at_step_size',
                'algorithmic_bytes': nbytes, 'ms': k_ms, 'peak_source': peak_src,
                'shape': {'N': Nn, 'E': Ee, 'graphs': sg}, 'l2': 'flushed before every forward and every backward',
                'timing': 'CUDA events recorded around each launch on its own stream (yolat_prof_*), avg of %d' % reps,
                'passes': per_kernel,
                'forward_all_passes': {'what': 'passes A + B + C (training-mode BatchNorm needs global statistics twice '
                                               'before any output exists)', 'ms': fwd_ms, 'algorithmic_bytes': nbytes,
                                       'achieved': nbytes / (fwd_ms * 1e-3) / 1e9, 'frac': nbytes / (fwd_ms * 1e-3) / 1e9 / hbm},
                'backward': {'what': 'eb::k_edge_bwd passes D1 + D2T + D2S (recompute, no [E,C] tensor in HBM)',
                             'ms': bwd_ms, 'algorithmic_bytes': nbytes_bwd,
                             'achieved': nbytes_bwd / (bwd_ms * 1e-3) / 1e9 if bwd_ms else None,
                             'frac': nbytes_bwd / (bwd_ms * 1e-3) / 1e9 / hbm if bwd_ms else None},
                'whole_op': {'what': 'yolat_gp2_fwd / yolat_gp2_bwd: lin_r + P/Q GEMMs, K-EDGE passes, BN finalizes, node branch',
                             'fwd_ms': tf / reps, 'bwd_ms': tb / reps, 'algorithmic_bytes_fwd': op_bytes,
                             'fwd_frac': op_bytes / (tf / reps * 1e-3) / 1e9 / hbm}}
        del big, xin, xnode, graph, go, gn
        # the same kernels inside one eager step of the timed workload (L2 flushed before the step)
        flush.zero_()
        torch.cuda.synchronize()
        lib.yolat_prof_enable(1)
        if sync is not None:
            sync.enabled = False      # rank 0 only from here on: no collective may be issued
        eager_step(resident)
        if sync is not None:
            sync.enabled = True
        torch.cuda.synchronize()
        lib.yolat_prof_enable(0)
        roof['at_step_size'] = {'what': 'per-launch averages inside one eager step of the timed workload '
                                        '(N=%d, E=%d: L2-resident, latency-bound)' % (resident.x.shape[0], resident.edge.shape[0]),
                                'passes': prof_read(KEDGE_IDS)}

        # (2) MLP path: the three products of fusion_block (torch_nn.py:58, architecture...py:40,62) on the tcgen05 GEMM
        M_, K_, N_ = resident.x.shape[0], model.cls_net.fusion_dims, 1024
        xa = torch.randn(M_, K_, device=dev)
        wa = torch.randn(N_, K_, device=dev)
        dya = torch.randn(M_, N_, device=dev)
        shapes = {'NT y = x W^T': (ops.GEMM_NT, xa, wa), 'NN dx = dy W': (ops.GEMM_NN, dya, wa), 'TN dW = dy^T x': (ops.GEMM_TN, dya, xa)}
        flops = 2.0 * M_ * K_ * N_
        peak_3xtf32 = tf_sust / 2.0 / 3.0        # kind::tf32 issues at half the bf16 rate; 3 MMAs per fp32-accurate product
        gem = {}
        for name, (mode, a_, b_) in shapes.items():
            for _ in range(3):
                ops.gemm(mode, a_, b_)
            tot = 0.0
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.gemm(mode, a_, b_)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            gms = tot / 10
            gem[name] = {'ms': gms, 'tflops': flops / (gms * 1e-3) / 1e12,
                         'frac_vs_bf16_sustained': flops / (gms * 1e-3) / 1e12 / tf_sust,
                         'frac_vs_3xtf32_peak': flops / (gms * 1e-3) / 1e12 / peak_3xtf32}
        nt = gem['NT y = x W^T']
        roof_mlp = {'bound': 'tensor', 'kernel': 'tc::k_tc_gemm_ws<MODE,128,fast,TMA> (persistent warp-specialised tcgen05 GEMM, TMEM '
                                                 'accumulators, operands by cp.async.bulk.tensor loads + in-place hi/lo split, C by TMA stores)',
                    'ncu': ncu_gemm(),
                    'shape': {'M': M_, 'K': K_, 'N': N_}, 'flops': flops, 'ms': nt['ms'], 'achieved': nt['tflops'],
                    'unit': 'TFLOP/s', 'peak': tf_sust, 'frac': nt['frac_vs_bf16_sustained'],
                    'mma': '3xTF32: three tcgen05.mma.kind::tf32 per fp32-accurate product (hi*hi + hi*lo + lo*hi)',
                    'derating': 'kind::tf32 runs at half the bf16 rate and every product is issued three times: the useful-'
                                'flop ceiling of this scheme is bf16_tflops_sustained / 6 = %.0f TFLOP/s (no measured tf32 '
                                'peak exists in MEASURED_PEAKS.json)' % peak_3xtf32,
                    'frac_vs_3xtf32_peak': nt['frac_vs_3xtf32_peak'], 'products': gem, 'peak_source': peak_src,
                    'l2': 'flushed before every call'}
        del xa, wa, dya

        # (3) whole step against both rooflines (floorplans figures)
        gps = args.graphs * world / (ms * 1e-3)
        if args.workload == 'floorplans':
            step_roof = {'graphs_per_s': gps, 'gflop_per_graph': STEP_GF_PER_GRAPH, 'mb_per_graph': STEP_MB_PER_GRAPH,
                         'compute_roofline_graphs_per_s': tf_sust * 1e3 * world / STEP_GF_PER_GRAPH,
                         'compute_frac': gps * STEP_GF_PER_GRAPH / (tf_sust * 1e3 * world),
                         'compute_frac_vs_3xtf32_peak': gps * STEP_GF_PER_GRAPH / (peak_3xtf32 * 1e3 * world),
                         'hbm_roofline_graphs_per_s': hbm * 1e3 * world / STEP_MB_PER_GRAPH,
                         'hbm_frac': gps * STEP_MB_PER_GRAPH / (hbm * 1e3 * world),
                         'note': 'at this size the working set is L2-resident and the step is a chain of short kernels: '
                                 'neither roofline binds (SURVEY.md H4)'}

    # ---- CPU baseline (rank 0, N = 1) -------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference_run(args.graphs, args.cpu_steps, 2, args.workload)

    if rank == 0:
        line = {
            'metric': 'graphs_per_sec_fwd_bwd', 'value': args.graphs * world / (ms * 1e-3), 'unit': 'graphs/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'strong' if args.global_batch else 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': args.workload_text,
                       'step': 'fwd+loss+bwd' + ('+nccl grad all-reduce' if world > 1 else ''),
                       'graphs_per_step_per_gpu': args.graphs, 'global_batch': args.graphs * world,
                       'in_channels': opt.in_channels, 'n_blocks': opt.n_blocks, 'n_filters': opt.n_filters,
                       'n_classes': opt.n_classes, 'l2': 'flushed between timed steps (256 MiB memset, untimed)',
                       'launch': 'one CUDA-graph replay per step' if args.mode == 'graph' else 'eager',
                       'parallelism': 'dp%d' % world},
            'clocks': clk, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roof, 'roofline_mlp': roof_mlp,
            'step_roofline': step_roof, 'collective': collective, 'cpu_baseline': cpu,
            'launch_mode': args.mode,
            'other_mode': {'mode': 'eager' if args.mode == 'graph' else 'graph', 'ms_per_step': ms_other},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down order matters: the captured NCCL kernels keep the communicator referenced until their graphs are
        # destroyed.  The line above is already out; if the communicator still refuses to go, leave without it.
        import threading
        sys.stdout.flush()
        sys.stderr.flush()
        bail = threading.Timer(20.0, lambda: os._exit(0))
        bail.daemon = True
        bail.start()
        barrier()
        graphed.release()
        if sync is not None:
            sync.close()
        dist.destroy_process_group()
        bail.cancel()


if __name__ == '__main__':
    main()
