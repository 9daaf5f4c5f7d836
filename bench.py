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

WORKLOAD = 'floorplans-shaped synthetic: batch 4 x (5000 nodes, 20000 edges), 16-node proposals (B=1250)'
L2_FLUSH_BYTES = 256 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--graphs', type=int, default=4, help='graphs per rank per step')
    ap.add_argument('--mode', default='graph', choices=['graph', 'eager'],
                    help='graph: the step is replayed as one CUDA graph (GraphedStep); eager: launched from Python')
    ap.add_argument('--cpu-steps', type=int, default=8, help='timed steps of the cpu_baseline leg')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--scatter-graphs', type=int, default=64,
                    help='graphs in the > L2 K-EDGE roofline measurement (0 = use the step workload)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (the reference tree itself cannot travel to the box)
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(graphs, steps, warmup):
    from oracle import restatement as R
    from yolat_vectorgraphicsrecognition_b200 import synth
    from yolat_vectorgraphicsrecognition_b200 import architecture3cc_rpn_gp_iter2 as arch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    state = R.clone_state(arch.SparseCADGCN(opt).state_dict(), torch.float32)
    batch = synth.floorplans_batch(graphs=graphs, seed=1)
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
    base, med = cpu_reference_run(args.graphs, max(1, min(args.steps, 10)), max(1, min(args.warmup, 3)))
    line = {
        'impl': 'reference', 'metric': 'graphs_per_sec_fwd_bwd', 'value': base['value'], 'unit': 'graphs/s',
        'n_gpus': args.gpus, 'steps': max(1, min(args.steps, 10)), 'warmup': max(1, min(args.warmup, 3)),
        'ms_per_step': med * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'step': 'fwd+loss+bwd', 'graphs_per_step': args.graphs},
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


def edge_bytes(N, E, Cin, Cn, C):
    """Algorithmic bytes of one GraphConv('attr_edge_gp2') forward: x + edge_index (int64 as
    delivered) + e_attr + out, plus the node branch x_node + x_node_out."""
    return 4 * N * Cin + 16 * E + 16 * E + 4 * N * C + 4 * N * Cn + 4 * N * C


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference_arm(args)

    import torch.distributed as dist
    from yolat_vectorgraphicsrecognition_b200 import _lib, synth, dp
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

    opt = synth.make_opt(n_classes=17)
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt).to(dev).train()
    crit = arch.DetectionLoss(opt)
    params = [p for p in model.parameters()]
    host = synth.floorplans_batch(graphs=args.graphs, seed=1 if world == 1 else 1000 + rank).pin_memory()
    resident = host.to(dev)
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    flat = dp.FlatGradients(params) if world > 1 else None

    def eager_step(batch):
        for p in params:
            p.grad = None
        out = model(batch, None)
        loss = crit(out, batch)['loss']
        loss.backward()
        if flat is not None:
            flat.all_reduce_mean()
        return loss

    # The step is ~150 short kernels: eager launching is host-bound, so the product path replays it as one CUDA
    # graph (graphed.py); the gradient all-reduce of N > 1 follows the replay on the same stream.
    graphed = GraphedStep(model, crit)

    def graph_step(batch):
        loss = graphed(batch)
        if flat is not None:
            flat.all_reduce_mean()
        return loss

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

    # ---- end to end through the public API with host buffers: e2e -----------------------------------
    e2e = None
    if not args.no_e2e:
        if step is graph_step:
            # the public API a training loop uses: a collate that writes ONE pinned buffer (batch.PackedBatch), one
            # host->device copy per step, the next step's copy staged on a copy stream while this step computes
            # (GraphedStep.prefetch), the loss read back every step (train.py:286)
            from yolat_vectorgraphicsrecognition_b200.batch import PackedBatch
            packed = PackedBatch.from_batch(synth.floorplans_batch(graphs=args.graphs, seed=1 if world == 1 else 1000 + rank))

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

    # ---- roofline of the scatter path: the K-EDGE kernels of one block-layer GraphConv forward ----------
    roof = None
    if rank == 0:
        hbm, which = peaks()
        sg = args.scatter_graphs if args.scatter_graphs > 0 else args.graphs
        big = synth.floorplans_batch(graphs=sg, seed=7).to(dev)
        Nn, Ee = big.x.shape[0], big.edge.shape[0]
        conv = model.cls_net.backbone[0].body
        xin = torch.randn(Nn, 64, device=dev)
        xnode = torch.randn(Nn, 64, device=dev)
        graph = CSRGraph(big.edge.T, Nn)
        import ctypes as C
        with torch.no_grad():
            for _ in range(3):
                conv(xin, graph, None, big.e_attr, x_node=xnode)
            torch.cuda.synchronize()
            lib.yolat_prof_enable(1)     # CUDA events around the K-EDGE launches, on the launching stream
            reps, tot = 10, 0.0
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                conv(xin, graph, None, big.e_attr, x_node=xnode)
                b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            lib.yolat_prof_enable(0)
        per_kernel = {}
        for name, kid in (('k_edge_stats1', 0), ('k_edge_fused<F_STATS>', 1), ('k_edge_fused<F_AGG>', 2)):
            n, t = C.c_int64(0), C.c_double(0.0)
            lib.yolat_prof_read(kid, C.byref(n), C.byref(t))
            per_kernel[name] = {'launches': int(n.value), 'ms': (t.value / n.value) if n.value else None}
        op_ms = tot / reps
        k_ms = per_kernel['k_edge_fused<F_AGG>']['ms']
        nbytes = edge_path_bytes(Nn, Ee, 64, 64)
        ach = nbytes / (k_ms * 1e-3) / 1e9
        op_bytes = edge_bytes(Nn, Ee, 64, 64, 64)
        roof = {'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm,
                'traffic': ncu_traffic(),
                'kernel': 'ef::k_edge_fused<F_AGG> (K-EDGE pass C: gather + Lin1 + BN1/ReLU + tcgen05 Lin2 + BN2/ReLU + '
                          'segmented mean-scatter) of one block-layer GraphConv(64->64) forward, training mode',
                'algorithmic_bytes': nbytes, 'ms': k_ms, 'peak_source': which,
                'shape': {'N': Nn, 'E': Ee, 'graphs': sg}, 'l2': 'flushed before every call',
                'timing': 'CUDA events recorded around the launch on its own stream (yolat_prof_*), avg of %d' % reps,
                'passes': per_kernel,
                'whole_op': {'what': 'yolat_gp2_fwd: lin_r + P/Q GEMMs, passes A/B/C, BN finalizes, node branch',
                             'ms': op_ms, 'algorithmic_bytes': op_bytes,
                             'achieved': op_bytes / (op_ms * 1e-3) / 1e9, 'frac': op_bytes / (op_ms * 1e-3) / 1e9 / hbm}}
        del big, xin, xnode, graph

    # ---- CPU baseline (rank 0, N = 1) -------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference_run(args.graphs, args.cpu_steps, 2)

    if rank == 0:
        line = {
            'metric': 'graphs_per_sec_fwd_bwd', 'value': args.graphs * world / (ms * 1e-3), 'unit': 'graphs/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'step': 'fwd+loss+bwd' + ('+nccl grad all-reduce' if world > 1 else ''),
                       'graphs_per_step_per_gpu': args.graphs, 'in_channels': 5, 'n_blocks': 2, 'n_filters': 64,
                       'n_classes': 17, 'l2': 'flushed between timed steps (256 MiB memset, untimed)',
                       'launch': 'one CUDA-graph replay per step' if args.mode == 'graph' else 'eager',
                       'parallelism': 'dp%d' % world},
            'clocks': clk, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roof, 'cpu_baseline': cpu,
            'launch_mode': args.mode,
            'other_mode': {'mode': 'eager' if args.mode == 'graph' else 'graph', 'ms_per_step': ms_other},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
