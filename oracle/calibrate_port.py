"""Calibrate the CPU baseline port (build container only): time one fwd + loss + bwd step of config 2 with
(a) the UNMODIFIED reference model imported through oracle/shims and (b) oracle/restatement.py (the `cpu_baseline.kind =
"port"` that bench.py times on the GPU box, where /root/reference does not exist), same seeds, same threads.

    python oracle/calibrate_port.py [graphs] [steps]      -> one JSON line (recorded in BASELINE.md section 4)"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, restatement as R          # noqa: E402
from yolat_vectorgraphicsrecognition_b200 import synth    # noqa: E402


def main():
    graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    opt = synth.make_opt(n_classes=17)
    batch = synth.floorplans_batch(graphs=graphs, seed=1)
    arch = ref_loader.load()
    torch.manual_seed(0)
    model = arch.SparseCADGCN(opt).train()
    crit = arch.DetectionLoss(opt)
    state = R.clone_state(model.state_dict(), torch.float32)

    def ref_step():
        for p in model.parameters():
            p.grad = None
        loss = crit(model(batch, None), batch)['loss']
        loss.backward()
        return float(loss)

    def port_step():
        return float(R.run_step(state, opt, batch, training=True)['loss'])

    out = {}
    for name, fn in (('reference', ref_step), ('port', port_step)):
        ts = []
        for it in range(steps + 2):
            t0 = time.perf_counter()
            loss = fn()
            if it >= 2:
                ts.append(time.perf_counter() - t0)
        ts.sort()
        out[name] = {'s_per_step': ts[len(ts) // 2], 'graphs_per_s': graphs / ts[len(ts) // 2], 'loss': loss}
    out['port_over_reference_speed'] = out['reference']['s_per_step'] / out['port']['s_per_step']
    out.update(cores=cores, graphs=graphs, torch=torch.__version__,
               note='reference = unmodified /root/reference model through oracle/shims (PyG propagate + torch_scatter stand-ins)')
    print(json.dumps(out))


if __name__ == '__main__':
    main()
