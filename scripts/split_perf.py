"""Experiment: the same 16 384-point DFF workload as ONE plan vs K concurrent plans (one host thread + stream each)
on one GPU, to see whether the tail waves / latency-bound kernels of one half overlap the other half's work."""
import sys, time, os, threading; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import bench
from cedarsim.jl_b200 import circuits, engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
span = float(sys.argv[2]) if len(sys.argv) > 2 else 6e-7
splits = [int(x) for x in (sys.argv[3].split(',') if len(sys.argv) > 3 else ['1', '2', '4'])]
fc, ms = circuits.dff()
P = circuits.dff_mc_params(fc, B)
x0 = bench.nodeset(fc)
c = engine.Circuit(fc, ms)
ts = np.linspace(0, span, int(round(span / 6e-7 * 1800)) + 1)
o = dict(bench.OPTS); o.update(bench.ENGINE_OPTS)
for K in splits:
    per = B // K
    plans = []
    for k in range(K):
        p = c.plan(per); p.set_params(np.ascontiguousarray(P[:, k * per:(k + 1) * per])); p.set_x0(x0); plans.append(p)
    for rep in range(2):
        t = time.time()
        th = [threading.Thread(target=lambda p=p: p.tran_device(0.0, span, ts, engine.default_options(**o))) for p in plans]
        [x.start() for x in th]; [x.join() for x in th]
        el = time.time() - t
        print(f'K={K} rep{rep} wall {el:.3f}s points/s {B / el:.1f}', flush=True)
    for p in plans: p.close()
