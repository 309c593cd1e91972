import sys, time, os; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from cedarsim.jl_b200 import circuits, engine, models
from helpers import x0_from
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mode = sys.argv[2] if len(sys.argv) > 2 else 'adaptive'
span = float(sys.argv[3]) if len(sys.argv) > 3 else 6e-7
fc, ms = circuits.dff()
P = circuits.dff_mc_params(fc, B)
x0 = x0_from(fc, dict(q=0.0, q_neg=0.7, net0=0.0, net7=0.0, vdd=0.7, clkn=0.7, ncki=0.0, cki=0.7))
c = engine.Circuit(fc, ms)
print('lu', c.lu_info(), 'N', fc.n_unknowns)
p = c.plan(B); p.set_params(P); p.set_x0(x0)
ts = np.linspace(0, span, int(round(span / 6e-7 * 1800)) + 1)
kw = dict(reltol=float(os.environ.get('RELTOL', '1e-3'))) if mode == 'adaptive' else dict(fixed_step=1, dt=25e-12)
for k in ('nr_reltol', 'nr_vabstol', 'nr_iabstol', 'vabstol', 'iabstol'):
    if k.upper() in os.environ: kw[k] = float(os.environ[k.upper()])
for k in ('nr_rate_test', 'value_rounds'):
    if k.upper() in os.environ: kw[k] = int(os.environ[k.upper()])
for rep in range(2):
    t = time.time()
    dy, ds, st = p.tran_device(0.0, span, ts, engine.default_options(**kw))
    el = time.time() - t
    print(f'rep{rep} B={B} {mode} wall {el:.3f}s dev {st["solve_seconds"]:.3f}s rounds {st["rounds"]} newton {st["newton_iters"]} acc {st["steps_accepted"]} rej {st["steps_rejected"]} '
          f'full_it {st["full_iters"]} vrounds {st["value_rounds"]} eval {st["eval_seconds"]:.3f} newton_k {st["newton_seconds"]:.3f} evalv {st["evalv_seconds"]:.3f} newtonv_k {st["newtonv_seconds"]:.3f}  points/s {B/el:.1f} it/s {st["newton_iters"]/el:.3e}')
