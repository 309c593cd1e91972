import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from cedarsim.jl_b200.flat import *
from helpers import run_tran_both
fc = FlatCircuit()
fc.vsource("V", "in", "0", Wave(W_PULSE, v=[0, 1, 1e-8, 1e-9, 1e-9, 2e-6, 4e-6]))
fc.resistor("R", "in", "out", fc.param("r"))
fc.capacitor("C", "out", "0", 1e-9)
fc.set_outputs(["out", "v.i"])
for B, S in ((37, 101), (5, 101), (37, 501)):
    P = params_matrix([np.linspace(500.0, 2000.0, B)])
    ts = np.linspace(0, 5e-6, S)
    for method in (0, 1, 2):
        (yg, sg, stg), (yo, so, sto) = run_tran_both(fc, [], P, 0.0, 5e-6, ts, fixed_step=1, dt=1e-8, method=method)
        err = np.abs(yg - yo)
        k = np.unravel_index(np.argmax(err), err.shape)
        print(B, S, method, err.max(), k, ts[k[1]], yg[k], yo[k], stg['newton_iters'], sto['newton_iters'])
        bad = np.argwhere(err[0] > 1e-9)
        print('   bad samples', len(bad), bad[:6].tolist())
