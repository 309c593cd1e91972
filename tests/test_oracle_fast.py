"""The fast CPU arm (static-pivot sparse LU from the engine's symbolic analysis, bench.py's `cpu_fast` baseline) against
the checker (dense partial pivoting): same equations and step control, so DC and fixed-step transients agree to rounding
and adaptive runs to the tolerance.  The fast arm is a timing baseline only; parity tests never use it."""
import numpy as np

from cedarsim.jl_b200 import circuits
from oracle import orc


def _both(fn):
    orc.set_sparse(False)
    a = fn()
    orc.set_sparse(True)
    try:
        b = fn()
    finally:
        orc.set_sparse(False)
    return a, b


def test_sparse_lu_matches_dense_on_linear_dc_sweep():
    fc = circuits.two_resistor()
    r = np.arange(100, 2001, 100.0)
    P = np.stack([np.repeat(r, len(r)), np.tile(r, len(r))])
    (xa, _, sa, _), (xb, _, sb, _) = _both(lambda: orc.dc(fc, P))
    assert sa.max() == 0 and sb.max() == 0
    assert np.abs(xa - xb).max() < 1e-15


def test_sparse_lu_matches_dense_on_bsimcmg_dff_transient(host_bsimcmg):
    fc, _ = circuits.dff(host=True, tscale=0.01)
    P = circuits.dff_mc_params(fc, 4)
    from bench import nodeset
    ts = np.linspace(0, 1.5e-9, 31)
    kw = dict(fixed_step=1, dt=2.5e-12, nr_rate_test=1, nr_reltol=1e-5, nr_vabstol=1e-7)

    def run():
        orc.set_x0(nodeset(fc))
        try:
            return orc.tran(fc, 0.0, 1.5e-9, ts, params=P, opts=orc.default_options(**kw))
        finally:
            orc.set_x0(None)
    (ya, sa, sta), (yb, sb, stb) = _both(run)
    assert sa.max() == 0 and sb.max() == 0
    assert np.abs(ya - yb).max() < 1e-9
    assert abs(sta["newton_iters"] - stb["newton_iters"]) <= 0.01 * sta["newton_iters"]
