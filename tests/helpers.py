"""Shared helpers of the parity tests: run the same flat circuit through the CUDA engine (via the
C ABI) and through the CPU oracle."""
import numpy as np

from cedarsim.jl_b200 import engine
from oracle import orc

OPT_KEYS = ("reltol", "vabstol", "iabstol", "nr_reltol", "nr_vabstol", "nr_iabstol", "dc_abstol", "dv_max",
            "max_newton_dc", "max_newton_tran", "method", "fixed_step", "dt", "dt_min", "dt_max", "gmin_steps", "skip_dc",
            "temp", "gmin")


def both_options(**kw):
    return engine.default_options(**kw), orc.default_options(**kw)


def x0_from(fc, nodeset):
    from cedarsim.jl_b200.flat import nodeset_vector
    return nodeset_vector(fc, nodeset)


def run_dc_both(fc, models, P, x0=None, **kw):
    eo, oo = both_options(**kw)
    B = P.shape[1]
    c = engine.Circuit(fc, models)
    p = c.plan(B)
    p.set_params(P)
    p.set_x0(x0)
    xg, xfg, sg, stg = p.dc(eo)
    p.close()
    orc.set_x0(x0)
    xo, xfo, so, sto = orc.dc(fc, P, opts=oo)
    orc.set_x0(None)
    return (xg, xfg, sg, stg), (xo, xfo, so, sto)


def run_tran_both(fc, models, P, t0, t1, saveat, x0=None, B=None, engine_only=None, **kw):
    """engine_only: options given to the CUDA engine but not to the oracle (throughput options whose results
    are checked against the oracle's plain Newton)."""
    eo, oo = both_options(**kw)
    for k, v in (engine_only or {}).items():
        setattr(eo, k, v)
    B = P.shape[1] if P is not None and P.size else (B or 1)
    c = engine.Circuit(fc, models)
    p = c.plan(B)
    p.set_params(P)
    p.set_x0(x0)
    yg, sg, stg = p.tran(t0, t1, saveat, eo)
    p.close()
    orc.set_x0(x0)
    yo, so, sto = orc.tran(fc, t0, t1, saveat, params=P if P is not None and P.size else None, B=B, opts=oo)
    orc.set_x0(None)
    return (yg, sg, stg), (yo, so, sto)
