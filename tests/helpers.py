"""Shared helpers of the parity tests: run the same flat circuit through the CUDA engine (via the
C ABI) and through the CPU oracle."""
import numpy as np

from cedarsim.jl_b200 import engine
from oracle import orc

OPT_KEYS = ("reltol", "vabstol", "iabstol", "nr_reltol", "nr_vabstol", "nr_iabstol", "dc_abstol", "dv_max",
            "max_newton_dc", "max_newton_tran", "method", "fixed_step", "dt", "dt_min", "dt_max", "gmin_steps", "skip_dc",
            "temp", "gmin", "nr_rate_test", "value_rounds", "mixed_rounds", "source_steps")


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


def _oracle_tran_cached(fc, models, P, t0, t1, saveat, x0, B, oo, nthreads):
    """The oracle's answer for one (circuit, inputs, options) set, kept under tests/_cache/ (git-ignored, travels to the
    GPU box like built .so files) keyed by a hash of the oracle source, the generated model sources and every input:
    the long config-3 runs take minutes of CPU, which would otherwise be spent while a GPU box is held."""
    import ctypes as C
    import hashlib
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha1()
    import tempfile
    from cedarsim.jl_b200.flat import save_flatckt
    for path in ("oracle/oracle.cpp", "oracle/Makefile", "cedarsim.jl_b200/csrc/symbolic.hpp"):
        h.update(open(os.path.join(root, path), "rb").read())
    with tempfile.NamedTemporaryFile(suffix=".flatckt") as tf:      # the whole flat circuit + generated model sources
        save_flatckt(fc, models, tf.name)
        h.update(open(tf.name, "rb").read())
    for arr in (P, np.asarray(saveat, dtype=np.float64), np.asarray([t0, t1], dtype=np.float64),
                np.zeros(0) if x0 is None else np.asarray(x0, dtype=np.float64)):
        h.update(np.ascontiguousarray(arr).tobytes())
    h.update(C.string_at(C.addressof(oo), C.sizeof(oo)))
    h.update(repr(B).encode())
    path = os.path.join(root, "tests", "_cache", f"orc_tran_{h.hexdigest()[:20]}.npz")
    if os.path.exists(path):
        z = np.load(path, allow_pickle=True)
        return z["y"], z["st"], z["stats"].item()
    orc.set_x0(x0)
    yo, so, sto = orc.tran(fc, t0, t1, saveat, params=P if P is not None and P.size else None, B=B, opts=oo, nthreads=nthreads)
    orc.set_x0(None)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez(path + f".{os.getpid()}.tmp.npz", y=yo, st=so, stats=np.array(sto, dtype=object))
    os.replace(path + f".{os.getpid()}.tmp.npz", path)
    return yo, so, sto


def run_tran_both(fc, models, P, t0, t1, saveat, x0=None, B=None, engine_only=None, nthreads=1, cache_oracle=False, oracle_only=False, **kw):
    """engine_only: options given to the CUDA engine but not to the oracle (throughput options whose results
    are checked against the oracle's plain Newton).  nthreads: OpenMP threads of the oracle (points are independent).
    cache_oracle: keep / re-use the oracle's answer under tests/_cache/ (see _oracle_tran_cached)."""
    eo, oo = both_options(**kw)
    for k, v in (engine_only or {}).items():
        setattr(eo, k, v)
    B = P.shape[1] if P is not None and P.size else (B or 1)
    if oracle_only:      # scripts/fill_oracle_cache.py: compute the cached oracle answers where no GPU is needed
        return None, _oracle_tran_cached(fc, models, P, t0, t1, saveat, x0, B, oo, nthreads)
    c = engine.Circuit(fc, models)
    p = c.plan(B)
    p.set_params(P)
    p.set_x0(x0)
    yg, sg, stg = p.tran(t0, t1, saveat, eo)
    p.close()
    if cache_oracle:
        return (yg, sg, stg), _oracle_tran_cached(fc, models, P, t0, t1, saveat, x0, B, oo, nthreads)
    orc.set_x0(x0)
    yo, so, sto = orc.tran(fc, t0, t1, saveat, params=P if P is not None and P.size else None, B=B, opts=oo, nthreads=nthreads)
    orc.set_x0(None)
    return (yg, sg, stg), (yo, so, sto)
