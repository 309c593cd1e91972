"""ctypes wrapper of the CPU oracle (oracle/oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_lib = None
_fast_lib = None
_use_fast = False


def build():
    subprocess.run(["make", "-C", _here, "-s"], check=True)


def cpu_tag() -> str:
    """Identifies the host CPU's instruction set: -march=native builds (the fast arm) are per host, they must not travel
    from the build container to the GPU box."""
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith(("flags", "Features")):
                    flags = " ".join(sorted(line.split(":", 1)[1].split()))
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


def fast_lib():
    """bench.py's `cpu_fast` baseline arm: the same source at -O3 -march=native with the static-pivot sparse LU
    (oracle.cpp, "fast arm").  Built on first use on the host it runs on.  Never used as the checker."""
    global _fast_lib
    if _fast_lib is None:
        name = f"_build/liboracle_fast_{cpu_tag()}.so"
        path = os.path.join(_here, name)
        # always through make: a library left over from an older oracle.cpp / header (same CPU tag) must not be loaded
        subprocess.run(["make", "-C", _here, "-s", "fast", f"FAST_OUT={name}"], check=True)
        _fast_lib = C.CDLL(path)
        _fast_lib.orc_wave_value.restype = C.c_double
    return _fast_lib


class fast_arm:
    """with orc.fast_arm(): ... -- route orc.dc / orc.tran through the fast library (timing baseline only)."""

    def __enter__(self):
        global _use_fast
        fast_lib()
        self.prev, _use_fast = _use_fast, True
        return self

    def __exit__(self, *exc):
        global _use_fast
        _use_fast = self.prev


def lib():
    global _lib
    if _use_fast:
        return fast_lib()
    if _lib is None:
        path = os.path.join(_here, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.orc_wave_value.restype = C.c_double
    return _lib


def set_sparse(on: bool):
    """Static-pivot sparse LU instead of dense partial pivoting in orc.dc / orc.tran (what the fast arm defaults to);
    tests use it to check the sparse path against the dense checker inside the checker's own build."""
    lib().orc_set_sparse(C.c_int(1 if on else 0))


def _flat():
    import cedarsim.jl_b200.flat as flat
    return flat


def default_options(**kw):
    flat = _flat()
    o = flat.cb_options()
    lib().orc_options_default(C.byref(o))
    for k, v in kw.items():
        if k in ("temp", "gmin"):
            setattr(o, k, flat._pref(v))
        else:
            if not hasattr(o, k):
                raise KeyError(k)
            setattr(o, k, v)
    return o


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


_x0_keep = None


def set_x0(x0=None):
    """Initial guess of the DC Newton: None, x0[N] (shared) or x0[N, B] (per point)."""
    global _x0_keep
    if x0 is None:
        _x0_keep = None
        lib().orc_set_x0(None, C.c_int64(0))
        return
    _x0_keep = np.ascontiguousarray(x0, dtype=np.float64)
    stride = 0 if _x0_keep.ndim == 1 else _x0_keep.shape[1]
    lib().orc_set_x0(_dp(_x0_keep), C.c_int64(stride))


def dc(fc, params=None, B=1, opts=None, nthreads=1):
    """fc: FlatCircuit. Returns (x_out [O,B], x_full [N,B], status [B], stats dict)."""
    flat = _flat()
    pk = fc.pack()
    params = np.zeros((max(1, len(fc.param_names)), B)) if params is None else np.ascontiguousarray(params, dtype=np.float64)
    if params.size:
        B = params.shape[1]
    opts = opts or default_options()
    O, N = len(fc.outputs), fc.n_unknowns
    x_out = np.zeros((O, B)); x_full = np.zeros((N, B)); status = np.zeros(B, dtype=np.int32)
    st = flat.cb_stats()
    rc = lib().orc_dc(pk.ref(), _dp(params), C.c_int64(B), C.byref(opts), _dp(x_out), _dp(x_full),
                      status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(st), C.c_int(nthreads))
    assert rc == 0
    return x_out, x_full, status, st.as_dict()


def tran(fc, t0, t1, saveat, params=None, B=1, opts=None, nthreads=1):
    """Returns (y [O,S,B], status [B], stats dict)."""
    flat = _flat()
    pk = fc.pack()
    params = np.zeros((max(1, len(fc.param_names)), B)) if params is None else np.ascontiguousarray(params, dtype=np.float64)
    if params.size:
        B = params.shape[1]
    opts = opts or default_options()
    saveat = np.ascontiguousarray(saveat, dtype=np.float64)
    O, S = len(fc.outputs), len(saveat)
    y = np.zeros((O, S, B)); status = np.zeros(B, dtype=np.int32)
    st = flat.cb_stats()
    rc = lib().orc_tran(pk.ref(), _dp(params), C.c_int64(B), C.c_double(t0), C.c_double(t1), _dp(saveat),
                        C.c_int64(S), C.byref(opts), _dp(y), status.ctypes.data_as(C.POINTER(C.c_int32)),
                        C.byref(st), C.c_int(nthreads))
    assert rc == 0
    return y, status, st.as_dict()


def ac(fc, freqs, params=None, B=1, opts=None, nthreads=1):
    """AC response of the outputs about each point's DC operating point: (y complex [O,F,B], status [B])."""
    pk = fc.pack()
    params = np.zeros((max(1, len(fc.param_names)), B)) if params is None else np.ascontiguousarray(params, dtype=np.float64)
    if params.size:
        B = params.shape[1]
    opts = opts or default_options()
    freqs = np.ascontiguousarray(freqs, dtype=np.float64)
    O, F = len(fc.outputs), len(freqs)
    y = np.zeros((O, F, B, 2)); status = np.zeros(B, dtype=np.int32)
    rc = lib().orc_ac(pk.ref(), _dp(params), C.c_int64(B), _dp(freqs), C.c_int64(F), C.byref(opts), _dp(y),
                      status.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(nthreads))
    assert rc == 0
    return y[..., 0] + 1j * y[..., 1], status


def noise(fc, freqs, params=None, B=1, opts=None, nthreads=1):
    """Output noise power spectral density about each point's DC operating point: (psd [O,F,B], status [B])."""
    pk = fc.pack()
    params = np.zeros((max(1, len(fc.param_names)), B)) if params is None else np.ascontiguousarray(params, dtype=np.float64)
    if params.size:
        B = params.shape[1]
    opts = opts or default_options()
    freqs = np.ascontiguousarray(freqs, dtype=np.float64)
    O, F = len(fc.outputs), len(freqs)
    psd = np.zeros((O, F, B)); status = np.zeros(B, dtype=np.int32)
    rc = lib().orc_noise(pk.ref(), _dp(params), C.c_int64(B), _dp(freqs), C.c_int64(F), C.byref(opts), _dp(psd),
                         status.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(nthreads))
    assert rc == 0
    return psd, status


def eval_system(fc, x, t=0.0, dcop=False, params=None, b=0, opts=None):
    """One assembled evaluation: returns f[N], q[N], G[N,N], C[N,N]."""
    pk = fc.pack()
    params = np.zeros((max(1, len(fc.param_names)), 1)) if params is None else np.ascontiguousarray(params, dtype=np.float64)
    B = params.shape[1]
    opts = opts or default_options()
    N = fc.n_unknowns
    x = np.ascontiguousarray(x, dtype=np.float64)
    f = np.zeros(N); q = np.zeros(N); G = np.zeros((N, N)); Cm = np.zeros((N, N))
    lib().orc_eval(pk.ref(), _dp(params), C.c_int64(B), C.c_int64(b), C.byref(opts), _dp(x), C.c_double(t),
                   C.c_int(1 if dcop else 0), _dp(f), _dp(q), _dp(G), _dp(Cm))
    return f, q, G, Cm


def wave_value(fc, wave, t, dcop=False, params=None, b=0):
    pk = fc.pack()
    params = np.zeros((max(1, len(fc.param_names)), 1)) if params is None else np.ascontiguousarray(params, dtype=np.float64)
    return lib().orc_wave_value(pk.ref(), C.c_int(wave), _dp(params), C.c_int64(params.shape[1]), C.c_int64(b),
                                C.c_double(t), C.c_int(1 if dcop else 0))
