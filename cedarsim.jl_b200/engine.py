"""ctypes binding of the C-ABI engine (include/cedarb200.h, csrc/libcedarb200.so).

This is the thin host layer the north star describes: the host flattens the netlist and
generates CUDA C for the device models; everything numeric happens behind the `cb_*` calls.
There is no CPU fallback: `load()` raises when the CUDA library has not been built and the
library itself fails with CB_ERR_NO_DEVICE when no GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import flat as F

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CB_ENGINE_LIB") or os.path.join(_HERE, "csrc", "libcedarb200.so")   # CB_ENGINE_LIB: experiment builds
CUBIN_CACHE = os.path.join(_HERE, os.environ.get("CB_GEN_DIR", "_gen"), "cubin")
_lib = None

SYMBOLS = [
    "cb_version", "cb_last_error", "cb_options_default", "cb_options_size", "cb_stats_size", "cb_options_init", "cb_circuit_create", "cb_circuit_load", "cb_circuit_set_cuda_source",
    "cb_circuit_compile", "cb_circuit_lu_info", "cb_plan_create", "cb_plan_create_lanes", "cb_plan_create_multi", "cb_plan_devices", "cb_plan_lanes", "cb_plan_set_params", "cb_dc", "cb_tran", "cb_sens_dc",
    "cb_ac", "cb_noise", "cb_plan_device_params", "cb_plan_set_x0", "cb_plan_set_timing", "cb_measure_fp64_peak", "cb_tran_device", "cb_dc_device", "cb_plan_destroy", "cb_circuit_destroy",
]


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cedarb200 error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(-100, f"{LIB_PATH} is missing: build the CUDA engine first "
                              "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        lib.cb_last_error.restype = C.c_char_p
        lib.cb_plan_destroy.restype = None
        lib.cb_circuit_destroy.restype = None
        lib.cb_options_default.restype = None
        lib.cb_options_size.restype = C.c_size_t
        lib.cb_stats_size.restype = C.c_size_t
        # layout guard: this binding's mirrors of the option / statistics structs must be the library's
        if lib.cb_options_size() != C.sizeof(F.cb_options) or lib.cb_stats_size() != C.sizeof(F.cb_stats):
            raise EngineError(-101, f"{LIB_PATH}: struct layout mismatch (cb_options {lib.cb_options_size()} vs "
                              f"{C.sizeof(F.cb_options)} bytes, cb_stats {lib.cb_stats_size()} vs {C.sizeof(F.cb_stats)}); rebuild the library")
        _lib = lib
    return _lib


def _check(rc: int):
    if rc != 0:
        raise EngineError(rc, load().cb_last_error().decode(errors="replace"))


def default_options(**kw) -> F.cb_options:
    o = F.cb_options()
    _check(load().cb_options_init(C.byref(o), C.c_size_t(C.sizeof(o))))
    for k, v in kw.items():
        if k in ("temp", "gmin"):
            setattr(o, k, F._pref(v))
        else:
            if not hasattr(o, k):
                raise KeyError(f"unknown option {k!r}")
            setattr(o, k, v)
    return o


def cuda_source(models: Sequence) -> str:
    """Concatenate generated model sources with their shape macros (see csrc/va_prelude.h)."""
    parts = []
    for cm in models:
        parts.append(f"#undef NT\n#undef NPARAM\n#undef NCACHE\n#undef NOUT\n#undef NJ\n"
                     f"#define NT {len(cm.terminals)}\n#define NPARAM {max(1, len(cm.params))}\n"
                     f"#define NCACHE {max(1, cm.ncache)}\n#define NJ {len(cm.jrow)}\n"
                     f"#define NOUT {2 * len(cm.terminals) + 2 * len(cm.jrow)}\n"
                     "#undef VA_LAYOUT\n#define VA_LAYOUT VA_CACHE_LAYOUT\n")
        parts.append(cm.source)
        if getattr(cm, "source_v", ""):   # value-only variant: its own cache (and cache layout, csrc/va_prelude.h)
            parts.append(f"#undef NCACHE\n#define NCACHE {max(1, cm.ncache_v)}\n#undef VA_LAYOUT\n#define VA_LAYOUT VA_CACHE_LAYOUT_V\n")
            parts.append(cm.source_v)
            parts.append("#undef VA_LAYOUT\n#define VA_LAYOUT VA_CACHE_LAYOUT\n")
        if getattr(cm, "source_n", ""):   # noise variant: source powers at the operating point (cb_noise)
            K = len(cm.noise_sources)
            parts.append(f"#undef NCACHE\n#undef NOUT\n#undef NNOISE\n#define NCACHE {max(1, cm.ncache_n)}\n"
                         f"#define NNOISE {K}\n#define NOUT {2 * K}\n")
            parts.append(cm.source_n)
    return "\n".join(parts)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def measure_fp64_peak(device: int = 0) -> float:
    """FP64 FMA TFLOP/s of `device` (DFMA microbenchmark in the engine library)."""
    tf = C.c_double(0.0)
    _check(load().cb_measure_fp64_peak(C.c_int(device), C.byref(tf)))
    return tf.value


class Circuit:
    """A compiled circuit: flat description + generated device code + symbolic LU."""

    def __init__(self, fc: F.FlatCircuit, models: Sequence = (), cache_dir: Optional[str] = CUBIN_CACHE, _handle=None):
        self.lib = load()
        self.fc = fc
        self.handle = C.c_void_p()
        if _handle is not None:        # created by the library itself (NativeNetlist: cb_netlist_circuit)
            self.packed = None
            self.handle = _handle
        else:
            self.packed = fc.pack()
            _check(self.lib.cb_circuit_create(self.packed.ref(), C.byref(self.handle)))
        if fc.va_models:
            by_name = {cm.name: cm for cm in models}
            src = cuda_source([by_name[m.name] for m in fc.va_models]).encode()
            _check(self.lib.cb_circuit_set_cuda_source(self.handle, src, C.c_size_t(len(src))))
        secs = C.c_double(0.0)
        if cache_dir:
            os.makedirs(cache_dir, exist_ok=True)
        _check(self.lib.cb_circuit_compile(self.handle, cache_dir.encode() if cache_dir else None, C.byref(secs)))
        self.compile_seconds = secs.value

    def lu_info(self):
        a, l, f = C.c_int32(), C.c_int32(), C.c_int64()
        _check(self.lib.cb_circuit_lu_info(self.handle, C.byref(a), C.byref(l), C.byref(f)))
        return {"nnz_a": a.value, "nnz_lu": l.value, "lu_flops": f.value}

    def plan(self, n_inst: int, device: int = 0, lanes: int = 0, devices: Optional[Sequence[int]] = None) -> "Plan":
        """devices: several GPUs of this process (cb_plan_create_multi: contiguous block of points per GPU, `lanes`
        lanes on each); else one GPU."""
        return Plan(self, n_inst, device, lanes, devices)

    def __del__(self):
        try:
            if self.handle:
                self.lib.cb_circuit_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Plan:
    def __init__(self, circuit: Circuit, n_inst: int, device: int = 0, lanes: int = 0, devices: Optional[Sequence[int]] = None):
        """lanes: concurrent sub-plans on the device (0 = automatic, see cb_plan_create_lanes)."""
        self.circuit = circuit
        self.lib = circuit.lib
        self.B = int(n_inst)
        self.device = device if not devices else int(devices[0])
        self.handle = C.c_void_p()
        if devices is not None and len(devices) > 1:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            _check(self.lib.cb_plan_create_multi(circuit.handle, C.c_int64(self.B), ids, C.c_int(len(devices)), C.c_int(lanes),
                                                 C.byref(self.handle)))
        else:
            _check(self.lib.cb_plan_create_lanes(circuit.handle, C.c_int64(self.B), C.c_int(self.device), C.c_int(lanes),
                                                 C.byref(self.handle)))
        self.lanes = int(self.lib.cb_plan_lanes(self.handle))
        self.last_status_ptr = None
        self.n_devices = int(self.lib.cb_plan_devices(self.handle))

    def set_params(self, params: Optional[np.ndarray]):
        P = len(self.circuit.fc.param_names)
        if P == 0:
            _check(self.lib.cb_plan_set_params(self.handle, None))
            return
        params = np.ascontiguousarray(params, dtype=np.float64)
        if params.shape != (P, self.B):
            raise ValueError(f"params must have shape ({P}, {self.B}), got {params.shape}")
        self._params = params
        _check(self.lib.cb_plan_set_params(self.handle, _dp(params)))

    def set_timing(self, enable: bool):
        _check(self.lib.cb_plan_set_timing(self.handle, C.c_int(1 if enable else 0)))

    def set_x0(self, x0: Optional[np.ndarray]):
        """DC initial guess: None, x0[N] shared by all points, or x0[N, B]."""
        if x0 is None:
            _check(self.lib.cb_plan_set_x0(self.handle, None, C.c_int(0)))
            return
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        N = self.circuit.fc.n_unknowns
        if x0.shape not in ((N,), (N, self.B)):
            raise ValueError(f"x0 must have shape ({N},) or ({N}, {self.B})")
        _check(self.lib.cb_plan_set_x0(self.handle, _dp(x0), C.c_int(1 if x0.ndim == 2 else 0)))

    def dc(self, opts: Optional[F.cb_options] = None, want_full: bool = True):
        opts = opts or default_options()
        fc = self.circuit.fc
        O, N = len(fc.outputs), fc.n_unknowns
        x_out = np.zeros((O, self.B))
        x_full = np.zeros((N, self.B)) if want_full else None
        status = np.zeros(self.B, dtype=np.int32)
        st = F.cb_stats()
        _check(self.lib.cb_dc(self.handle, C.byref(opts), _dp(x_out), _dp(x_full) if want_full else None,
                              status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(st)))
        return x_out, x_full, status, st.as_dict()

    def sens_dc(self, params_plus: np.ndarray, params_minus: np.ndarray, step: np.ndarray, opts: Optional[F.cb_options] = None):
        """Operating point and its forward sensitivities by the direct method (cb_sens_dc): direction d moves the
        parameters to params_plus[d] / params_minus[d] ([n_dir, P, B]), step[d] ([n_dir, B]) is the distance moved.
        Returns (x_out [O, B], sens [n_dir, O, B], status, stats)."""
        opts = opts or default_options()
        fc = self.circuit.fc
        O, P = len(fc.outputs), len(fc.param_names)
        step = np.ascontiguousarray(step, dtype=np.float64)
        n_dir = step.shape[0]
        pp = np.ascontiguousarray(params_plus, dtype=np.float64).reshape(n_dir, max(1, P), self.B) if P else np.zeros((n_dir, 1, self.B))
        pm = np.ascontiguousarray(params_minus, dtype=np.float64).reshape(n_dir, max(1, P), self.B) if P else np.zeros((n_dir, 1, self.B))
        x_out = np.zeros((O, self.B))
        sens = np.zeros((n_dir, O, self.B))
        status = np.zeros(self.B, dtype=np.int32)
        st = F.cb_stats()
        _check(self.lib.cb_sens_dc(self.handle, C.byref(opts), C.c_int64(n_dir), _dp(pp), _dp(pm), _dp(step), _dp(x_out), _dp(sens),
                                   status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(st)))
        return x_out, sens, status, st.as_dict()

    def tran(self, t0: float, t1: float, saveat, opts: Optional[F.cb_options] = None, out: Optional[np.ndarray] = None):
        opts = opts or default_options()
        saveat = np.ascontiguousarray(saveat, dtype=np.float64)
        O, S = len(self.circuit.fc.outputs), len(saveat)
        y = out if out is not None else np.zeros((O, S, self.B))
        status = np.zeros(self.B, dtype=np.int32)
        st = F.cb_stats()
        _check(self.lib.cb_tran(self.handle, C.c_double(t0), C.c_double(t1), _dp(saveat), C.c_int64(S), C.byref(opts),
                                _dp(y), status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(st)))
        return y, status, st.as_dict()

    def _small_signal(self, fn, freqs, opts, width):
        opts = opts or default_options()
        freqs = np.ascontiguousarray(freqs, dtype=np.float64)
        O, Fq = len(self.circuit.fc.outputs), len(freqs)
        out = np.zeros((O, Fq, self.B) + ((2,) if width == 2 else ()))
        status = np.zeros(self.B, dtype=np.int32)
        st = F.cb_stats()
        _check(fn(self.handle, _dp(freqs), C.c_int64(Fq), C.byref(opts), _dp(out),
                  status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(st)))
        return out, status, st.as_dict()

    def ac(self, freqs, opts: Optional[F.cb_options] = None):
        """AC response of the outputs about every point's operating point: (y complex [O, F, B], status, stats)."""
        out, status, st = self._small_signal(self.lib.cb_ac, freqs, opts, 2)
        return out[..., 0] + 1j * out[..., 1], status, st

    def noise(self, freqs, opts: Optional[F.cb_options] = None):
        """Output noise PSD about every point's operating point: (psd [O, F, B], status, stats)."""
        return self._small_signal(self.lib.cb_noise, freqs, opts, 1)

    def tran_device(self, t0: float, t1: float, saveat, opts: Optional[F.cb_options] = None):
        """Results stay in HBM; returns (device pointer of y, device pointer of status, stats)."""
        opts = opts or default_options()
        saveat = np.ascontiguousarray(saveat, dtype=np.float64)
        dy, ds = C.c_void_p(), C.c_void_p()
        st = F.cb_stats()
        _check(self.lib.cb_tran_device(self.handle, C.c_double(t0), C.c_double(t1), _dp(saveat), C.c_int64(len(saveat)),
                                       C.byref(opts), C.byref(dy), C.byref(ds), C.byref(st)))
        self.last_status_ptr = ds.value
        return dy.value, ds.value, st.as_dict()

    def dc_device(self, opts: Optional[F.cb_options] = None):
        opts = opts or default_options()
        dx, ds = C.c_void_p(), C.c_void_p()
        st = F.cb_stats()
        _check(self.lib.cb_dc_device(self.handle, C.byref(opts), C.byref(dx), C.byref(ds), C.byref(st)))
        return dx.value, ds.value, st.as_dict()

    def device_params_ptr(self) -> int:
        dp = C.c_void_p()
        _check(self.lib.cb_plan_device_params(self.handle, C.byref(dp)))
        return dp.value

    def close(self):
        if self.handle:
            self.lib.cb_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NativeNetlist:
    """The library's own netlist front end (include/cedarb200.h cb_netlist_*, csrc/spice_front.hpp): a SPICE-subset deck and
    the sweep's values go in as text and arrays, the flat circuit and the parameter matrix come back -- no Python parser,
    flattener or struct packing on the way.  `fc` is a read-only FlatCircuit VIEW of what the library built (names,
    devices, waves, outputs), so results can be named and the CPU oracle can be handed the same circuit."""

    def __init__(self, text: str, sweep: Optional[dict] = None, outputs: Optional[Sequence[str]] = None, base_dir: Optional[str] = None,
                 lang: str = "spice"):
        lib = load()
        lib.cb_netlist_flat.restype = C.POINTER(F.cb_flat_circuit)
        lib.cb_netlist_params.restype = C.POINTER(C.c_double)
        lib.cb_netlist_unknown_name.restype = C.c_char_p
        lib.cb_netlist_param_name.restype = C.c_char_p
        lib.cb_netlist_option.restype = C.c_double
        self.lib = lib
        sweep = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in (sweep or {}).items()}
        B = len(next(iter(sweep.values()))) if sweep else 1
        names = (C.c_char_p * max(1, len(sweep)))(*[k.encode() for k in sweep])
        vals = np.ascontiguousarray(np.stack(list(sweep.values())) if sweep else np.zeros((0, B)))
        outs = list(outputs or [])
        onames = (C.c_char_p * max(1, len(outs)))(*[o.encode() for o in outs])
        self.handle = C.c_void_p()
        flatten = lib.cb_netlist_flatten_spectre if lang == "spectre" else lib.cb_netlist_flatten
        _check(flatten(text.encode(), base_dir.encode() if base_dir else None, names, C.c_int(len(sweep)),
                                      _dp(vals) if sweep else None, C.c_int64(B), onames, C.c_int(len(outs)), C.byref(self.handle)))
        self.B = B
        P = int(lib.cb_netlist_n_params(self.handle))
        self.params = np.ctypeslib.as_array(lib.cb_netlist_params(self.handle), shape=(P, B)).copy() if P else np.zeros((0, B))
        self.fc = self._view()

    def _view(self) -> F.FlatCircuit:
        lib, h = self.lib, self.handle
        f = lib.cb_netlist_flat(h).contents
        fc = F.FlatCircuit()
        names = [lib.cb_netlist_unknown_name(h, i).decode() for i in range(f.n_unknowns)]
        fc.node_names = names[:f.n_nodes]
        fc.branch_names = names[f.n_nodes:]
        fc._node_index = {n: i for i, n in enumerate(fc.node_names)}
        fc.param_names = [lib.cb_netlist_param_name(h, i).decode() for i in range(f.n_params)]

        def val(p):
            return F.Col(p.col) if p.col >= 0 else float(p.value)
        for i in range(f.n_waves):
            w = f.waves[i]
            wv = F.Wave(int(w.kind), dc=(val(w.dc) if w.has_dc else None))
            if w.kind == F.W_PWL:
                wv.t = [float(w.t[k]) for k in range(w.npts)]
                wv.y = [val(w.y[k]) for k in range(w.npts)]
            elif w.kind in (F.W_PULSE, F.W_SIN):
                wv.v = [val(w.v[k]) for k in range(7)]
            wv.ac = float(w.ac_mag)
            fc.waves.append(wv)
        for i in range(f.n_devices):
            d = f.devices[i]
            nn = {F.DEV_VCVS: 4, F.DEV_VCCS: 4}.get(int(d.kind), 2)
            fc.devices.append(F.Device(int(d.kind), f"d{i}", [int(d.n[k]) for k in range(nn)], val(d.value), int(d.branch), int(d.wave),
                                       float(d.mult)))
        fc.outputs = [int(f.outputs[k]) for k in range(f.n_outputs)]
        fc._finalized = True
        return fc

    def unknown(self, name: str) -> int:
        u = int(self.lib.cb_netlist_unknown(self.handle, name.encode()))
        if u < 0:
            raise KeyError(f"no unknown named {name!r}")
        return u

    def option(self, name: str, default: float = float("nan")) -> float:
        return float(self.lib.cb_netlist_option(self.handle, name.encode(), C.c_double(default)))

    def circuit(self, cache_dir: Optional[str] = CUBIN_CACHE) -> Circuit:
        """cb_netlist_circuit + cb_circuit_compile: the engine circuit of this deck, created by the library from its own
        flat circuit."""
        ch = C.c_void_p()
        _check(self.lib.cb_netlist_circuit(self.handle, C.byref(ch)))
        return Circuit(self.fc, (), cache_dir, _handle=ch)

    def close(self):
        if self.handle:
            self.lib.cb_netlist_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
