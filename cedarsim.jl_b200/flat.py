"""Flat circuit description: the neutral form handed across the C ABI.

A `FlatCircuit` is what the host produces by flattening a netlist for one
`CircuitSweep`: MNA unknowns, built-in device instances, source waveforms,
Verilog-A model instances, and -- for every numeric device parameter -- either a
constant or a reference to one per-instance parameter column (what `ParamSim`
turns into a runtime parameter, reference src/circuitodesystem.jl:66-97).

`pack()` lowers it to the `cb_flat_circuit` struct of include/cedarb200.h.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

# ---------------------------------------------------------------- ctypes mirror of cedarb200.h


class cb_pref(C.Structure):
    _fields_ = [("value", C.c_double), ("col", C.c_int32), ("_pad", C.c_int32)]


class cb_device(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("n", C.c_int32 * 4),
        ("branch", C.c_int32),
        ("wave", C.c_int32),
        ("_pad", C.c_int32),
        ("value", cb_pref),
        ("mult", C.c_double),
    ]


class cb_wave(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("has_dc", C.c_int32),
        ("dc", cb_pref),
        ("npts", C.c_int32),
        ("_pad", C.c_int32),
        ("t", C.POINTER(C.c_double)),
        ("y", C.POINTER(cb_pref)),
        ("v", cb_pref * 7),
        ("ac_mag", C.c_double),
    ]


class cb_va_model(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("nterm", C.c_int32),
        ("nparam", C.c_int32),
        ("ncache", C.c_int32),
        ("nj", C.c_int32),
        ("jrow", C.POINTER(C.c_int32)),
        ("jcol", C.POINTER(C.c_int32)),
        ("host_setup", C.c_void_p),
        ("host_eval", C.c_void_p),
        ("n_noise", C.c_int32),
        ("ncache_n", C.c_int32),
        ("noise_pos", C.POINTER(C.c_int32)),
        ("noise_neg", C.POINTER(C.c_int32)),
        ("host_setupn", C.c_void_p),
        ("host_noise", C.c_void_p),
        ("linear", C.c_int32),
        ("ncache_v", C.c_int32),
        ("host_setupv", C.c_void_p),
        ("host_evalv", C.c_void_p),
    ]


class cb_va_inst(C.Structure):
    _fields_ = [
        ("model", C.c_int32),
        ("_pad", C.c_int32),
        ("term", C.POINTER(C.c_int32)),
        ("par", C.POINTER(cb_pref)),
        ("given", C.POINTER(C.c_uint8)),
        ("mult", C.c_double),
    ]


class cb_flat_circuit(C.Structure):
    _fields_ = [
        ("n_unknowns", C.c_int32),
        ("n_nodes", C.c_int32),
        ("n_params", C.c_int32),
        ("n_devices", C.c_int32),
        ("devices", C.POINTER(cb_device)),
        ("n_waves", C.c_int32),
        ("n_va_models", C.c_int32),
        ("waves", C.POINTER(cb_wave)),
        ("va_models", C.POINTER(cb_va_model)),
        ("n_va_insts", C.c_int32),
        ("n_outputs", C.c_int32),
        ("va_insts", C.POINTER(cb_va_inst)),
        ("outputs", C.POINTER(C.c_int32)),
    ]


class cb_options(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("abi_version", C.c_uint32),
        ("temp", cb_pref),
        ("gmin", cb_pref),
        ("reltol", C.c_double),
        ("vabstol", C.c_double),
        ("iabstol", C.c_double),
        ("nr_reltol", C.c_double),
        ("nr_vabstol", C.c_double),
        ("nr_iabstol", C.c_double),
        ("dc_abstol", C.c_double),
        ("dv_max", C.c_double),
        ("max_newton_dc", C.c_int32),
        ("max_newton_tran", C.c_int32),
        ("method", C.c_int32),
        ("fixed_step", C.c_int32),
        ("dt", C.c_double),
        ("dt_min", C.c_double),
        ("dt_max", C.c_double),
        ("gmin_steps", C.c_int32),
        ("skip_dc", C.c_int32),
        ("nr_rate_test", C.c_int32),
        ("value_rounds", C.c_int32),
        ("mixed_rounds", C.c_int32),
        ("source_steps", C.c_int32),
        ("t0_reinit", C.c_int32),
        ("pivot_repair", C.c_int32),
        ("pivot_growth_max", C.c_double),
    ]


class cb_stats(C.Structure):
    _fields_ = [
        ("newton_iters", C.c_int64),
        ("lu_factors", C.c_int64),
        ("steps_accepted", C.c_int64),
        ("steps_rejected", C.c_int64),
        ("rounds", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("solve_seconds", C.c_double),
        ("h2d_seconds", C.c_double),
        ("d2h_seconds", C.c_double),
        ("eval_seconds", C.c_double),
        ("newton_seconds", C.c_double),
        ("value_rounds", C.c_int64),
        ("full_iters", C.c_int64),
        ("evalv_seconds", C.c_double),
        ("newtonv_seconds", C.c_double),
        ("pivot_fallbacks", C.c_int64),
        ("dc_source_stepped", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


DEV_R, DEV_C, DEV_L, DEV_VSRC, DEV_ISRC, DEV_VCVS, DEV_VCCS = range(7)
W_DC, W_PWL, W_PULSE, W_SIN = range(4)
METHOD_BE, METHOD_TRAP, METHOD_BDF2 = range(3)
ST_SUCCESS, ST_MAXITERS, ST_INITIAL_FAILURE, ST_DT_LESS_THAN_MIN, ST_UNSTABLE = range(5)
RETCODES = {0: "Success", 1: "MaxIters", 2: "InitialFailure", 3: "DtLessThanMin", 4: "Unstable"}

# ---------------------------------------------------------------- python-side description


@dataclass(frozen=True)
class Col:
    """Reference to per-instance parameter column `index` of params[P][B]."""

    index: int


Value = Union[float, int, Col]


def _pref(v: Value) -> cb_pref:
    if isinstance(v, Col):
        return cb_pref(0.0, int(v.index), 0)
    return cb_pref(float(v), -1, 0)


@dataclass
class Wave:
    kind: int = W_DC
    dc: Optional[Value] = None  # None = not given (`something(dc, tran, 0)`)
    t: Sequence[float] = ()
    y: Sequence[Value] = ()
    v: Sequence[Value] = ()  # PULSE: v1 v2 td tr tf pw per | SIN: vo va freq td theta phase ncycles
    ac: float = 0.0          # small-signal magnitude |ac| (src/simpledevices.jl:292-294; the phase is ignored there too)


@dataclass
class Device:
    kind: int
    name: str
    nodes: Sequence[int]  # p, n[, cp, cn] unknown indices (-1 ground)
    value: Value = 0.0
    branch: int = -1
    wave: int = -1
    mult: float = 1.0


@dataclass
class VAModelShape:
    """Shape of one compiled Verilog-A model (see va/compiler.py)."""

    name: str
    terminals: List[str]
    params: List[str]
    ncache: int
    jrow: List[int]
    jcol: List[int]
    host_setup: int = 0  # function addresses, CPU oracle only
    host_eval: int = 0
    # noise sources (va/compiler.py noise variant): independent current sources between two terminals
    ncache_n: int = 0
    noise_pos: List[int] = field(default_factory=list)   # terminal index, -1 = ground
    noise_neg: List[int] = field(default_factory=list)
    host_setupn: int = 0
    host_noise: int = 0
    ncache_v: int = 0        # value-only variant on the host (the oracle's chord iterations)
    host_setupv: int = 0
    host_evalv: int = 0
    branch_terms: List[int] = field(default_factory=list)   # terminals that are branch currents (V() <+ branches, I() probes)
    linear: bool = False   # every Jacobian entry is bias-independent: no Newton step limiting on its account


def shape_of(cm) -> VAModelShape:
    """Shape of a va.compiler.CompiledModel without host function addresses (all the engine needs)."""
    # (the counts include the model's uniform slots: rows are over-allocated by them on the GPU, where they live in a table)
    return VAModelShape(cm.name, list(cm.terminals), list(cm.params), cm.ncache + getattr(cm, "nuni", 0), list(cm.jrow), list(cm.jcol),
                        ncache_n=getattr(cm, "ncache_n", 0) + getattr(cm, "nuni_n", 0),
                        noise_pos=[int(s[0]) for s in getattr(cm, "noise_sources", [])],
                        noise_neg=[int(s[1]) for s in getattr(cm, "noise_sources", [])],
                        branch_terms=list(getattr(cm, "branch_terms", [])), linear=bool(getattr(cm, "linear", False)))


@dataclass
class VAInst:
    model: int
    name: str
    term: Sequence[int]
    par: Dict[int, Value]  # param index -> value, only the given ones
    mult: float = 1.0


@dataclass
class FlatCircuit:
    node_names: List[str] = field(default_factory=list)  # unknown index -> name (voltages first)
    branch_names: List[str] = field(default_factory=list)
    devices: List[Device] = field(default_factory=list)
    waves: List[Wave] = field(default_factory=list)
    va_models: List[VAModelShape] = field(default_factory=list)
    va_insts: List[VAInst] = field(default_factory=list)
    param_names: List[str] = field(default_factory=list)  # P columns
    outputs: List[int] = field(default_factory=list)
    _node_index: Dict[str, int] = field(default_factory=dict)
    _finalized: bool = False
    va_branches: List[str] = field(default_factory=list)   # names of branch-current unknowns of Verilog-A instances
    aliases: Dict[str, str] = field(default_factory=dict)   # subcircuit port 'x1.pos' -> the parent's net ('vcc' or '0')

    # ---- construction helpers -------------------------------------------------
    def node(self, name: str) -> int:
        """Unknown index of net `name`; '0' and 'gnd' are ground (src/spectre.jl:736-749)."""
        name = str(name).lower()
        if name in ("0", "gnd", "gnd!"):
            return -1
        if name not in self._node_index:
            assert not self._finalized, "cannot add nodes after branches were allocated"
            self._node_index[name] = len(self.node_names)
            self.node_names.append(name)
        return self._node_index[name]

    def param(self, name: str) -> Col:
        if name not in self.param_names:
            self.param_names.append(name)
        return Col(self.param_names.index(name))

    def _dev(self, kind, name, nodes, value=0.0, wave=None, mult=1.0, branch=False):
        d = Device(kind, name.lower(), [self.node(n) if isinstance(n, str) else n for n in nodes],
                   value, -1, -1, float(mult))
        if wave is not None:
            self.waves.append(wave)
            d.wave = len(self.waves) - 1
        if branch:
            d.branch = -2  # allocated in finalize(), after all nodes are known
        self.devices.append(d)
        return d

    def resistor(self, name, p, n, r, m=1.0):
        return self._dev(DEV_R, name, (p, n), r, mult=m)

    def capacitor(self, name, p, n, c, m=1.0):
        return self._dev(DEV_C, name, (p, n), c, mult=m)

    def inductor(self, name, p, n, l, m=1.0):
        return self._dev(DEV_L, name, (p, n), l, mult=m, branch=True)

    def vsource(self, name, p, n, wave: Union[Wave, Value], m=1.0):
        if not isinstance(wave, Wave):
            wave = Wave(W_DC, dc=wave)
        return self._dev(DEV_VSRC, name, (p, n), 0.0, wave=wave, mult=m, branch=True)

    def isource(self, name, p, n, wave: Union[Wave, Value], m=1.0):
        if not isinstance(wave, Wave):
            wave = Wave(W_DC, dc=wave)
        return self._dev(DEV_ISRC, name, (p, n), 0.0, wave=wave, mult=m)

    def vcvs(self, name, p, n, cp, cn, gain, m=1.0):
        return self._dev(DEV_VCVS, name, (p, n, cp, cn), gain, mult=m, branch=True)

    def vccs(self, name, p, n, cp, cn, gain, m=1.0):
        return self._dev(DEV_VCCS, name, (p, n, cp, cn), gain, mult=m)

    def va_model(self, shape: VAModelShape) -> int:
        for i, m in enumerate(self.va_models):
            if m.name == shape.name:
                return i
        self.va_models.append(shape)
        return len(self.va_models) - 1

    def va_instance(self, name, model: int, ports: Sequence[Union[str, int]], params: Dict[str, Value], m=1.0):
        """Instantiate VA model `model`; internal nodes are created as '<name>.<node>'."""
        shape = self.va_models[model]
        name = name.lower()
        nports = len(ports)
        term = [self.node(p) if isinstance(p, str) else p for p in ports]
        for k, internal in enumerate(shape.terminals[nports:], start=nports):
            if k in shape.branch_terms:
                # a branch current of the device (its own MNA unknown): allocated after the node voltages by finalize();
                # until then a placeholder -2 - <index into va_branches>
                self.va_branches.append(f"{name}.{internal.lower()}")
                term.append(-2 - (len(self.va_branches) - 1))
            else:
                term.append(self.node(f"{name}.{internal}"))
        lut = {p.lower(): i for i, p in enumerate(shape.params)}
        par = {}
        for k, v in params.items():
            if k.lower() not in lut:
                raise KeyError(f"model {shape.name} has no parameter {k!r}")
            par[lut[k.lower()]] = v
        self.va_insts.append(VAInst(model, name, term, par, float(m)))
        return self.va_insts[-1]

    def finalize(self):
        """Allocate branch-current unknowns after the node voltages."""
        if self._finalized:
            return self
        self._finalized = True
        nn = len(self.node_names)
        for d in self.devices:
            if d.branch == -2:
                d.branch = nn + len(self.branch_names)
                self.branch_names.append(d.name + ".i")
        slot = {}
        for k, bname in enumerate(self.va_branches):
            slot[-2 - k] = nn + len(self.branch_names)
            self.branch_names.append(bname)
        for vi in self.va_insts:
            vi.term = [slot.get(t, t) for t in vi.term]
        return self

    @property
    def n_nodes(self):
        return len(self.node_names)

    @property
    def n_unknowns(self):
        self.finalize()
        return len(self.node_names) + len(self.branch_names)

    def unknown(self, name: str) -> int:
        """Index of a node voltage ('q', 'node_q') or branch current ('v1.i', 'v1.I')."""
        self.finalize()
        key = name.lower().replace(" ", "")   # `x1.I(p, n)` as the reference prints it == `x1.i(p,n)`
        head, _, last = key.rpartition(".")
        if last.startswith("node_"):             # sys.node_q, sys.x1.node_pos (src/spectre.jl:736-749)
            key = (head + "." if head else "") + last[5:]
        seen = set()
        while key in self.aliases and key not in seen:   # a subcircuit port is the parent's net (test/alias.jl)
            seen.add(key)
            key = self.aliases[key]
        if key in ("0", "gnd", "gnd!"):
            raise KeyError(f"{name!r} is the ground net (0 V, not an unknown)")
        if key in self._node_index:
            return self._node_index[key]
        if key in self.branch_names:
            return len(self.node_names) + self.branch_names.index(key)
        raise KeyError(f"no unknown named {name!r}")

    def set_outputs(self, names: Sequence[Union[str, int]]):
        self.outputs = [n if isinstance(n, int) else self.unknown(n) for n in names]
        return self

    # ---- lowering ---------------------------------------------------------------
    def pack(self) -> "PackedCircuit":
        self.finalize()
        keep = []
        devs = (cb_device * max(1, len(self.devices)))()
        for i, d in enumerate(self.devices):
            n = list(d.nodes) + [-1] * (4 - len(d.nodes))
            devs[i] = cb_device(d.kind, (C.c_int32 * 4)(*n), d.branch, d.wave, 0, _pref(d.value), d.mult)
        waves = (cb_wave * max(1, len(self.waves)))()
        for i, w in enumerate(self.waves):
            cw = cb_wave()
            cw.kind = w.kind
            cw.has_dc = 0 if w.dc is None else 1
            cw.dc = _pref(0.0 if w.dc is None else w.dc)
            cw.npts = len(w.t)
            if w.kind == W_PWL:
                if len(w.t) != len(w.y):
                    raise ValueError("PWL must have an equal number of x and y values")
                ts = (C.c_double * len(w.t))(*[float(x) for x in w.t])
                ys = (cb_pref * len(w.y))(*[_pref(y) for y in w.y])
                keep += [ts, ys]
                cw.t = C.cast(ts, C.POINTER(C.c_double))
                cw.y = C.cast(ys, C.POINTER(cb_pref))
            vv = list(w.v)
            if w.kind == W_PULSE:
                vv = vv + [math.inf] * (7 - len(vv)) if len(vv) >= 5 else vv
                if len(vv) != 7:
                    raise ValueError("PULSE needs v1 v2 td tr tf [pw [per]]")
                for k in range(2, 7):
                    if isinstance(vv[k], Col):
                        raise ValueError("PULSE timing parameters cannot be swept (shared breakpoints)")
                # a period <= 0 means "no period" (one pulse), as in SPICE; fmod(t, 0) would be NaN
                if not float(vv[6]) > 0.0:
                    vv[6] = math.inf
                if any(not (0.0 <= float(vv[k]) < math.inf) for k in range(2, 6)):
                    raise ValueError("PULSE td / tr / tf / pw must be finite and >= 0")
            elif w.kind == W_SIN:
                dflt = [0.0, 0.0, 1.0, 0.0, 0.0, 0.0, math.inf]
                vv = vv + dflt[len(vv):]
            else:
                vv = [0.0] * 7
            for k in range(7):
                cw.v[k] = _pref(vv[k])
            cw.ac_mag = abs(float(w.ac))
            waves[i] = cw
        models = (cb_va_model * max(1, len(self.va_models)))()
        for i, m in enumerate(self.va_models):
            jr = (C.c_int32 * max(1, len(m.jrow)))(*m.jrow)
            jc = (C.c_int32 * max(1, len(m.jcol)))(*m.jcol)
            npos = (C.c_int32 * max(1, len(m.noise_pos)))(*m.noise_pos)
            nneg = (C.c_int32 * max(1, len(m.noise_neg)))(*m.noise_neg)
            keep += [jr, jc, npos, nneg]
            models[i] = cb_va_model(m.name.encode(), len(m.terminals), len(m.params), m.ncache, len(m.jrow),
                                    C.cast(jr, C.POINTER(C.c_int32)), C.cast(jc, C.POINTER(C.c_int32)),
                                    m.host_setup or None, m.host_eval or None, len(m.noise_pos), m.ncache_n,
                                    C.cast(npos, C.POINTER(C.c_int32)), C.cast(nneg, C.POINTER(C.c_int32)),
                                    m.host_setupn or None, m.host_noise or None, 1 if m.linear else 0, m.ncache_v,
                                    m.host_setupv or None, m.host_evalv or None)
        insts = (cb_va_inst * max(1, len(self.va_insts)))()
        for i, vi in enumerate(self.va_insts):
            shape = self.va_models[vi.model]
            npar = len(shape.params)
            term = (C.c_int32 * len(vi.term))(*vi.term)
            par = (cb_pref * max(1, npar))()
            given = (C.c_uint8 * max(1, npar))()
            for k, v in vi.par.items():
                par[k] = _pref(v)
                given[k] = 1
            keep += [term, par, given]
            insts[i] = cb_va_inst(vi.model, 0, C.cast(term, C.POINTER(C.c_int32)),
                                  C.cast(par, C.POINTER(cb_pref)), C.cast(given, C.POINTER(C.c_uint8)), vi.mult)
        outs = (C.c_int32 * max(1, len(self.outputs)))(*self.outputs)
        fc = cb_flat_circuit()
        fc.n_unknowns = self.n_unknowns
        fc.n_nodes = self.n_nodes
        fc.n_params = len(self.param_names)
        fc.n_devices = len(self.devices)
        fc.devices = C.cast(devs, C.POINTER(cb_device))
        fc.n_waves = len(self.waves)
        fc.n_va_models = len(self.va_models)
        fc.waves = C.cast(waves, C.POINTER(cb_wave))
        fc.va_models = C.cast(models, C.POINTER(cb_va_model))
        fc.n_va_insts = len(self.va_insts)
        fc.n_outputs = len(self.outputs)
        fc.va_insts = C.cast(insts, C.POINTER(cb_va_inst))
        fc.outputs = C.cast(outs, C.POINTER(C.c_int32))
        keep += [devs, waves, models, insts, outs]
        return PackedCircuit(fc, keep)


@dataclass
class PackedCircuit:
    struct: cb_flat_circuit
    keepalive: list

    def ref(self):
        return C.byref(self.struct)


def nodeset_vector(fc: "FlatCircuit", nodeset: Dict[str, float]) -> np.ndarray:
    """Initial guess x0[N] from named node voltages.  Internal nodes of Verilog-A devices that sit
    behind a series resistance (BSIM `di`/`si` behind `d`/`s`) start at their port's voltage."""
    x0 = np.zeros(fc.n_unknowns)
    for n, v in nodeset.items():
        x0[fc.unknown(n)] = v
    for vi in fc.va_insts:
        terms = fc.va_models[vi.model].terminals
        for k, tname in enumerate(terms):
            if tname.endswith("i") and tname[:-1] in terms and vi.term[k] >= 0:
                port = vi.term[terms.index(tname[:-1])]
                x0[vi.term[k]] = x0[port] if port >= 0 else 0.0
    return x0


def params_matrix(cols: Sequence[np.ndarray]) -> np.ndarray:
    """Stack per-instance parameter columns into the C-ABI layout [P][B]."""
    if len(cols) == 0:
        return np.zeros((0, 0))
    return np.ascontiguousarray(np.stack([np.asarray(c, dtype=np.float64) for c in cols], axis=0))


def save_flatckt(fc: "FlatCircuit", models: Sequence, path: str) -> None:
    """Write `fc` and the generated CUDA C of its Verilog-A `models` (va.compiler.CompiledModel) as a "flatckt" file,
    the input of cb_circuit_load (include/cedarb200.h): what a binding without a front end of its own hands to the
    engine (ext/CedarSimB200Ext.jl).  Layout: see cb_circuit_load in csrc/cedarb200.cu."""
    import struct as S
    from . import engine
    pk = fc.pack()
    f = pk.struct
    out = [b"CBFC", S.pack("<I", 1),
           S.pack("<8i", f.n_unknowns, f.n_nodes, f.n_params, f.n_devices, f.n_waves, f.n_va_models, f.n_va_insts, f.n_outputs)]
    raw = lambda obj: bytes(memoryview(obj).cast("B")) if not isinstance(obj, C.Structure) else C.string_at(C.addressof(obj), C.sizeof(obj))  # noqa: E731
    for i in range(f.n_devices):
        out.append(raw(f.devices[i]))
    for i in range(f.n_waves):
        w = f.waves[i]
        out += [S.pack("<2i", w.kind, w.has_dc), raw(w.dc), S.pack("<i", w.npts if w.kind == W_PWL else 0)]
        if w.kind == W_PWL:
            out.append(S.pack(f"<{w.npts}d", *[w.t[k] for k in range(w.npts)]))
            out += [raw(w.y[k]) for k in range(w.npts)]
        out += [raw(w.v[k]) for k in range(7)]
        out.append(S.pack("<d", w.ac_mag))
    for i in range(f.n_va_models):
        m = f.va_models[i]
        out += [S.pack("<I", len(m.name)), m.name, S.pack("<4i", m.nterm, m.nparam, m.ncache, m.nj),
                S.pack(f"<{m.nj}i", *[m.jrow[k] for k in range(m.nj)]), S.pack(f"<{m.nj}i", *[m.jcol[k] for k in range(m.nj)]),
                S.pack("<2i", m.n_noise, m.ncache_n),
                S.pack(f"<{m.n_noise}i", *[m.noise_pos[k] for k in range(m.n_noise)]),
                S.pack(f"<{m.n_noise}i", *[m.noise_neg[k] for k in range(m.n_noise)]), S.pack("<i", m.linear)]
    for i in range(f.n_va_insts):
        v = f.va_insts[i]
        m = f.va_models[v.model]
        out += [S.pack("<i", v.model), S.pack(f"<{m.nterm}i", *[v.term[k] for k in range(m.nterm)])]
        out += [raw(v.par[k]) for k in range(m.nparam)]
        out += [bytes(bytearray(v.given[k] for k in range(m.nparam))), S.pack("<d", v.mult)]
    out.append(S.pack(f"<{f.n_outputs}i", *[f.outputs[k] for k in range(f.n_outputs)]))
    src = b""
    if fc.va_models:
        by_name = {cm.name: cm for cm in models}
        src = engine.cuda_source([by_name[m.name] for m in fc.va_models]).encode()
    out += [S.pack("<Q", len(src)), src]
    with open(path, "wb") as fh:
        fh.write(b"".join(out))
