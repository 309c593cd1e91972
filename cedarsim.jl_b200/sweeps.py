"""Sweep API: mirror of reference src/sweeps.jl, with the batched engine behind `dc_` / `tran_`.

    Sweep / ProductSweep / TandemSweep / SerialSweep / sweepify / split_axes / sweepvars /
    find_param_ranges                        src/sweeps.jl:52-60, 98-146, 175-354, 507-546
    CircuitSweep                             src/sweeps.jl:390-435
    dc_(cs)   == dc!(cs::CircuitSweep)       src/sweeps.jl:437-448, 471-486
    tran_(cs) == tran!(cs::CircuitSweep, tspan)  src/sweeps.jl:450-463, 488-502 (with the defects
                                             listed in SURVEY.md 3.2 fixed: tspan is an argument)

    sensitivities_(cs, wrt)                  test/sensitivity.jl:14-68 (forward sensitivities over the ParamSim's
                                             parameters), as a batched difference stencil: extra sweep points

Iteration order and shapes are the reference's: a point is a tuple of (name, value) pairs sorted by
name; a product varies its FIRST axis fastest and `size(cs)` is the tuple of axis lengths (Julia
column-major), tandem zips, serial concatenates with None ("keep default") for inactive names.

Instead of the reference's serial `broadcast` over points (src/sweeps.jl:473,490) the whole sweep
is flattened once into per-point parameter columns and solved in one batched call on the GPU(s).
"""
from __future__ import annotations

import itertools
import threading
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Set, Tuple, Union

import numpy as np

from . import engine
from .flat import FlatCircuit, RETCODES
from .netlist import Flattened, Netlist, flatten

Point = Tuple[Tuple[str, object], ...]


def _expand(items) -> Point:
    """flatten nested point tuples and sort by name (src/sweeps.jl:131-137)"""
    out: List[Tuple[str, object]] = []

    def rec(x):
        if isinstance(x, tuple) and len(x) == 2 and isinstance(x[0], str):
            out.append(x)
        else:
            for y in x:
                rec(y)

    rec(items)
    return tuple(sorted(out, key=lambda kv: kv[0]))


class SweepBase:
    shape: Tuple[int, ...] = ()

    def __len__(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n

    def size(self, d: Optional[int] = None):
        """size(s) / size(s, d) with Julia's 1-based d; missing dimensions are 1"""
        if d is None:
            return self.shape
        return self.shape[d - 1] if 1 <= d <= len(self.shape) else 1

    def sweepvars(self) -> Set[str]:
        raise NotImplementedError

    def __iter__(self) -> Iterator[Point]:
        raise NotImplementedError

    def first(self) -> Point:
        for p in self:
            return p
        return ()

    def columns(self) -> Dict[str, np.ndarray]:
        """{name: float array over all points in iteration order}; None -> NaN"""
        names = sorted(self.sweepvars())
        cols = {n: np.full(len(self), np.nan) for n in names}
        for i, pt in enumerate(self):
            for k, v in pt:
                if v is not None:
                    cols[k][i] = float(v)
        return cols

    def children(self) -> Sequence["SweepBase"]:
        return ()


class Sweep(SweepBase):
    """Sweep("R1", values) / Sweep(R1=values); a scalar becomes a one-point sweep."""

    def __init__(self, selector=None, values=None, **kw):
        if isinstance(selector, SweepBase):
            raise TypeError("use sweepify() to pass sweeps through")
        if selector is None:
            if len(kw) != 1:
                raise ValueError("`Sweep` takes a single variable at a time!")
            (selector, values), = kw.items()
        elif isinstance(selector, tuple) and values is None:
            selector, values = selector
        self.selector = str(selector)
        if np.isscalar(values):
            values = [values]
        self.values = list(values)
        self.shape = (len(self.values),)

    def __eq__(self, other):
        return isinstance(other, Sweep) and self.selector == other.selector and self.values == other.values

    def __hash__(self):
        return hash((self.selector, tuple(self.values)))

    def __iter__(self):
        for v in self.values:
            yield ((self.selector, v),)

    def sweepvars(self):
        return {self.selector}

    def __repr__(self):
        if len(self.values) > 1:
            return f"Sweep of {self.selector} with {len(self.values)} values over [{min(self.values)} .. {max(self.values)}]"
        return f"Sweep of {self.selector} set to {self.values[0]}"


def _as_sweeps(args, kwargs) -> List[SweepBase]:
    out: List[SweepBase] = []
    for a in args:
        out.append(a if isinstance(a, SweepBase) else Sweep(a) if not isinstance(a, dict) else None)
        if out[-1] is None:
            out.pop()
            out.extend(Sweep(k, v) for k, v in a.items())
    out.extend(Sweep(k, v) for k, v in kwargs.items())
    return out


class _Product(SweepBase):
    def __init__(self, its: List[SweepBase]):
        self.iterators = its
        self.shape = tuple(itertools.chain.from_iterable(it.shape for it in its))

    def __iter__(self):
        if not self.iterators:
            yield ()
            return
        lists = [list(it) for it in self.iterators]
        # first iterator varies fastest (Base.Iterators.product)
        for combo in itertools.product(*reversed(lists)):
            yield _expand(tuple(reversed(combo)))

    def __len__(self):
        n = 1
        for it in self.iterators:
            n *= len(it)
        return n

    def sweepvars(self):
        return set().union(*[it.sweepvars() for it in self.iterators]) if self.iterators else set()

    def children(self):
        return self.iterators


class _Tandem(SweepBase):
    def __init__(self, its: List[SweepBase]):
        lens = [len(it) for it in its]
        if any(n != lens[0] for n in lens):
            raise ValueError("TandemSweep requires all sweeps be of the same length!")
        self.iterators = its
        self.shape = (lens[0],) if lens else ()

    def __iter__(self):
        for combo in zip(*self.iterators):
            yield _expand(combo)

    def sweepvars(self):
        return set().union(*[it.sweepvars() for it in self.iterators])

    def children(self):
        return self.iterators


class _Serial(SweepBase):
    def __init__(self, its: List[SweepBase]):
        self.iterators = its
        self.vars = set().union(*[it.sweepvars() for it in its]) if its else set()
        self.shape = (sum(len(it) for it in its),)

    def __iter__(self):
        for it in self.iterators:
            for pt in it:
                m = {v: None for v in self.vars}
                m.update(dict(_expand(pt)))
                yield tuple(sorted(m.items()))

    def sweepvars(self):
        return set(self.vars)

    def children(self):
        return self.iterators


def ProductSweep(*args, **kwargs) -> SweepBase:
    its = _as_sweeps(args, kwargs)
    if len(its) == 1:
        return its[0]
    return _Product(its)


def TandemSweep(*args, **kwargs) -> SweepBase:
    its = _as_sweeps(args, kwargs)
    if len(its) == 1:
        return its[0]
    return _Tandem(its)


def SerialSweep(*args, **kwargs) -> SweepBase:
    its = _as_sweeps(args, kwargs)
    if len(its) == 1:
        return its[0]
    return _Serial(its)


def sweepvars(*sweeps) -> Set[str]:
    out: Set[str] = set()
    for s in sweeps:
        out |= s.sweepvars()
    return out


def sweepify(x) -> SweepBase:
    """lists -> SerialSweep, dicts -> ProductSweep, (name, values) -> Sweep (src/sweeps.jl:349-354)"""
    if isinstance(x, SweepBase):
        return x
    if isinstance(x, dict):
        return ProductSweep(**x)
    if isinstance(x, tuple) and len(x) == 2 and isinstance(x[0], str):
        return Sweep(x[0], x[1])
    if isinstance(x, (list, tuple)):
        return SerialSweep(*[sweepify(y) for y in x])
    raise TypeError(f"cannot turn {x!r} into a sweep")


def split_axes(sweep: SweepBase, axes: Iterable[str]):
    """split a ProductSweep into (outer, inner) products; inner holds `axes` (src/sweeps.jl:98-129)"""
    if not isinstance(sweep, _Product):
        raise ValueError("split_axes only works with ProductSweep objects!")
    axes = list(axes)
    idx = []
    for ax in axes:
        found = [i for i, it in enumerate(sweep.iterators) if isinstance(it, Sweep) and it.selector == ax]
        if not found:
            raise ValueError(f"Unable to find product axis matching '{ax}'")
        idx.append(found[0])
    inner = _Product([sweep.iterators[i] for i in idx])
    outer = _Product([it for i, it in enumerate(sweep.iterators) if i not in idx])
    return outer, inner


def find_param_ranges(sweep: SweepBase) -> Dict[str, Tuple[float, float, int]]:
    """(min, max, number of values) explored along each name (src/sweeps.jl:507-546)"""
    acc: Dict[str, List[Tuple[float, float, int]]] = {}

    def rec(it):
        if isinstance(it, Sweep):
            acc.setdefault(it.selector, []).append((min(it.values), max(it.values), len(it.values)))
        else:
            for c in it.children():
                rec(c)

    rec(sweep)
    out = {}
    for name, ranges in acc.items():
        lo, hi, n = ranges[0]
        for a, b, m in ranges[1:]:
            lo, hi, n = min(lo, a), max(hi, b), n + m
        out[name] = (lo, hi, n)
    return out


# ---------------------------------------------------------------- circuit sweep + solutions


class ScopeRef:
    """cs.sys.node_q / cs.sys.v1.I / cs.sys.x1.r1.I -- names results as the reference does
    (src/simulate_ir.jl:79-91, src/spectre.jl:736-749, test/sweep.jl:336-339,363-369)."""

    def __init__(self, path: str = ""):
        object.__setattr__(self, "_path", path)

    def __getattr__(self, name: str):
        return ScopeRef(f"{self._path}.{name}" if self._path else name)

    def __getitem__(self, name: str):
        return getattr(self, name)

    def __repr__(self):
        return f"sys.{self._path}"


def _resolve(fc: FlatCircuit, ref: Union[str, ScopeRef]) -> int:
    key = ref._path if isinstance(ref, ScopeRef) else str(ref)
    return fc.unknown(key.lower())


class _Observable:
    """A named result: an unknown, or a branch observable of a two-terminal primitive reconstructed from the node
    voltages the way the reference does (src/simulate_ir.jl:112-120: `<inst>.V` = V(net+) - V(net-), `<inst>.I` flows
    net+ -> net- through the device; src/simpledevices.jl:62-77 I = V/R, :99-109 I = C dV/dt)."""

    def __init__(self, fc: FlatCircuit, ref: Union[str, ScopeRef]):
        from .flat import DEV_C, DEV_R
        key = (ref._path if isinstance(ref, ScopeRef) else str(ref)).lower()
        self.key, self.terms, self.div, self.cap = key, [], None, None
        try:
            self.terms = [(1.0, fc.unknown(key))]
            return
        except KeyError as e:
            if "ground net" in str(e):   # sys.x1.node_neg of a port tied to ground: identically 0 (test/alias.jl:33)
                self.terms = []
                return
        if not (key.endswith(".i") or key.endswith(".v")):
            raise KeyError(f"no unknown or observable named {key!r}")
        dev = next((d for d in fc.devices if d.name == key[:-2]), None)
        if dev is None:
            raise KeyError(f"no device named {key[:-2]!r}")
        self.terms = [(sg, n) for sg, n in ((1.0, dev.nodes[0]), (-1.0, dev.nodes[1])) if n >= 0]
        if key.endswith(".i"):
            if dev.kind == DEV_R:
                self.div = dev.value
            elif dev.kind == DEV_C:
                self.cap = dev.value
            else:
                raise KeyError(f"{key!r}: the current of this device kind is not reconstructible from node voltages")

    def unknowns(self) -> List[int]:
        return [u for _, u in self.terms]

    def value(self, sol: "SweepSolution", pts) -> np.ndarray:
        """values for the points `pts` (slice or index): DC -> [..], transient -> [S, ..]"""
        y = sol.y
        acc = 0.0
        for sg, u in self.terms:
            if u not in sol.out_index:
                raise KeyError(f"{self.key!r} needs unknown {sol.fc.node_names[u] if u < sol.fc.n_nodes else u} among the outputs")
            acc = acc + sg * y[sol.out_index[u]][..., pts]
        acc = np.asarray(acc, dtype=float) if self.terms else np.zeros(np.shape(y[0][..., pts]))
        for val, op in ((self.div, "div"), (self.cap, "cap")):
            if val is None:
                continue
            from .flat import Col
            v = sol.cs.flat.params[val.index][pts] if isinstance(val, Col) else float(val)
            if op == "div":
                acc = acc / v
            elif sol.t is None:
                acc = acc * 0.0                       # DC: no current through a capacitor
            else:
                acc = np.gradient(acc, sol.t, axis=0) * v
        return acc


class PointSolution:
    """One element of the result array of dc_ / tran_."""

    def __init__(self, parent: "SweepSolution", index: int):
        self._p, self._i = parent, index

    @property
    def retcode(self) -> str:
        return RETCODES[int(self._p.status[self._i])]

    @property
    def params(self) -> Dict[str, float]:
        return self._p.point_params(self._i)

    @property
    def t(self) -> np.ndarray:
        return self._p.t

    def __getitem__(self, ref):
        v = _Observable(self._p.fc, ref).value(self._p, self._i)
        return float(v) if self._p.t is None else v

    def __call__(self, t, idxs=None):
        """sol(t; idxs=sys.node_q): linear interpolation of the saved waveform"""
        if self._p.t is None:
            raise TypeError("DC solutions are not functions of time")
        refs = idxs if isinstance(idxs, (list, tuple)) else [idxs]
        vals = [np.interp(t, self._p.t, self[r]) for r in refs]
        return vals if isinstance(idxs, (list, tuple)) else vals[0]


class SweepSolution:
    """Array of per-point solutions with `size(cs)` (column-major like the Julia result)."""

    def __init__(self, cs: "CircuitSweep", y: np.ndarray, status: np.ndarray, stats: dict, t: Optional[np.ndarray]):
        self.cs, self.fc, self.y, self.status, self.stats, self.t = cs, cs.flat.fc, y, status, stats, t
        self.shape = cs.shape
        self.out_index = {u: k for k, u in enumerate(self.fc.outputs)}

    def __len__(self):
        return len(self.status)

    def point_params(self, i: int) -> Dict[str, float]:
        return {k: (None if np.isnan(v[i]) else float(v[i])) for k, v in self.cs.columns.items()}

    def _linear(self, idx) -> int:
        if isinstance(idx, (int, np.integer)):
            return int(idx)
        return int(np.ravel_multi_index(tuple(idx), self.shape, order="F"))

    def __getitem__(self, idx) -> PointSolution:
        return PointSolution(self, self._linear(idx))

    def __iter__(self):
        for i in range(len(self)):
            yield PointSolution(self, i)

    def array(self, ref) -> np.ndarray:
        """values of one unknown over the whole sweep, shaped size(cs) (+ time axis last for tran)"""
        v = _Observable(self.fc, ref).value(self, slice(None))
        if self.t is None:
            return v.reshape(self.shape, order="F")
        return np.moveaxis(v, 0, -1).reshape(self.shape + (len(self.t),), order="F")

    def default_name_map(self) -> Dict[str, str]:
        """top-level nets among the outputs -> column name, `node_` stripped, ground left out (src/util.jl:239-260)"""
        out = {}
        for u in self.fc.outputs:
            if u < self.fc.n_nodes and "." not in self.fc.node_names[u]:
                out["node_" + self.fc.node_names[u]] = self.fc.node_names[u]
        return out

    def write_csv(self, file: str, index=0, name_map: Optional[Dict[str, str]] = None):
        """CSV.write(file, sol) of one sweep point (ext/CedarSimCSVExt.jl:13-19): column `t` followed by one column
        per entry of `name_map` (ScopeRef path -> column name; default: the top-level nets)."""
        if self.t is None:
            raise TypeError("write_csv needs a transient solution")
        name_map = name_map or self.default_name_map()
        pt = self[index]
        cols = [("t", self.t)] + [(name, pt[ref]) for ref, name in name_map.items()]
        with open(file, "w") as f:
            f.write(",".join(n for n, _ in cols) + "\n")
            for k in range(len(self.t)):
                f.write(",".join(repr(float(c[k])) for _, c in cols) + "\n")
        return file

    def plot_spec(self, index=0, name_map: Optional[Dict[str, str]] = None, title: Optional[str] = None) -> dict:
        """PlotlyLight.Plot(sol; name_map, title) of one sweep point (ext/CedarSimPlotlyLightExt.jl:11-35): one
        "scatter" / "lines" trace per entry of `name_map`, keys in sorted order, x = sol.t."""
        if self.t is None:
            raise TypeError("plot_spec needs a transient solution")
        name_map = name_map or self.default_name_map()
        pt = self[index]
        data = [{"x": [float(v) for v in self.t], "y": [float(v) for v in pt[ref]], "type": "scatter", "mode": "lines",
                 "name": name_map[ref]} for ref in sorted(name_map)]
        layout = {"template": "plotly"}
        if title is not None:
            layout["title"] = title
        return {"data": data, "layout": layout}

    def save_html(self, file: str, index=0, **kw) -> str:
        """Cobweb.save(sol, filename) (ext/CedarSimPlotlyLightExt.jl:37-46): a self-contained page with the plot."""
        import json
        spec = json.dumps(self.plot_spec(index, **kw))
        with open(file, "w") as f:
            f.write("<!DOCTYPE html><html><head><meta charset=\"utf-8\"><script src=\"https://cdn.plot.ly/plotly-latest.min.js\">"
                    "</script></head><body><div id=\"plot\"></div><script>const s = " + spec +
                    "; Plotly.newPlot(\"plot\", s.data, s.layout);</script></body></html>\n")
        return file

    @property
    def retcodes(self) -> np.ndarray:
        return self.status.reshape(self.shape, order="F")


class CircuitSweep:
    """CircuitSweep(circuit, sweep): compile once for the set of swept names, then solve all points
    in one batched call.  `circuit` is a parsed `Netlist` (or SPICE text), or a callable
    `builder(columns: dict, B: int) -> Flattened`.  front_end="native": SPICE text is read and flattened by the engine
    library itself (cb_netlist_*, csrc/spice_front.hpp: R C L V I E G X decks) instead of the Python front end."""

    def __init__(self, circuit, iterator, outputs: Optional[Sequence[str]] = None, devices: Optional[Sequence[int]] = None,
                 host: bool = False, include_dirs: Optional[Sequence[str]] = None, lang: str = "spice", front_end: str = "python"):
        self.iterator = sweepify(iterator)
        self.shape = self.iterator.shape
        self.circuit = circuit
        self._ctor = dict(outputs=outputs, host=host, include_dirs=include_dirs, lang=lang)
        self.columns = self.iterator.columns()
        B = len(self.iterator)
        if B == 0:   # the reference compiles from `first(iterator)` (src/sweeps.jl:414-417) and fails on an empty one as well
            raise ValueError("empty sweep: a CircuitSweep needs at least one point")
        self._native = None
        if isinstance(circuit, str) and front_end == "native":
            import math
            from .flat import Col
            nn = engine.NativeNetlist(circuit, self.columns, outputs, base_dir=(include_dirs[0] if include_dirs else None), lang=lang)
            opts = {}
            t = nn.option("temp")
            if "temp" in nn.fc.param_names:
                opts["temp"] = Col(nn.fc.param_names.index("temp"))
            elif not math.isnan(t):
                opts["temp"] = t
            self._native = nn
            circuit = lambda columns, B_: Flattened(nn.fc, nn.params, [], opts, None)   # noqa: E731
        elif front_end != "python":
            raise ValueError("front_end must be 'python' or 'native'")
        if isinstance(circuit, str):
            from .netlist import parse_netlist
            if lang == "spectre":
                from .spectre import parse_spectre
                circuit = parse_spectre(circuit, include_dirs=include_dirs)
            else:
                circuit = parse_netlist(circuit, include_dirs=include_dirs)   # include_dirs as solve_spice_code(...; include_dirs)
        if isinstance(circuit, Netlist):
            self.flat: Flattened = flatten(circuit, self.columns, B=B, outputs=outputs, host=host)
        elif callable(circuit):
            self.flat = circuit(self.columns, B)
        else:
            raise TypeError("circuit must be a Netlist, SPICE text or a builder callable")
        self.sys = ScopeRef()
        self.devices = list(devices) if devices is not None else [0]
        self._compiled: Optional[engine.Circuit] = None
        self._plans: List[Tuple[engine.Plan, slice]] = []
        self.x0: Optional[np.ndarray] = None

    def __len__(self):
        return len(self.iterator)

    def size(self, d: Optional[int] = None):
        return self.iterator.size(d)

    def sweepvars(self):
        return self.iterator.sweepvars()

    def __iter__(self):
        """yields the parameter assignment of each point (the reference yields ParamSim objects)"""
        return iter(self.iterator)

    def nodeset(self, **node_voltages):
        """initial guess for the DC Newton (warm start, cf. remake(prob, u0=...) src/sweeps.jl:474-477)"""
        from .flat import nodeset_vector
        self.x0 = nodeset_vector(self.flat.fc, node_voltages)
        return self

    # ---- engine plumbing: contiguous block of points per GPU, one host thread per device
    def _ensure_plans(self):
        if self._plans:
            return
        self._compiled = self._native.circuit() if self._native is not None else engine.Circuit(self.flat.fc, self.flat.models)
        # one plan over all the GPUs named in `devices` (cb_plan_create_multi: contiguous block of points per GPU, one
        # host thread per lane inside the library, results copied straight into the caller's arrays)
        plan = self._compiled.plan(len(self), devices=self.devices)
        plan.set_params(np.ascontiguousarray(self.flat.params) if self.flat.params.size else None)
        self._plans.append((plan, slice(0, len(self))))

    def _options(self, kw):
        opts = dict(kw)
        for f_ in ("temp", "gmin"):
            if f_ in self.flat.options and f_ not in opts:
                opts[f_] = self.flat.options[f_]
        if "abstol" in opts:      # reference kwargs (test/sweep.jl:333): abstol -> DC residual tolerance
            opts["dc_abstol"] = min(opts.pop("abstol"), 1e-10)
        return engine.default_options(**opts)

    def _run(self, fn):
        self._ensure_plans()
        results = [None] * len(self._plans)
        errors = []

        def work(k, plan, sl):
            try:
                plan.set_x0(self.x0 if self.x0 is None or self.x0.ndim == 1 else self.x0[:, sl])
                results[k] = fn(plan)
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        threads = [threading.Thread(target=work, args=(k, p, sl)) for k, (p, sl) in enumerate(self._plans)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results


# Host retry policy for the points a batched solve leaves unconverged (the reference's CedarDCOp
# restarts a failed initialisation up to ten times with a more robust algorithm, src/dcop.jl:53-94): the failed points
# alone are solved again as a small batch of their own, with progressively more patient options -- longer gmin / source
# stepping ladders and a tighter Newton step limit for the operating point, plain full Newton (no rate test, no chord
# iterations) and more iterations per step for the transient, finally backward Euler; every rung also switches on the
# partial-pivoting repair pass of k_lu for iterations the pivot-growth monitor flags (cb_options.pivot_repair).  Points that converge on a rung
# replace their entries; the others keep their status code.  `retry=False` switches it off, `retry=[dict, ...]` gives
# another ladder.
RETRY_LADDER = (
    dict(gmin_steps=20, source_steps=40, max_newton_dc=400, dv_max=0.25, nr_rate_test=0, value_rounds=0, max_newton_tran=50,
         pivot_repair=1),
    dict(gmin_steps=20, source_steps=100, max_newton_dc=1000, dv_max=0.1, nr_rate_test=0, value_rounds=0, max_newton_tran=100,
         method=0, pivot_repair=1),
)


def _retry_failed(cs: CircuitSweep, kw: dict, retry, solve, y: np.ndarray, status: np.ndarray, stats: dict, point_axis: int):
    """solve(plan, opts) -> (y_sub with the points on `point_axis`, status_sub).  Mutates y / status / stats in place."""
    ladder = RETRY_LADDER if retry is True else tuple(retry or ())
    stats["retried_points"] = stats["recovered_points"] = 0
    for rung in ladder:
        failed = np.flatnonzero(status != 0)
        if failed.size == 0:
            break
        stats["retried_points"] = max(stats["retried_points"], int(failed.size))
        opts = cs._options(dict(kw, **rung))
        plan = cs._compiled.plan(int(failed.size), devices=[cs.devices[0]])
        try:
            P = cs.flat.params
            plan.set_params(np.ascontiguousarray(P[:, failed]) if P.size else None)
            plan.set_x0(cs.x0 if cs.x0 is None or cs.x0.ndim == 1 else np.ascontiguousarray(cs.x0[:, failed]))
            y_sub, st_sub = solve(plan, opts)
        finally:
            plan.close()
        ok = st_sub == 0
        if ok.any():
            idx = [slice(None)] * y.ndim
            idx[point_axis] = failed[ok]
            sub = [slice(None)] * y.ndim
            sub[point_axis] = np.flatnonzero(ok)
            y[tuple(idx)] = y_sub[tuple(sub)]
            status[failed[ok]] = 0
            stats["recovered_points"] += int(ok.sum())


def dc_(cs: CircuitSweep, retry=True, **kw) -> SweepSolution:
    """dc!(cs): DC operating point of every sweep point.  retry: see RETRY_LADDER."""
    opts = cs._options(kw)
    res = cs._run(lambda plan: plan.dc(opts, want_full=False))
    y = np.concatenate([r[0] for r in res], axis=1)
    status = np.concatenate([r[2] for r in res])
    stats = _merge_stats([r[3] for r in res])
    if retry and status.any():
        def solve(plan, o):
            r = plan.dc(o, want_full=False)
            return r[0], r[2]
        _retry_failed(cs, kw, retry, solve, y, status, stats, point_axis=1)
    return SweepSolution(cs, y, status, stats, None)


def tran_(cs: CircuitSweep, tspan: Optional[Tuple[float, float]] = None, saveat=None, retry=True, **kw) -> SweepSolution:
    """tran!(cs, tspan): transient of every sweep point from its DC operating point.  retry: see RETRY_LADDER."""
    if tspan is None:
        if cs.flat.tran is None:
            raise ValueError("no tspan given and the netlist has no .tran card")
        tspan = (0.0, cs.flat.tran[1])
    t0, t1 = tspan
    if saveat is None:
        step = cs.flat.tran[0] if cs.flat.tran else (t1 - t0) / 100.0
        saveat = t0 + np.arange(int(round((t1 - t0) / step)) + 1) * step
    elif np.isscalar(saveat):
        saveat = t0 + np.arange(int(round((t1 - t0) / saveat)) + 1) * saveat
    saveat = np.asarray(saveat, dtype=float)
    opts = cs._options(kw)
    res = cs._run(lambda plan: plan.tran(t0, t1, saveat, opts))
    y = np.concatenate([r[0] for r in res], axis=2)
    status = np.concatenate([r[1] for r in res])
    stats = _merge_stats([r[2] for r in res])
    if retry and status.any():
        def solve(plan, o):
            r = plan.tran(t0, t1, saveat, o)
            return r[0], r[1]
        _retry_failed(cs, kw, retry, solve, y, status, stats, point_axis=2)
    return SweepSolution(cs, y, status, stats, saveat)


class FreqSolution:
    """Result of ac_ / noise_: per sweep point the small-signal response about its DC operating point.
    Mirrors how the reference's ACSol / NoiseSol are queried (src/ac.jl:257-284): `freqresp(sol, ref)` /
    `PSD(sol, ref)` give arrays shaped size(cs) + (len(freqs),)."""

    def __init__(self, cs: "CircuitSweep", y: np.ndarray, status: np.ndarray, stats: dict, freqs: np.ndarray, kind: str):
        self.cs, self.fc, self.y, self.status, self.stats, self.freqs, self.kind = cs, cs.flat.fc, y, status, stats, freqs, kind
        self.shape = cs.shape
        self.out_index = {u: k for k, u in enumerate(self.fc.outputs)}

    def __len__(self):
        return len(self.status)

    def array(self, ref) -> np.ndarray:
        o = self.out_index[_resolve(self.fc, ref)]
        return np.moveaxis(self.y[o], 0, -1).reshape(self.shape + (len(self.freqs),), order="F")

    def point(self, idx, ref) -> np.ndarray:
        i = int(idx) if isinstance(idx, (int, np.integer)) else int(np.ravel_multi_index(tuple(idx), self.shape, order="F"))
        return self.y[self.out_index[_resolve(self.fc, ref)], :, i]

    @property
    def retcodes(self) -> np.ndarray:
        return np.array([RETCODES[int(s)] for s in self.status]).reshape(self.shape, order="F")


def acdec(nd: int, fstart: float, fstop: float) -> np.ndarray:
    """`.ac dec nd fstart fstop` frequency vector in Hz (src/ac.jl:286-303)."""
    a, b = np.log10(fstart), np.log10(fstop)
    return 10.0 ** np.linspace(a, b, int(np.ceil((b - a) * nd)) + 1)


def ac_(cs: CircuitSweep, freqs, **kw) -> FreqSolution:
    """ac!(circ) for every sweep point (src/ac.jl:166-180; the reference has no CircuitSweep method): complex
    response of the outputs to the sources carrying `AC mag`.  `freqs` in Hz (the reference's freqresp takes
    omega = 2 pi f)."""
    freqs = np.asarray(freqs, dtype=float)
    opts = cs._options(kw)
    res = cs._run(lambda plan: plan.ac(freqs, opts))
    y = np.concatenate([r[0] for r in res], axis=2)
    status = np.concatenate([r[1] for r in res])
    return FreqSolution(cs, y, status, _merge_stats([r[2] for r in res]), freqs, "ac")


def noise_(cs: CircuitSweep, freqs, **kw) -> FreqSolution:
    """noise!(circ) for every sweep point (src/ac.jl:182-190): output noise power spectral density
    PSD(sol, ref, 2 pi f) (src/ac.jl:265-284), V^2/Hz."""
    freqs = np.asarray(freqs, dtype=float)
    opts = cs._options(kw)
    res = cs._run(lambda plan: plan.noise(freqs, opts))
    y = np.concatenate([r[0] for r in res], axis=2)
    status = np.concatenate([r[1] for r in res])
    return FreqSolution(cs, y, status, _merge_stats([r[2] for r in res]), freqs, "noise")


def freqresp(sol: FreqSolution, ref) -> np.ndarray:
    return sol.array(ref)


def PSD(sol: FreqSolution, ref) -> np.ndarray:
    return sol.array(ref)


def _merge_stats(stats: List[dict]) -> dict:
    out = dict(stats[0])
    for s in stats[1:]:
        for k, v in s.items():
            out[k] = max(out[k], v) if k.endswith("seconds") else out[k] + v
    return out


# ---- forward parameter sensitivities, batched (reference test/sensitivity.jl:31-41, :58-67:
# ODEForwardSensitivityProblem over the circuit's parameters -> one d(solution)/d(parameter) per parameter).
# The reference integrates the sensitivity ODEs next to the circuit (SciMLSensitivity, un-vendored).  Here the
# perturbed circuits of a central-difference stencil are simply MORE SWEEP POINTS of the same batched call: a
# sweep of B points with n parameters becomes one batch of B*(1 + K*n) points (K = 2 or 4 stencil offsets), solved
# by the same kernels, and the stencil is combined on the host.  No new device code, any observable, DC, transient, AC or noise.

_STENCILS = {2: ((-1, 1), (-0.5, 0.5)),
             4: ((-2, -1, 1, 2), (1.0 / 12.0, -2.0 / 3.0, 2.0 / 3.0, -1.0 / 12.0))}


class _ColumnSweep(SweepBase):
    """A sweep given as explicit per-point columns (internal: the expanded batch of a sensitivity stencil)."""

    def __init__(self, cols: Dict[str, np.ndarray]):
        self._cols = {k: np.asarray(v, dtype=float) for k, v in cols.items()}
        self.shape = (len(next(iter(self._cols.values()))),)

    def sweepvars(self):
        return set(self._cols)

    def columns(self):
        return dict(self._cols)

    def __iter__(self):
        names = sorted(self._cols)
        for i in range(self.shape[0]):
            yield tuple((k, None if np.isnan(self._cols[k][i]) else float(self._cols[k][i])) for k in names)


def sensitivity_columns(columns: Dict[str, np.ndarray], wrt: Sequence[str], rel_step: float = 1e-3, order: int = 4):
    """Columns of the expanded batch and the step of every (parameter, point).

    Block 0 (the first B points) is the sweep itself; block 1 + j*K + k is the sweep with parameter wrt[j] moved by
    offsets[k] * h_j, where h_j = rel_step * |p_j| per point (rel_step itself where p_j == 0)."""
    if order not in _STENCILS:
        raise ValueError("order must be 2 or 4")
    offs, _ = _STENCILS[order]
    B = len(next(iter(columns.values())))
    steps = {}
    blocks = {k: [np.asarray(v, dtype=float)] for k, v in columns.items()}
    for name in wrt:
        if name not in columns:
            raise KeyError(f"sensitivity parameter {name!r} is not a swept variable of this CircuitSweep "
                           f"(add it as a one-value Sweep to differentiate at its nominal value)")
        p = np.asarray(columns[name], dtype=float)
        if np.isnan(p).any():
            raise ValueError(f"sensitivity parameter {name!r} is left at its default (None) at some points")
        h = np.where(p != 0.0, rel_step * np.abs(p), rel_step)
        steps[name] = h
        for o in offs:
            for k, v in columns.items():
                blocks[k].append(np.asarray(v, dtype=float) + (o * h if k == name else 0.0))
    assert all(len(b[0]) == B for b in blocks.values())
    return {k: np.concatenate(b) for k, b in blocks.items()}, steps


def combine_stencil(v: np.ndarray, B: int, j: int, h: np.ndarray, order: int = 4) -> np.ndarray:
    """d v / d p_j from values `v[..., B*(1+K*n)]` laid out as `sensitivity_columns` does."""
    offs, coef = _STENCILS[order]
    K = len(offs)
    acc = 0.0
    for k, c in enumerate(coef):
        blk = 1 + j * K + k
        acc = acc + c * v[..., blk * B:(blk + 1) * B]
    return acc / h


class SensitivitySolution:
    """Result of `sensitivities_`: `.solution` is the SweepSolution of the sweep itself, `.array(ref, name)` /
    `.point(idx, ref, name)` the derivative of any unknown or observable with respect to swept parameter `name`
    (`analysis="ac"` / `"noise"`: `.solution` is a FreqSolution and the derivative has a frequency axis)."""

    def __init__(self, cs: "CircuitSweep", big, wrt: List[str], steps: Dict[str, np.ndarray], order: int):
        self.cs, self.big, self.wrt, self.steps, self.order = cs, big, list(wrt), steps, order
        B = len(cs)
        self.freq = isinstance(big, FreqSolution)
        self.freqs = big.freqs if self.freq else None
        self.t = None if self.freq else big.t
        if self.freq:
            self.solution = FreqSolution(cs, big.y[..., :B], big.status[:B], big.stats, big.freqs, big.kind)
        else:
            self.solution = SweepSolution(cs, big.y[..., :B], big.status[:B], big.stats, big.t)
        # a derivative is valid where every point of its stencil converged
        self.status = big.status.reshape(-1, B).max(axis=0)

    def _d(self, ref, name: str) -> np.ndarray:
        j = self.wrt.index(name)
        if self.freq:
            v = self.big.y[self.big.out_index[_resolve(self.big.fc, ref)]]
        else:
            v = _Observable(self.big.fc, ref).value(self.big, slice(None))
        return combine_stencil(v, len(self.cs), j, self.steps[name], self.order)

    def array(self, ref, name: str) -> np.ndarray:
        """d ref / d name over the whole sweep, shaped size(cs) (+ time or frequency axis last)"""
        d = self._d(ref, name)
        if d.ndim == 1:
            return d.reshape(self.cs.shape, order="F")
        return np.moveaxis(d, 0, -1).reshape(self.cs.shape + (d.shape[0],), order="F")

    def point(self, idx, ref, name: str):
        """d ref / d name at one sweep point: a number (DC) or the waveform over `.t` (transient)"""
        i = int(idx) if isinstance(idx, (int, np.integer)) else int(np.ravel_multi_index(tuple(idx), self.cs.shape, order="F"))
        return self._d(ref, name)[..., i]

    @property
    def retcodes(self) -> np.ndarray:
        return np.array([RETCODES.get(int(s), "Failure") for s in self.status], dtype=object).reshape(self.cs.shape, order="F")


class DirectSensitivitySolution:
    """Result of `sensitivities_(..., method="direct")` (DC): `.solution` is the SweepSolution of the sweep itself,
    `.array(ref, name)` / `.point(idx, ref, name)` the derivative of an unknown among the outputs (or of a `<inst>.V`
    observable, a difference of two of them) with respect to swept parameter `name`."""

    def __init__(self, cs: "CircuitSweep", x: np.ndarray, sens: np.ndarray, status: np.ndarray, stats, wrt: List[str]):
        self.cs, self.sens, self.wrt, self.status, self.stats = cs, sens, list(wrt), status, stats
        self.solution = SweepSolution(cs, x, status, stats, None)
        self.t = self.freqs = None
        self.freq = False

    def _d(self, ref, name: str) -> np.ndarray:
        obs = _Observable(self.cs.flat.fc, ref)
        if obs.div is not None or obs.cap is not None:
            raise KeyError(f"{obs.key!r}: the direct method differentiates unknowns and `.V` observables; use method='stencil'")
        j = self.wrt.index(name)
        acc = np.zeros(len(self.cs))
        for sg, u in obs.terms:
            if u not in self.solution.out_index:
                raise KeyError(f"{obs.key!r} needs its unknowns among the outputs")
            acc = acc + sg * self.sens[j, self.solution.out_index[u]]
        return acc

    def array(self, ref, name: str) -> np.ndarray:
        return self._d(ref, name).reshape(self.cs.shape, order="F")

    def point(self, idx, ref, name: str):
        i = int(idx) if isinstance(idx, (int, np.integer)) else int(np.ravel_multi_index(tuple(idx), self.cs.shape, order="F"))
        return self._d(ref, name)[i]

    @property
    def retcodes(self) -> np.ndarray:
        return np.array([RETCODES.get(int(s), "Failure") for s in self.status], dtype=object).reshape(self.cs.shape, order="F")


def _direct_dc_sensitivities(cs: CircuitSweep, wrt: List[str], rel_step: float, **kw) -> DirectSensitivitySolution:
    """cb_sens_dc: one operating-point solve of the sweep, then per swept quantity two chord updates with the stored
    factors of J(x*) and the parameters moved by +-h (h = rel_step |p|): no nonlinear re-solve, B points instead of the
    B (1 + 4 n) of the stencil form.  A swept quantity may feed several parameter columns through netlist expressions: the
    moved columns come from re-flattening the circuit with the moved value (directional derivative)."""
    B = len(cs)
    netlist_circuit = isinstance(cs.circuit, str) or not callable(cs.circuit)
    nl = cs.circuit
    if isinstance(nl, str):
        from .netlist import parse_netlist
        if cs._ctor["lang"] == "spectre":
            from .spectre import parse_spectre
            nl = parse_spectre(nl, include_dirs=cs._ctor["include_dirs"])
        else:
            nl = parse_netlist(nl, include_dirs=cs._ctor["include_dirs"])

    def flat_for(cols):
        # every sweep-dependent value stays a parameter column (a swept value that is the same at all points would
        # otherwise be folded into the circuit as a constant and could not be moved)
        if netlist_circuit:
            return flatten(nl, cols, B=B, outputs=cs._ctor["outputs"], host=cs._ctor["host"], force_columns=True)
        return cs.circuit(cols, B)

    flat0 = flat_for(cs.columns)
    base = flat0.params
    plan = engine.Circuit(flat0.fc, flat0.models).plan(B, devices=cs.devices)
    plan.set_params(np.ascontiguousarray(base) if base.size else None)

    def params_for(cols):
        fl = flat_for(cols)
        if list(fl.fc.param_names) != list(flat0.fc.param_names):
            raise ValueError("moving a swept parameter changed the set of parameter columns")
        return fl.params

    pp, pm, steps = [], [], []
    for name in wrt:
        if name not in cs.columns:
            raise KeyError(f"sensitivity parameter {name!r} is not a swept variable of this CircuitSweep")
        p = np.asarray(cs.columns[name], dtype=float)
        h = np.where(p != 0.0, rel_step * np.abs(p), rel_step)
        steps.append(h)
        for sign, dst in ((1.0, pp), (-1.0, pm)):
            cols = dict(cs.columns)
            cols[name] = p + sign * h
            dst.append(params_for(cols) if base.size else np.zeros((0, B)))
    plan.set_x0(cs.x0)
    x, sens, status, stats = plan.sens_dc(np.stack(pp) if base.size else np.zeros((len(wrt), 0, B)),
                                          np.stack(pm) if base.size else np.zeros((len(wrt), 0, B)), np.stack(steps), cs._options(kw))
    plan.close()
    return DirectSensitivitySolution(cs, x, sens, status, stats, wrt)


def sensitivities_(cs: CircuitSweep, wrt: Optional[Sequence[str]] = None, analysis: str = "dc", tspan=None, saveat=None, freqs=None,
                   rel_step: float = 1e-3, order: int = 4, method: str = "stencil", **kw):
    """Forward sensitivities of every sweep point with respect to the swept parameters `wrt` (default: all of
    them, like the reference's sensitivity problem over the ParamSim's parameters, test/sensitivity.jl:58-67).
    method="stencil": every stencil point is one more sweep point of the same batched solve (any analysis);
    method="direct" (analysis="dc"): the direct method of cb_sens_dc -- one solve, then one pair of triangular solves per
    parameter with the factors of the Newton matrix.
    For `analysis="tran"` use `fixed_step=1` (or tolerances well below the derivative accuracy wanted), so that the
    stencil points share their time grid and step-control noise does not enter the difference."""
    wrt = [w for w in (sorted(cs.columns) if wrt is None else wrt)]
    if method == "direct":
        if analysis != "dc":
            raise ValueError("method='direct' is available for analysis='dc'")
        return _direct_dc_sensitivities(cs, wrt, 1e-4 if rel_step == 1e-3 else rel_step, **kw)
    if method != "stencil":
        raise ValueError("method must be 'stencil' or 'direct'")
    cols, steps = sensitivity_columns(cs.columns, wrt, rel_step, order)
    big_cs = CircuitSweep(cs.circuit, _ColumnSweep(cols), devices=cs.devices, **cs._ctor)
    if cs.x0 is not None:
        reps = len(big_cs) // len(cs)
        big_cs.x0 = cs.x0 if cs.x0.ndim == 1 else np.tile(cs.x0, (1, reps))
    if analysis == "dc":
        big = dc_(big_cs, **kw)
    elif analysis == "tran":
        big = tran_(big_cs, tspan, saveat, **kw)
    elif analysis in ("ac", "noise"):      # derivative of the complex response / of the output PSD over `freqs`
        if freqs is None:
            raise ValueError("analysis='ac' / 'noise' needs freqs")
        big = (ac_ if analysis == "ac" else noise_)(big_cs, freqs, **kw)
    else:
        raise ValueError("analysis must be 'dc', 'tran', 'ac' or 'noise'")
    return SensitivitySolution(cs, big, wrt, steps, order)


dc = dc_
tran = tran_
sensitivities = sensitivities_
ac = ac_
noise = noise_
