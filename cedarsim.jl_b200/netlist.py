"""SPICE-subset netlist reader and flattener: the step *before* the hot path.

Grammar and semantics follow what the reference implements in
SpectreNetlistParser.jl/src/SPICE (lexer.jl:61,82-95 comments / continuations; parse.jl:750-770
device kind by first letter, :842-875 sources, :946-970 M and X cards, :71-96 .tran) and
src/spectre.jl (names lower-cased :17-23; nets `0`/`gnd` grounded :736-749; SI suffixes :383-455;
subckt parameters with implicit m=1 :928-948; dynamic scoping of unresolved names :494-512; `.model`
level -> device :603-607; `m=` multiplicity :1177-1179; .option/.temp -> SimSpec :648-673).

Flattening evaluates every device parameter with numpy over the whole sweep at once: a swept
name (top-level `.param`, `<inst>.<param>`, `<x-path>.<param>`, or a SimSpec field temp/gmin) is an
array with one value per sweep point, so an expression that depends on it becomes one per-instance
parameter column of the flat circuit, and everything else folds to a constant -- the same split
`ParamSim` makes between runtime parameters and constants (src/circuitodesystem.jl:66-97).
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import models
from .expr import ExprError, evaluate, free_vars, parse_expr, parse_number
from .flat import (Col, FlatCircuit, VAModelShape, Wave, W_DC, W_PULSE, W_PWL, W_SIN, shape_of)
from .modelcard import ModelCard, NoBinException, bins_of, find_bin, parse_model_cards


class NetlistError(Exception):
    pass


SIMSPEC_FIELDS = ("temp", "gmin", "scale")


@dataclass
class Card:
    kind: str                    # first letter, lower case
    name: str
    nodes: List[str]
    model: Optional[str] = None
    params: Dict[str, str] = field(default_factory=dict)   # raw expression text
    value: Optional[str] = None
    source: Optional[dict] = None


@dataclass
class Subckt:
    name: str
    ports: List[str]
    params: Dict[str, str] = field(default_factory=dict)
    cards: List[Card] = field(default_factory=list)
    local_params: Dict[str, str] = field(default_factory=dict)
    subckts: Dict[str, "Subckt"] = field(default_factory=dict)


@dataclass
class Netlist:
    title: str = ""
    top: Subckt = field(default_factory=lambda: Subckt("<top>", []))
    cards: Dict[str, ModelCard] = field(default_factory=dict)
    options: Dict[str, str] = field(default_factory=dict)
    tran: Optional[Tuple[float, float]] = None   # (tstep, tstop)
    includes: List[str] = field(default_factory=list)
    va_modules: Dict[str, str] = field(default_factory=dict)   # Verilog-A module name (lower case) -> file, from `.hdl`
    include_dirs: List[str] = field(default_factory=list)


# ---------------------------------------------------------------- lexical level

def logical_lines(text: str, first_is_title: bool = True) -> Tuple[str, List[str]]:
    raw = text.splitlines()
    title = ""
    if first_is_title and raw:
        title = raw[0].strip()
        raw = raw[1:]
    out: List[str] = []
    for line in raw:
        s = line.strip()
        if not s or s.startswith("*"):
            continue
        # trailing comments: `$` or `;` (not inside quotes)
        s = re.split(r"\s[;$]", " " + s, maxsplit=1)[0].strip() if (";" in s or "$" in s) else s
        if not s:
            continue
        if s.startswith("+"):
            if out:
                out[-1] += " " + s[1:].strip()
            continue
        out.append(s)
    return title, out


_TOKEN = re.compile(r"""'[^']*'|\{[^}]*\}|"[^"]*"|\(|\)|,|=|[^\s(),=]+""")


def tokens(line: str) -> List[str]:
    return _TOKEN.findall(line)


def split_params(toks: List[str]) -> Tuple[List[str], Dict[str, str]]:
    """positional tokens, then k=v pairs"""
    pos: List[str] = []
    kv: Dict[str, str] = {}
    i = 0
    while i < len(toks):
        if i + 2 < len(toks) + 0 and i + 1 < len(toks) and toks[i + 1] == "=":
            kv[toks[i].lower()] = toks[i + 2] if i + 2 < len(toks) else ""
            i += 3
        else:
            pos.append(toks[i])
            i += 1
    return pos, kv


# ---------------------------------------------------------------- parsing

_JLPKG = {"asap7pdk/7nm_tt.pm": "asap7", "asap7pdk/7nm_tt.scs": "asap7"}


def _parse_source(toks: List[str]) -> dict:
    """V/I source spec after the two nodes: [value] [DC v] [AC mag [phase]] [PWL(...)|PULSE(...)|SIN(...)]"""
    src = {"dc": None, "ac": None, "tran": None}
    i = 0
    flat = [t for t in toks if t not in (",", "=")]   # `dc=1` and `DC 1` are the same thing (test/basic.jl:373)
    while i < len(flat):
        t = flat[i].lower()
        if t == "dc":
            src["dc"] = flat[i + 1]
            i += 2
        elif t == "ac":
            src["ac"] = flat[i + 1]
            i += 2
            if i < len(flat) and re.match(r"^[-+.\d]", flat[i]):
                i += 1
        elif t in ("pwl", "pulse", "sin"):
            args = []
            i += 1
            if i < len(flat) and flat[i] == "(":
                i += 1
                while i < len(flat) and flat[i] != ")":
                    args.append(flat[i])
                    i += 1
                i += 1
            else:
                while i < len(flat) and flat[i].lower() not in ("dc", "ac"):
                    args.append(flat[i])
                    i += 1
            src["tran"] = (t, args)
        elif t in ("(", ")"):
            i += 1
        else:
            if src["dc"] is None and src["tran"] is None:
                src["dc"] = flat[i]   # a bare value means DC (parse.jl:831-840)
            i += 1
    return src


def parse_netlist(text: str, path: Optional[str] = None, first_is_title: bool = True, _nl: Optional[Netlist] = None,
                  _lib_section: Optional[str] = None, include_dirs: Optional[Sequence[str]] = None) -> Netlist:
    nl = _nl or Netlist()
    if include_dirs:
        nl.include_dirs.extend(include_dirs)
    title, lines = logical_lines(text, first_is_title)
    if _nl is None:
        nl.title = title
    stack: List[Subckt] = [nl.top]
    cond_stack: List[bool] = []     # is the current branch of each open .if live
    cond_taken: List[bool] = []     # has any branch of that .if been live yet (.elseif / .else chains)
    in_lib: Optional[str] = None
    base = os.path.dirname(path) if path else "."
    for line in lines:
        toks = tokens(line)
        if not toks:
            continue
        head = toks[0].lower()
        cur = stack[-1]
        if head == ".lib" and len(toks) == 2:
            in_lib = toks[1].strip("'\"").lower()   # `.LIB name` ... `.ENDL`: a library section definition
            continue
        if head == ".endl":
            in_lib = None
            continue
        # a section is inert where it is defined and is read only through `.LIB "file" name` -- also from the file
        # itself (test/basic.jl:312-336)
        if (in_lib is not None) if _lib_section is None else (in_lib != _lib_section):
            continue
        if head in (".if", ".elseif", ".else", ".endif"):
            cond_text = re.sub(r"^\s*\.\w+\s*", "", line).strip()   # from the raw line: the card tokenizer splits `==`
            if head == ".if":
                cond_stack.append(bool(evaluate(parse_expr(cond_text), _const_env(nl))))
                cond_taken.append(cond_stack[-1])
            elif head == ".else":
                cond_stack[-1] = not cond_taken[-1]
                cond_taken[-1] = True
            elif head == ".elseif":
                cond_stack[-1] = (not cond_taken[-1]) and bool(evaluate(parse_expr(cond_text), _const_env(nl)))
                cond_taken[-1] = cond_taken[-1] or cond_stack[-1]
            else:
                cond_stack.pop()
                cond_taken.pop()
            continue
        if cond_stack and not all(cond_stack):
            continue
        if head.startswith("."):
            if head in (".param", ".parameter", ".parameters"):
                _, kv = split_params(toks[1:])
                (cur.local_params if cur is not nl.top else cur.params).update(kv)
            elif head == ".subckt":
                pos, kv = split_params(toks[1:])
                pos = [p for p in pos if p.lower() != "params:"]
                sub = Subckt(pos[0].lower(), [p.lower() for p in pos[1:]], dict(kv))
                cur.subckts[sub.name] = sub
                stack.append(sub)
            elif head == ".ends":
                if len(stack) > 1:
                    stack.pop()
            elif head == ".model":
                nl.cards.update(parse_model_cards(line))
            elif head == ".hdl":   # Verilog-A include (src/spectre.jl: `.hdl` / `ahdl_include`, test/basic.jl:353-380)
                fname = toks[1].strip("'\"")
                vpath = _resolve(nl, fname, base)
                with open(vpath, "r", errors="replace") as f:
                    for mname in re.findall(r"^\s*module\s+([A-Za-z_][A-Za-z0-9_$]*)", f.read(), re.M):
                        nl.va_modules[mname.lower()] = vpath
            elif head in (".include", ".inc", ".lib"):
                fname = toks[1].strip("'\"")
                section = toks[2].lower() if head == ".lib" and len(toks) > 2 else None
                _include(nl, fname, base, section)
            elif head in (".option", ".options"):
                _, kv = split_params(toks[1:])
                nl.options.update(kv)
            elif head == ".temp":
                nl.options["temp"] = toks[1]
            elif head == ".tran":
                vals = [parse_number(t) for t in toks[1:3]]
                nl.tran = (vals[0], vals[1])
            elif head in (".end", ".global", ".control", ".endc", ".print", ".plot", ".save", ".ic", ".nodeset", ".op",
                          ".dc", ".ac", ".noise", ".probe", ".title"):
                pass
            else:
                pass   # unknown dot-cards are ignored, as the reference warns and continues (spectre.jl:1520-1522)
            continue
        cur.cards.append(_parse_card(toks))
    return nl


def _const_env(nl: Netlist) -> Dict[str, float]:
    env: Dict[str, float] = {}
    for k, v in nl.top.params.items():
        try:
            env[k] = evaluate(parse_expr(v), env)
        except ExprError:
            pass
    return env


def _resolve(nl: Netlist, fname: str, base: str) -> str:
    """A file named in the deck: absolute, next to the including file, or in one of the include directories."""
    if os.path.isabs(fname):
        return fname
    for d in [base] + list(nl.include_dirs):
        cand = os.path.join(d, fname)
        if os.path.exists(cand):
            return cand
    raise NetlistError(f"cannot find {fname!r} (searched {[base] + list(nl.include_dirs)})")


def _include(nl: Netlist, fname: str, base: str, section: Optional[str]):
    nl.includes.append(fname)
    low = fname.lower()
    if low.startswith("jlpkg://"):
        key = low[len("jlpkg://"):]
        if _JLPKG.get(key) == "asap7":
            nl.cards.update(models.asap7_cards())
            return
        raise NetlistError(f"package include {fname!r} is not available (BSIM4 / GF180 / sky130 PDKs are not in the "
                           "reference tree, SURVEY.md fact 5)")
    path = _resolve(nl, fname, base)
    with open(path, "r", errors="replace") as f:
        text = f.read()
    if re.search(r"^\s*simulator\s+lang\s*=\s*spectre", text, re.M | re.I) or path.endswith(".scs"):
        nl.cards.update(parse_model_cards(text))
        return
    parse_netlist(text, path, first_is_title=False, _nl=nl, _lib_section=section)


def _parse_card(toks: List[str]) -> Card:
    name = toks[0].lower()
    kind = name[0]
    rest = toks[1:]
    if kind in "rcl":
        pos, kv = split_params(rest)
        nodes = [p.lower() for p in pos[:2]]
        value = pos[2] if len(pos) > 2 else None
        model = None
        if value is not None and re.match(r"^[A-Za-z_]", value) and not value.startswith(("'", "{")):
            model, value = value.lower(), (pos[3] if len(pos) > 3 else None)
        return Card(kind, name, nodes, model, kv, value)
    if kind in "vi":
        nodes = [rest[0].lower(), rest[1].lower()]
        return Card(kind, name, nodes, source=_parse_source(rest[2:]))
    if kind == "b" or (kind in "eg" and any(t.lower() in ("vol", "cur", "value") for t in rest[2:4])):
        # behavioural source: `B1 p n v=<expr>` / `i=<expr>`, `E1 p n vol=<expr>`, `G1 p n cur=<expr>` (test/basic.jl:207-235)
        eq = rest.index("=")
        key = rest[eq - 1].lower()
        expr = " ".join(rest[eq + 1:]).strip()
        if len(expr) > 1 and expr[0] in "'{" and expr[-1] in "'}":
            expr = expr[1:-1]
        return Card("b", name, [rest[0].lower(), rest[1].lower()], None, {"v" if key in ("v", "vol", "value") else "i": expr})
    if kind in "eg":
        pos, kv = split_params([t for t in rest if t not in ("(", ")", ",")])
        nodes = [p.lower() for p in pos[:4]]
        value = pos[4] if len(pos) > 4 else kv.get("gain")
        return Card(kind, name, nodes, None, kv, value)
    if kind == "m":
        pos, kv = split_params(rest)
        return Card(kind, name, [p.lower() for p in pos[:4]], pos[4].lower(), kv)
    if kind == "x":
        pos, kv = split_params(rest)
        return Card(kind, name, [p.lower() for p in pos[:-1]], pos[-1].lower(), kv)  # last bare id = subckt
    raise NetlistError(f"unsupported device card {toks[0]!r}")


def parse_file(path: str) -> Netlist:
    with open(path, "r", errors="replace") as f:
        return parse_netlist(f.read(), path)


# ---------------------------------------------------------------- flattening

_VA_FUNCS = {"abs", "exp", "ln", "log", "sqrt", "pow", "min", "max", "sin", "cos", "tan", "atan", "tanh", "sinh", "cosh", "limexp"}

Num = Union[float, np.ndarray]


class _Scope:
    """Parameter scope with lazy, memoised evaluation and dynamic scoping to the parent."""

    def __init__(self, exprs: Dict[str, str], parent: Optional["_Scope"], overrides: Dict[str, Num]):
        self.exprs = {k.lower(): v for k, v in exprs.items()}
        self.parent = parent
        self.values: Dict[str, Num] = dict(overrides)
        self._busy: set = set()

    def lookup(self, name: str) -> Num:
        name = name.lower()
        if name in self.values:
            return self.values[name]
        if name in self.exprs:
            if name in self._busy:
                # `.subckt inner a b foo=foo+2000`: inside its own default a name means the enclosing scope's value
                # (dynamic scoping, src/spectre.jl:494-512; test/params.jl:58-99)
                if self.parent is not None:
                    return self.parent.lookup(name)
                raise NetlistError(f"circular parameter definition for {name!r}")
            self._busy.add(name)
            v = self.eval(self.exprs[name])
            self._busy.discard(name)
            self.values[name] = v
            return v
        if self.parent is not None:
            return self.parent.lookup(name)
        raise ExprError(f"undefined parameter {name!r}")

    def eval(self, text: str) -> Num:
        e = parse_expr(text)
        env = {v: self.lookup(v) for v in free_vars(e)}
        return evaluate(e, env)


@dataclass
class Flattened:
    fc: FlatCircuit
    params: np.ndarray            # [P][B]
    models: list                  # CompiledModel objects used
    options: Dict[str, Num]       # SimSpec values (temp, gmin, ...) possibly per point
    tran: Optional[Tuple[float, float]]


_TIME_NET = "time__"   # hidden net carrying simulation time for `$time` in behavioural sources


class _Flattener:
    def __init__(self, nl: Netlist, sweep: Dict[str, np.ndarray], B: int, host: bool, outputs: Optional[Sequence[str]] = None,
                 force_columns: bool = False):
        self.nl, self.sweep, self.B, self.host = nl, {k.lower(): v for k, v in sweep.items()}, B, host
        self.force_columns = force_columns   # keep every sweep-dependent value a parameter column, even a uniform one
        # branch currents of Verilog-A instances asked for as observables: `<inst>.i(a,b)`
        self.want_branches: Dict[str, List[Tuple[str, str]]] = {}
        for o in outputs or ():
            m = re.match(r"^(.*)\.i\(([^,()]+),([^,()]+)\)$", str(o).lower().replace(" ", "")) if isinstance(o, str) else None
            if m:
                self.want_branches.setdefault(m.group(1), []).append((m.group(2), m.group(3)))
        self.fc = FlatCircuit()
        self.columns: List[np.ndarray] = []
        self.models: list = []
        self.used_sweep: set = set()

    def col(self, name: str, arr: np.ndarray) -> Col:
        c = self.fc.param(name)
        if c.index == len(self.columns):
            self.columns.append(np.broadcast_to(np.asarray(arr, dtype=float), (self.B,)).copy())
        return c

    def value(self, tag: str, v: Num):
        """constant or per-point column"""
        if isinstance(v, np.ndarray) and v.ndim > 0:
            if np.all(v == v.flat[0]) and not self.force_columns:
                return float(v.flat[0])
            return self.col(tag, v)
        return float(v)

    def overrides(self, prefix: str, names: Sequence[str], scope_defaults: Optional[_Scope] = None) -> Dict[str, Num]:
        """swept values addressed as `<prefix><name>`; None entries (NaN) keep the default"""
        out: Dict[str, Num] = {}
        for n in names:
            key = (prefix + n).lower()
            if key in self.sweep:
                self.used_sweep.add(key)
                out[n.lower()] = self.sweep[key]
        return out

    def run(self) -> Flattened:
        nl = self.nl
        top_over = self.overrides("", list(nl.top.params))
        top = _Scope(nl.top.params, None, top_over)
        self._fill_defaults(top, nl.top.params, top_over)
        # `temper` in expressions = the simulation temperature in Celsius: a swept `temp`, else `.option temp=` / `.temp`,
        # else 27 (src/spectre_env.jl temper, src/simulate_ir.jl:12-20; test/basic.jl:469-517)
        if "temp" in self.sweep:
            top.values["temper"] = self.sweep["temp"]
        elif "temp" in {k.lower() for k in nl.options}:
            top.values["temper"] = float(parse_number(str({k.lower(): v for k, v in nl.options.items()}["temp"])))
        else:
            top.values["temper"] = 27.0
        self.instantiate(nl.top, top, prefix="", portmap={}, mult_ctx=1.0)
        opts: Dict[str, Num] = {}
        for k, v in nl.options.items():
            try:
                opts[k.lower()] = top.eval(v)
            except ExprError:
                continue
        for f_ in SIMSPEC_FIELDS:
            if f_ in self.sweep:
                self.used_sweep.add(f_)
                opts[f_] = self.value(f_, self.sweep[f_])
        unused = set(self.sweep) - self.used_sweep
        if unused:
            raise NetlistError(f"sweep variable(s) {sorted(unused)} do not name any parameter of the circuit")
        P = np.stack(self.columns, axis=0) if self.columns else np.zeros((0, self.B))
        return Flattened(self.fc, np.ascontiguousarray(P), self.models, opts, nl.tran)

    @staticmethod
    def _fill_defaults(scope: _Scope, exprs: Dict[str, str], over: Dict[str, Num]):
        # `nothing` in a SerialSweep means "keep the default" (src/sweeps.jl:18-21): NaN -> default
        for k, v in list(over.items()):
            if isinstance(v, np.ndarray) and np.isnan(v).any() and k in {e.lower() for e in exprs}:
                del scope.values[k]
                default = scope.lookup(k)
                scope.values[k] = np.where(np.isnan(v), default, v)

    def instantiate(self, sub: Subckt, scope: _Scope, prefix: str, portmap: Dict[str, str], mult_ctx: float):
        # a subcircuit's ports are aliases of the parent's nets: sys.x1.node_pos == sys.node_vcc (net_alias,
        # src/spectre.jl:911-913,951; test/alias.jl:21-33)
        for port, parent in portmap.items():
            self.fc.aliases[(prefix + port).lower()] = parent.lower()

        def net(n: str) -> str:
            if n in ("0", "gnd", "gnd!"):
                return "0"
            if n in portmap:
                return portmap[n]
            return prefix + n

        for card in sub.cards:
            name = prefix + card.name
            inst_over = self.overrides(name + ".", list(card.params) + ["r", "c", "l", "dc", "gain", "w", "nfin", "m"])

            def par(key: str, default=None):
                if key in inst_over:
                    return inst_over[key]
                if key in card.params:
                    return scope.eval(card.params[key])
                return default

            mult = par("m", 1.0)
            if isinstance(mult, np.ndarray):
                raise NetlistError("multiplicity m cannot be swept")
            own_m = float(mult)
            mult = own_m * mult_ctx   # ParallelInstances nest multiplicatively (src/simulate_ir.jl:43-49)
            nodes = [net(n) for n in card.nodes]
            k = card.kind
            if k == "r":
                v = inst_over.get("r", scope.eval(card.value) if card.value is not None else par("r"))
                if v is None and card.model is not None and card.model not in self.nl.cards:
                    v = scope.eval(card.model)   # `R2 vcc 0 res`: a bare identifier that is a parameter, not a model (test/basic.jl:725-737)
                if v is None:   # model card / geometry: r = rsh*(l-short)/(w-narrow) (src/simpledevices.jl:62-70)
                    mp = {kk.lower(): vv for kk, vv in (self.nl.cards[card.model].params.items() if card.model in self.nl.cards else ())}
                    g = lambda key, d: par(key, mp.get(key, d))
                    v = mp.get("r") if "r" in mp else g("rsh", 50.0) * (g("l", 1e-6) - g("short", 0.0)) / (g("w", 1e-6) - g("narrow", 0.0))
                self.fc.resistor(name, nodes[0], nodes[1], self.value(name + ".r", v), m=mult)
            elif k == "c":
                v = inst_over.get("c", scope.eval(card.value) if card.value is not None else par("c", 1.0))
                self.fc.capacitor(name, nodes[0], nodes[1], self.value(name + ".c", v), m=mult)
            elif k == "l":
                v = inst_over.get("l", scope.eval(card.value) if card.value is not None else par("l"))
                self.fc.inductor(name, nodes[0], nodes[1], self.value(name + ".l", v), m=mult)
            elif k in "vi":
                w = self._wave(name, card.source, scope, inst_over)
                (self.fc.vsource if k == "v" else self.fc.isource)(name, nodes[0], nodes[1], w, m=mult)
            elif k in "eg":
                g = inst_over.get("gain", scope.eval(card.value) if card.value is not None else 1.0)
                (self.fc.vcvs if k == "e" else self.fc.vccs)(name, nodes[0], nodes[1], nodes[2], nodes[3],
                                                              self.value(name + ".gain", g), m=mult)
            elif k == "b":
                self._behavioural(name, card, nodes, scope, mult, net)
            elif k == "m":
                self._mosfet(name, card, nodes, scope, inst_over, mult)
            elif k == "x":
                child = self._find_subckt(sub, card.model)
                if child is None and card.model in self.nl.va_modules:
                    self._va_device(name, card, nodes, scope, mult)
                    continue
                if child is None:
                    raise NetlistError(f"unknown subcircuit {card.model!r} for {name}")
                # instance parameters may refer to one another (`x1 ... w=4 nrd='w/2'`, test/basic.jl:519-536) and to
                # the caller's scope; a name inside its own value means the caller's (`foo=foo+1`)
                inst_scope = _Scope({kk: vv for kk, vv in card.params.items() if kk != "m"}, scope, {})
                given = {kk: inst_scope.lookup(kk) for kk in card.params if kk != "m"}
                given.update(self.overrides(name + ".", list(child.params) + list(child.local_params)))
                exprs = dict(child.params)
                exprs.update(child.local_params)
                cs = _Scope(exprs, scope, given)
                self._fill_defaults(cs, exprs, given)
                if len(nodes) != len(child.ports):
                    raise NetlistError(f"{name}: {len(nodes)} nodes for subcircuit {child.name} with {len(child.ports)} ports")
                # the subckt's own `m` (instance value, else its declared default, else 1) scales everything inside
                sub_m = own_m if "m" in card.params or "m" in inst_over else float(cs.lookup("m")) if "m" in exprs else 1.0
                self.instantiate(child, cs, name + ".", dict(zip(child.ports, nodes)), mult_ctx * sub_m)
            else:
                raise NetlistError(f"unsupported device {name}")

    def _find_subckt(self, sub: Subckt, name: str) -> Optional[Subckt]:
        if name in sub.subckts:
            return sub.subckts[name]
        return self.nl.top.subckts.get(name)

    def _wave(self, name: str, src: dict, scope: _Scope, over: Dict[str, Num]) -> Wave:
        dc = over.get("dc", scope.eval(src["dc"]) if src["dc"] is not None else None)
        dcv = None if dc is None else self.value(name + ".dc", dc)
        w = self._wave_tran(name, src, scope, dcv)
        if src.get("ac") is not None:
            ac = scope.eval(src["ac"])
            if isinstance(ac, np.ndarray):
                raise NetlistError("the AC magnitude of a source cannot be swept")
            w.ac = abs(float(ac))
        return w

    def _wave_tran(self, name: str, src: dict, scope: _Scope, dcv) -> Wave:
        if src["tran"] is None:
            return Wave(W_DC, dc=0.0 if dcv is None else dcv)
        kind, args = src["tran"]
        vals = [scope.eval(a) for a in args]
        if kind == "pwl":
            if len(vals) % 2:
                raise NetlistError("PWL must have an equal number of x and y values")
            ts = [float(v) for v in vals[0::2]]
            ys = [self.value(f"{name}.pwl{i}", v) for i, v in enumerate(vals[1::2])]
            return Wave(W_PWL, dc=dcv, t=ts, y=ys)
        if kind == "pulse":
            v = [self.value(f"{name}.pulse{i}", x) for i, x in enumerate(vals)]
            return Wave(W_PULSE, dc=dcv, v=v)
        v = [self.value(f"{name}.sin{i}", x) for i, x in enumerate(vals)]
        return Wave(W_SIN, dc=dcv, v=v)

    def _behavioural(self, name: str, card: Card, nodes: List[str], scope: _Scope, mult: float, net):
        """Behavioural source as a generated Verilog-A module: `V(p,n) <+ expr` (a voltage branch with its own current
        unknown) or `I(p,n) <+ expr`; every net the expression probes with V(a) / V(a,b) becomes a port, every
        netlist parameter it names a module parameter (so it can be a sweep column).  The reference lowers B / E vol= /
        G cur= sources to closures over the same probes (src/spectre.jl:1020-1071)."""
        import hashlib
        kind, text = next(iter(card.params.items()))
        ports: List[str] = []

        def port(n: str) -> str:
            n = n.strip().lower()
            if n not in ports:
                ports.append(n)
            return f"c{ports.index(n)}"

        def probe(m):
            args = [a for a in m.group(1).split(",")]
            return "V(" + ", ".join(port(a) for a in args) + ")"

        # `$time` / `$abstime`: the device code takes no time argument, so simulation time enters as the voltage of a
        # hidden net driven by V(t) = t (one shared ramp source per circuit); the module probes it like any other net
        text = re.sub(r"\$(?:abs)?time\b", f"V({_TIME_NET})", text, flags=re.I)
        body = re.sub(r"\b[vV]\s*\(([^()]*)\)", probe, text)
        if re.search(r"\b[iI]\s*\(", body) or "$" in body:
            raise NetlistError(f"{name}: only V() probes are supported in behavioural sources ({text!r})")
        body = body.replace("**", "^")
        # numbers with SPICE magnitudes -> plain literals; identifiers that are netlist parameters -> module parameters
        from .expr import parse_number, _CONSTS
        pars: List[str] = []

        def atom(m):
            tok = m.group(0)
            if re.match(r"^(\d|\.\d)", tok):
                return repr(parse_number(tok))
            low = tok.lower()
            if low in ("v",) or re.match(r"^c\d+$", low) or low in _VA_FUNCS:
                return low if low in _VA_FUNCS else tok
            if low in _CONSTS:      # pi, e, M_PI, ...: constants of the expression language unless a parameter shadows them
                try:
                    scope.lookup(low)
                except ExprError:
                    return repr(float(_CONSTS[low]))
            if low not in pars:
                pars.append(low)
            return "P_" + low
        body = re.sub(r"(\d+\.?\d*(?:[eE][+-]?\d+)?[A-Za-z]*|\.\d+(?:[eE][+-]?\d+)?[A-Za-z]*|[A-Za-z_][A-Za-z0-9_]*)", atom, body)
        body = body.replace("^", "**")
        allp = ["p", "n"] + [f"c{k}" for k in range(len(ports))]
        tag = hashlib.sha1((kind + body + repr(pars)).encode()).hexdigest()[:10]
        mname = f"bsrc_{tag}"
        va = ["`include \"disciplines.vams\"", f"module {mname}({', '.join(allp)});", f"    inout {', '.join(allp)};",
              f"    electrical {', '.join(allp)};"]
        va += [f"    parameter real P_{q} = 0.0;" for q in pars]
        va += ["    analog begin", f"        {'V' if kind == 'v' else 'I'}(p, n) <+ {body};", "    end", "endmodule", ""]
        os.makedirs(models.GEN_DIR, exist_ok=True)
        path = os.path.join(models.GEN_DIR, mname + ".va")
        if not os.path.exists(path):
            with open(f"{path}.{os.getpid()}.tmp", "w") as f:
                f.write("\n".join(va))
            os.replace(f"{path}.{os.getpid()}.tmp", path)
        cm = models.compiled_model(mname, path, module=mname)
        if cm not in self.models:
            self.models.append(cm)
        if self.host:
            from .va.build import build_host
            shape = build_host(cm).shape()
        else:
            shape = shape_of(cm)
        vals = {f"P_{q}": self.value(f"{name}.{q}", scope.lookup(q)) for q in pars}
        self.fc.va_instance(name, self.fc.va_model(shape),
                            nodes[:2] + [self._time_net() if q == _TIME_NET else net(q) for q in ports], vals, m=mult)

    def _time_net(self) -> str:
        """The net whose voltage is the simulation time: a PWL source 0 -> T over [0, T] with T = 2^20 s (slope exactly
        1; 0 V at the DC operating point, like `$time` there).  Created on first use; shows up as unknown `$time`."""
        if not getattr(self, "_has_time_net", False):
            T = float(2 ** 20)
            self.fc.vsource(_TIME_NET, _TIME_NET, "0", Wave(W_PWL, dc=0.0, t=[0.0, T], y=[0.0, T]))
            self._has_time_net = True
        return _TIME_NET

    def _va_device(self, name: str, card: Card, nodes: List[str], scope: _Scope, mult: float):
        """Instance of a Verilog-A module brought in by `.hdl`: all module parameters stay run-time parameters (instance
        values, swept columns), names matched case-insensitively as SPICE does (src/spectre.jl:17-23)."""
        import hashlib
        path = self.nl.va_modules[card.model]
        with open(path, "rb") as f:
            tag = hashlib.sha1(f.read()).hexdigest()[:10]
        probes = tuple(self.want_branches.get(name, ()))   # sys.<inst>.var"I(p, n)" observables become unknowns of the device
        if probes:
            tag += "_" + hashlib.sha1(repr(probes).encode()).hexdigest()[:6]
        cm = models.compiled_model(f"{card.model}_{tag}", path, module=self._va_module_name(path, card.model),
                                   probe_branches=probes)
        if cm not in self.models:
            self.models.append(cm)
        if self.host:
            from .va.build import build_host
            shape = build_host(cm).shape()
        else:
            shape = shape_of(cm)
        # a module that reads $abstime has a hidden last port for the circuit's time net (va/compiler.py _lower_abstime)
        timed = cm.nports > 0 and cm.terminals[cm.nports - 1] == _TIME_NET
        if len(nodes) > cm.nports - (1 if timed else 0):
            raise NetlistError(f"{name}: {len(nodes)} nodes for Verilog-A module {cm.module} with {cm.nports - (1 if timed else 0)} ports")
        if timed:
            if len(nodes) < cm.nports - 1:
                raise NetlistError(f"{name}: a module that reads $abstime needs all of its {cm.nports - 1} ports connected")
            nodes = list(nodes) + [self._time_net()]
        lut = {p.lower(): p for p in cm.params}
        over = self.overrides(name + ".", list(lut))
        vals: Dict[str, Num] = {}
        for k in set(card.params) | set(over):
            if k == "m":
                continue
            if k not in lut:
                raise NetlistError(f"{name}: Verilog-A module {cm.module} has no parameter {k!r}")
            vals[lut[k]] = self.value(f"{name}.{k}", over[k] if k in over else scope.eval(card.params[k]))
        self.fc.va_instance(name, self.fc.va_model(shape), nodes, vals, m=mult)

    @staticmethod
    def _va_module_name(path: str, lower: str) -> str:
        with open(path, "r", errors="replace") as f:
            for mname in re.findall(r"^\s*module\s+([A-Za-z_][A-Za-z0-9_$]*)", f.read(), re.M):
                if mname.lower() == lower:
                    return mname
        raise NetlistError(f"module {lower!r} not found in {path}")

    def _binned(self, name: str, card: Card, scope: _Scope, over: Dict[str, Num]) -> ModelCard:
        """`<model>.<N>` bins: the bin whose [lmin,lmax) x [wmin,wmax) window holds scale*l, scale*w
        (src/spectre.jl:1160-1170; `scale` from `.option scale=`, :1217).  One bin per sweep: the card is
        compiled into the device code, so all points of a sweep must fall into the same bin."""
        geo = {}
        for k in ("l", "w"):
            if k in over:
                geo[k] = np.atleast_1d(np.asarray(over[k], dtype=float))
            elif k in card.params:
                geo[k] = np.atleast_1d(np.asarray(scope.eval(card.params[k]), dtype=float))
            else:
                raise NetlistError(f"{name}: binned model {card.model!r} needs both l= and w= on the instance")
        sc = self.nl.options.get("scale", 1.0)
        sc = float(scope.eval(sc)) if isinstance(sc, str) else float(sc)
        l, w = np.broadcast_arrays(geo["l"], geo["w"])
        try:
            picks = {find_bin(self.nl.cards, card.model, float(a), float(b), sc).name
                     for a, b in set(zip(l.tolist(), w.tolist()))}
        except NoBinException as e:
            raise NetlistError(f"{name}: {e}") from None
        if len(picks) != 1:
            raise NetlistError(f"{name}: sweep crosses bins {sorted(picks)} of {card.model!r}; the model card is "
                               "compiled into the device code -- split the sweep at the bin boundary")
        return self.nl.cards[picks.pop()]

    _BIN_KEYS = ("LMIN", "LMAX", "WMIN", "WMAX")

    def _mosfet(self, name: str, card: Card, nodes: List[str], scope: _Scope, over: Dict[str, Num], mult: float):
        mc = self.nl.cards.get(card.model)
        binned = mc is None and bool(bins_of(self.nl.cards, card.model))
        if binned:
            mc = self._binned(name, card, scope, over)
        if mc is None:
            raise NetlistError(f"unknown model {card.model!r} for {name}")
        if mc.exprs:   # card values written in terms of `.param`s: resolved in the instance's scope
            mc = ModelCard(mc.name, mc.master, dict(mc.params))
            for k, e in self.nl.cards[mc.name].exprs.items():
                v = scope.eval(e)
                if isinstance(v, np.ndarray) and v.ndim > 0:
                    if not np.all(v == v.flat[0]):
                        raise NetlistError(f"model {mc.name!r}: card parameter {k} varies over the sweep; cards are "
                                           "compiled into the device code, sweep an instance parameter instead")
                    v = v.flat[0]
                mc.params[k] = float(v)
        if not mc.master.startswith("bsimcmg"):
            raise NetlistError(f"model {card.model!r}: device family {mc.master!r} is not available "
                               "(only BSIM-CMG 107 is vendored in the reference tree)")
        if binned:
            # BSIM-CMG has no LMIN/LMAX/WMIN/WMAX/W parameters (they are BSIM4's): for this family the window
            # and the instance's w= only select the bin
            mc = ModelCard(mc.name, mc.master, {k: v for k, v in mc.params.items() if k not in self._BIN_KEYS})
        inst: Dict[str, Num] = {}
        for k in set(card.params) | set(over):
            if k == "m" or (binned and k == "w"):
                if k == "w" and k in over:
                    self.used_sweep.add((name + ".w").lower())
                continue
            inst[k.upper()] = over[k] if k in over else scope.eval(card.params[k])
        runtime = tuple(sorted(set(inst) | {"L", "NFIN"}))
        cm = models.specialized_model(mc, runtime)
        if cm not in self.models:
            self.models.append(cm)
        if self.host:
            from .va.build import build_host
            shape = build_host(cm).shape()
        else:
            shape = shape_of(cm)
        m = self.fc.va_model(shape)
        vals = {k: self.value(f"{name}.{k.lower()}", v) for k, v in inst.items()}
        for k in runtime:   # runtime parameters must always be supplied: card value or model default
            if k not in vals:
                vals[k] = float(mc.params.get(k, models.param_default(cm, k)))
        self.fc.va_instance(name, m, nodes, vals, m=mult)


def flatten(nl: Netlist, sweep: Optional[Dict[str, np.ndarray]] = None, B: int = 1, host: bool = False,
            outputs: Optional[Sequence[str]] = None, force_columns: bool = False) -> Flattened:
    """Flatten `nl` for a sweep given as {name: array of B values}.  force_columns: every value that depends on a swept
    name stays a per-point parameter column even when it is the same at all points (the direct sensitivity method moves
    those columns; by default a uniform value is folded into the circuit as a constant)."""
    sweep = sweep or {}
    if sweep:
        B = len(next(iter(sweep.values())))
    fl = _Flattener(nl, sweep, B, host, outputs, force_columns).run()
    fl.fc.finalize()
    if outputs is not None:
        fl.fc.set_outputs(list(outputs))
    else:
        fl.fc.set_outputs(list(range(fl.fc.n_unknowns)))
    return fl
