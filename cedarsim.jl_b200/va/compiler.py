"""Verilog-A -> C / CUDA C generator with staging and forward-mode AD.

Semantics follow the reference's VA lowering (src/vasim.jl:128-180 contributions, :337-454
probes / ddx / $param_given / noise / user functions, :459-492 assignment + vaconvert,
:503-569 analog functions, :603-626 case; src/va_env.jl:35-123 math + $temperature) but the
architecture is different: instead of tracing Dual numbers through a DAE compiler, one
flow-sensitive abstract interpretation of the analog block emits two straight C functions

  <model>_setup : everything that depends only on parameters / temperature ("static" stage).
                  Runs once per (instance, device); values the bias-dependent code needs are
                  written to numbered cache slots.
  <model>_eval  : everything downstream of a V() probe ("dynamic" stage), with explicit
                  partial derivatives carried only for the terminal voltages a value can
                  actually depend on.  Produces per-terminal currents I, charges Q and the
                  structurally non-zero entries of dI/dV (G) and dQ/dV (C).

The same text compiles as C (CPU oracle, via va/build.py) and CUDA C (engine, via NVRTC)
through the small macro vocabulary PAR / GIVEN / TEMP_K / GMIN_V / CACHE_ST / CACHE_LD / VT
/ OUT_I / OUT_Q / OUT_J.

MNA formulation: `I(a,b) <+ e` adds e to KCL(a) and -e to KCL(b); `ddt(q)` terms go to the
charge vector.  Branch currents of VA devices are observables, not unknowns.
"""
from __future__ import annotations

import hashlib
import math
import os
import re
from dataclasses import dataclass, field
from typing import Dict, FrozenSet, List, Optional, Sequence, Set, Tuple

from .parser import Function, Module, parse
from .preproc import Preprocessor


# cache streaming (see _Compiler._stage_cache): rows per ring chunk and residency window, in stream positions
GEN_VERSION = 7   # bump when the emitted code changes: models cached under _gen/ are regenerated
# experiment knobs (a non-default value needs its own CB_GEN_DIR: cached models are looked up by name)
CACHE_CHUNK_ROWS = int(os.environ.get("CB_CACHE_ROWS", "4"))
CACHE_WINDOW = int(os.environ.get("CB_CACHE_WINDOW", "8"))
UNIFORM_SLOTS = os.environ.get("CB_UNIFORM_SLOTS", "1") != "0"
FOLD_CONST_DIV = os.environ.get("CB_FOLD_CONST_DIV", "1") != "0"


class VACompileError(Exception):
    pass


NOISE_FUNCS = {"white_noise", "flicker_noise", "noise_table"}
MATH1 = {"exp", "ln", "log", "sqrt", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh",
         "asinh", "acosh", "atanh", "abs", "floor", "ceil", "limexp"}
MATH2 = {"pow", "min", "max", "atan2", "hypot"}

P_K, P_Q = 1.3806503e-23, 1.602176462e-19

_PYF = {"ln": math.log, "log": math.log10, "abs": abs, "min": min, "max": max, "exp": math.exp, "sqrt": math.sqrt,
        "pow": math.pow, "sin": math.sin, "cos": math.cos, "tan": math.tan, "asin": math.asin, "acos": math.acos,
        "atan": math.atan, "atan2": math.atan2, "sinh": math.sinh, "cosh": math.cosh, "tanh": math.tanh,
        "asinh": math.asinh, "acosh": math.acosh, "atanh": math.atanh, "hypot": math.hypot, "floor": math.floor,
        "ceil": math.ceil, "limexp": lambda x: math.exp(x) if x < 80.0 else math.exp(80.0) * (1.0 + x - 80.0)}


class _NotConst(Exception):
    pass


def _vaconvert_int(x: float) -> int:
    """real -> integer conversion, ties away from zero (src/va_env.jl:107)"""
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


class _ConstEval:
    """Compile-time evaluation of analog functions whose arguments are all constants: lets the
    generator fold everything that depends only on a model card (circuit-specialised code)."""

    def __init__(self, functions):
        self.functions = functions

    def call(self, fname: str, args: List[float]):
        f = self.functions[fname]
        if any(kind != "input" for _, kind in f.args):
            raise _NotConst()
        env = {v: (0 if t == "integer" else 0.0) for v, t in f.var_types.items()}
        types = dict(f.var_types)
        for (an, _), v in zip(f.args, args):
            env[an] = _vaconvert_int(v) if types.get(an) == "integer" and isinstance(v, float) else v
        self.stmt(f.body, env, types)
        return env[fname]

    def stmt(self, st, env, types):
        k = st[0]
        if k in ("nop", "task"):
            return
        if k == "block":
            if st[3]:
                raise _NotConst()
            for s_ in st[2]:
                self.stmt(s_, env, types)
        elif k == "assign":
            v = self.expr(st[2], env)
            if st[1] not in env:
                raise _NotConst()
            env[st[1]] = _vaconvert_int(v) if types.get(st[1]) == "integer" and isinstance(v, float) else \
                (float(v) if types.get(st[1]) != "integer" else v)
        elif k == "if":
            if self.expr(st[1], env):
                self.stmt(st[2], env, types)
            elif st[3] is not None:
                self.stmt(st[3], env, types)
        else:
            raise _NotConst()

    def expr(self, e, env):
        k = e[0]
        if k == "num":
            return e[1]
        if k == "var":
            if e[1] not in env:
                raise _NotConst()
            return env[e[1]]
        if k == "un":
            a = self.expr(e[2], env)
            if e[1] == "-":
                return -a
            if e[1] == "!":
                return int(not a)
            raise _NotConst()
        if k == "bin":
            a, b = self.expr(e[2], env), self.expr(e[3], env)
            both_int = isinstance(a, int) and isinstance(b, int)
            try:
                return _Compiler._fold(e[1], a, b, both_int)
            except (ValueError, ZeroDivisionError, OverflowError):
                raise _NotConst()
        if k == "cond":
            return self.expr(e[2], env) if self.expr(e[1], env) else self.expr(e[3], env)
        if k == "call":
            args = [self.expr(a, env) for a in e[2]]
            if e[1] in self.functions:
                return self.call(e[1], args)
            fn = _PYF.get(e[1])
            if fn is None:
                raise _NotConst()
            try:
                r = fn(*[float(a) for a in args])
            except (ValueError, OverflowError, ZeroDivisionError):
                raise _NotConst()
            return float(r)
        raise _NotConst()


def _lit(x) -> str:
    if isinstance(x, bool):
        return "1" if x else "0"
    if isinstance(x, int):
        return str(x)
    if math.isinf(x):
        return "INFINITY" if x > 0 else "(-INFINITY)"
    if math.isnan(x):
        return "NAN"
    r = repr(float(x))
    return f"({r})" if x < 0 else r


# ---------------------------------------------------------------------------------------
# liveness: drop assignments whose value can never reach a contribution
# ---------------------------------------------------------------------------------------

_NOISE_LIVE = [False]   # the noise variant keeps the arguments of white_noise()/flicker_noise() alive


def _expr_vars(e, out: Set[str], funcs):
    k = e[0]
    if k == "var":
        out.add(e[1])
    elif k == "bin":
        _expr_vars(e[2], out, funcs)
        _expr_vars(e[3], out, funcs)
    elif k == "un":
        _expr_vars(e[2], out, funcs)
    elif k == "cond":
        for s in e[1:]:
            _expr_vars(s, out, funcs)
    elif k == "call":
        if e[1] in NOISE_FUNCS and not _NOISE_LIVE[0]:
            return
        for a in e[2]:
            _expr_vars(a, out, funcs)


def _call_outs(e, funcs, out: Set[str]):
    """variables written through output/inout arguments of user functions inside e"""
    k = e[0]
    if k == "call":
        fn = funcs.get(e[1])
        if fn is not None:
            for (an, kind), a in zip(fn.args, e[2]):
                if kind != "input" and a[0] == "var":
                    out.add(a[1])
        for a in e[2]:
            _call_outs(a, funcs, out)
    elif k == "bin":
        _call_outs(e[2], funcs, out)
        _call_outs(e[3], funcs, out)
    elif k == "un":
        _call_outs(e[2], funcs, out)
    elif k == "cond":
        for s in e[1:]:
            _call_outs(s, funcs, out)


def _assigned(st, out: Set[str], funcs):
    k = st[0]
    if k == "assign":
        out.add(st[1])
        _call_outs(st[2], funcs, out)
    elif k == "contrib":
        _call_outs(st[3], funcs, out)
    elif k == "block":
        for s in st[2]:
            _assigned(s, out, funcs)
    elif k == "if":
        _assigned(st[2], out, funcs)
        if st[3]:
            _assigned(st[3], out, funcs)
    elif k == "case":
        for _, s in st[2]:
            _assigned(s, out, funcs)
        if st[3]:
            _assigned(st[3], out, funcs)
    elif k == "for":
        _assigned(st[1], out, funcs)
        _assigned(st[3], out, funcs)
        _assigned(st[4], out, funcs)
    elif k in ("while", "repeat"):
        _assigned(st[2], out, funcs)


def prune_dead(st, live: Set[str], funcs) -> Tuple[Optional[tuple], Set[str]]:
    """Backward liveness over structured code. Returns (pruned statement or None, live-before)."""
    k = st[0]
    if k == "assign":
        outs: Set[str] = set()
        _call_outs(st[2], funcs, outs)
        if st[1] not in live and not (outs & live):
            return None, live
        nl = set(live)
        nl.discard(st[1])
        _expr_vars(st[2], nl, funcs)
        return st, nl
    if k == "contrib":
        nl = set(live)
        _expr_vars(st[3], nl, funcs)
        return st, nl
    if k in ("task", "nop"):
        return None, live
    if k == "block":
        kept = []
        cur = live
        for s in reversed(st[2]):
            ns, cur = prune_dead(s, cur, funcs)
            if ns is not None:
                kept.append(ns)
        if not kept:
            return None, cur
        return ("block", st[1], kept[::-1], st[3]), cur
    if k == "if":
        a, la = prune_dead(st[2], live, funcs)
        b, lb = prune_dead(st[3], live, funcs) if st[3] else (None, live)
        if a is None and b is None:
            return None, live
        nl = set(la) | set(lb)
        _expr_vars(st[1], nl, funcs)
        return ("if", st[1], a if a is not None else ("nop",), b), nl
    if k == "case":
        items, nl, any_kept = [], set(), False
        for vals, s in st[2]:
            ns, ls = prune_dead(s, live, funcs)
            any_kept |= ns is not None
            items.append((vals, ns if ns is not None else ("nop",)))
            nl |= ls
            for v in vals:
                _expr_vars(v, nl, funcs)
        d, ld = prune_dead(st[3], live, funcs) if st[3] else (None, live)
        any_kept |= d is not None
        nl |= ld
        if not any_kept:
            return None, live
        _expr_vars(st[1], nl, funcs)
        return ("case", st[1], items, d), nl
    # loops: keep everything, everything read or written inside stays live
    nl = set(live)
    asg: Set[str] = set()
    _assigned(st, asg, funcs)

    def all_vars(s):
        kk = s[0]
        if kk == "assign":
            _expr_vars(s[2], nl, funcs)
        elif kk == "contrib":
            _expr_vars(s[3], nl, funcs)
        elif kk == "block":
            for x in s[2]:
                all_vars(x)
        elif kk == "if":
            _expr_vars(s[1], nl, funcs)
            all_vars(s[2])
            if s[3]:
                all_vars(s[3])
        elif kk == "case":
            _expr_vars(s[1], nl, funcs)
            for vals, x in s[2]:
                for v in vals:
                    _expr_vars(v, nl, funcs)
                all_vars(x)
            if s[3]:
                all_vars(s[3])
        elif kk == "for":
            all_vars(s[1]); _expr_vars(s[2], nl, funcs); all_vars(s[3]); all_vars(s[4])
        elif kk in ("while", "repeat"):
            _expr_vars(s[1], nl, funcs); all_vars(s[2])

    all_vars(st)
    return st, nl | asg


# ---------------------------------------------------------------------------------------
# values
# ---------------------------------------------------------------------------------------

@dataclass
class Val:
    """A generated value in the eval (dynamic) stream: an atom plus derivative atoms."""
    typ: str                      # 'r' | 'i'
    c: str                        # C atom (identifier or literal)
    d: Dict[int, object] = field(default_factory=dict)  # seed -> atom (str) or python float
    const: Optional[float] = None


@dataclass
class Lazy:
    """A not-yet-emitted static expression (depends on parameters / temperature only)."""
    e: tuple
    typ: str
    const: Optional[float] = None


@dataclass
class VS:
    dyn: bool = False
    deps: FrozenSet[int] = frozenset()
    slot: Optional[int] = None
    const: Optional[float] = 0.0   # compile-time constant value of a static variable


@dataclass
class CompiledModel:
    name: str
    module: str
    terminals: List[str]
    nports: int
    params: List[str]
    param_types: List[str]
    ncache: int
    jrow: List[int]
    jcol: List[int]
    source: str          # target-neutral body (macros), see c_prelude()/cuda wrappers in engine
    n_eval_lines: int = 0
    n_setup_lines: int = 0
    census: Dict[str, int] = field(default_factory=dict)  # static op counts of the eval function
    exec_ops: List[int] = field(default_factory=list)     # (add, mul, div, special) on the executed path, see models.py
    param_defaults: Dict[str, float] = field(default_factory=dict)  # constant defaults of the module's parameters
    # value-only variant (currents and charges, no derivatives) for the engine's chord iterations: its own setup /
    # eval pair (VA_SETUPV_* / VA_EVALV_* macros) and cache layout; empty when the module uses ddx()
    source_v: str = ""
    ncache_v: int = 0
    # noise variant (VA_SETUPN_* / VA_EVALN_*): power pwr_k (and flicker exponent) of every noise source at a bias
    # point; sources are independent current sources between two terminals (-1 = ground)
    source_n: str = ""
    ncache_n: int = 0
    noise_sources: List[Tuple[int, int, str, str]] = field(default_factory=list)
    branch_terms: List[int] = field(default_factory=list)   # terminals that are branch CURRENTS (voltage branches, I() probes)
    linear: bool = False   # every Jacobian entry is independent of the terminal values (no Newton step limiting needed)
    gen_version: int = 0   # GEN_VERSION of the generator that wrote `source` (cached models of another version are rebuilt)
    codegen_seconds: float = 0.0   # wall time of parsing + code generation of all variants (bench.py reports it as compile latency)
    # cached values that do not depend on any instance parameter (only on the model card, temperature, gmin): one small
    # table per model instead of a slot in every instance's cache row (CACHE_LDU / CACHE_STU); per variant
    nuni: int = 0
    nuni_v: int = 0
    nuni_n: int = 0

    @property
    def key(self) -> str:
        return hashlib.sha1(self.source.encode()).hexdigest()[:16]


class _Compiler:
    def __init__(self, mod: Module, name: str, const_params: Optional[Dict[str, float]] = None,
                 runtime_params: Optional[Sequence[str]] = None, no_deriv: bool = False, skip_funcs=(),
                 noise: bool = False, probe_branches: Sequence[Tuple[str, str]] = (), part: Optional[str] = None):
        self.mod = mod
        self.part = part                  # None | "I" | "Q": keep only the resistive / only the ddt() part of every contribution
        self.name = name
        self.noise = noise                # noise variant: outputs the power of every noise source, nothing else
        self.noise_sources: List[Tuple[int, int, str, str]] = []   # (pos terminal, neg terminal, kind, name), -1 = ground
        no_deriv = no_deriv or noise
        self.no_deriv = no_deriv          # value-only variant: probes carry no seeds, no Jacobian outputs
        self.skip_funcs = set(skip_funcs)  # helper functions the full variant already defined
        # specialisation: parameters with compile-time values (a model card) are folded; only
        # `runtime_params` are read through PAR(); every other parameter takes its default
        self.const_params = None if const_params is None else {k.upper(): v for k, v in const_params.items()}
        self.runtime_params = None if runtime_params is None else {k.upper() for k in runtime_params}
        self.terms = list(mod.nets)
        self.tindex = {n: i for i, n in enumerate(self.terms)}
        # Branches that need a current unknown (src/vasim.jl:792,802-808: "only used ones get a current unknown"): the
        # target of a `V(a,b) <+` contribution ('V') and a branch whose current is probed with `I(a,b)` ('P').  Each is
        # one more terminal "I(a,b)" after the internal nodes; the circuit allocates it among the branch currents.
        #   row of the unknown:  'V':  V(a) - V(b) - sum(contributions) = 0      'P':  I_br - sum(contributions) = 0
        #   KCL:                 +I_br leaves a, enters b (src/simulate_ir.jl:112-120)
        # `probe_branches`: branches whose current the caller wants as an observable (sys.<inst>.var"I(p, n)",
        # src/vasim.jl:786-808) -- they become 'P' branches, i.e. unknowns the solution carries
        self.branch_kind: Dict[Tuple[str, str], str] = {}
        for a, b in probe_branches:
            if a not in self.tindex and a != "0" or b not in self.tindex and b != "0":
                raise VACompileError(f"module {mod.name} has no branch ({a}, {b})")
            if (b, a) not in self.branch_kind:
                self.branch_kind[(a, b)] = "P"
        self._scan_branches(list(mod.analog))
        self.branch_terms: List[int] = []
        for (a, b) in self.branch_kind:
            self.tindex[f"I({a},{b})"] = len(self.terms)
            self.branch_terms.append(len(self.terms))
            self.terms.append(f"I({a},{b})")
        self.S: List[str] = []      # setup stream
        self.E: List[str] = []      # eval stream
        self.state: Dict[str, VS] = {}
        self.scopes: List[Dict[str, str]] = [{}]
        self.types: Dict[str, str] = {}           # cname -> 'r' | 'i'
        self.decl_deps: Dict[str, Set[int]] = {}  # cname -> all seeds ever carried
        self.dyn_vars: Set[str] = set()
        self.static_vars: Set[str] = set()
        self.ntemp = 0
        self.nslot = 0
        self.ninl = 0
        self.census: Dict[str, int] = {}
        self.forced_all: Set[str] = set()         # variables forced to carry all seeds (loops)
        self.dead_locals: Set[str] = set()        # locals of inlined functions / blocks out of scope
        self.block_ops: Dict[int, List[int]] = {}
        self.params = {p.name: p for p in mod.params}
        for p in mod.params:
            cn = "p_" + p.name
            self.types[cn] = "i" if p.type == "integer" else "r"
        for v, t in mod.var_types.items():
            self.types["v_" + v] = "i" if t == "integer" else "r"

    # ---- naming ---------------------------------------------------------------------
    def resolve(self, name: str) -> str:
        for sc in reversed(self.scopes):
            if name in sc:
                return sc[name]
        if name in self.params:
            return "p_" + name
        if name in self.mod.var_types:
            return "v_" + name
        raise VACompileError(f"undeclared identifier {name!r} in module {self.mod.name}")

    def tmp(self) -> str:
        self.ntemp += 1
        return f"t{self.ntemp}"

    def count(self, op: str, n: int = 1):
        self.census[op] = self.census.get(op, 0) + n
        # per-basic-block tallies, emitted as VA_OPS(add, mul, div, special) so that an
        # instrumented host build can report the op count of the *executed* path
        if n:
            blk = self.block_ops.setdefault(id(self.E), [0, 0, 0, 0])
            blk[{"add": 0, "mul": 1, "div": 2}.get(op, 3)] += n

    def flush_ops(self, lst: List[str]):
        blk = self.block_ops.pop(id(lst), None)
        if blk and any(blk):
            lst.append(f"VA_OPS({blk[0]}, {blk[1]}, {blk[2]}, {blk[3]});")

    def vs(self, cname: str) -> VS:
        s = self.state.get(cname)
        if s is None:
            s = VS(const=None) if cname.startswith("p_") else VS()
            self.state[cname] = s
        return s

    # ---- static (setup) expression generation: plain nested C --------------------------
    def gs(self, e) -> Tuple[str, str, Optional[float]]:
        """Returns (c expression, type, compile-time constant or None); may emit to self.S."""
        k = e[0]
        if k == "num":
            return _lit(e[1]), ("i" if e[2] else "r"), e[1]
        if k == "str":
            raise VACompileError("string expression in numeric context")
        if k == "var":
            cn = self.resolve(e[1])
            s = self.vs(cn)
            if s.dyn:
                raise VACompileError(f"internal: static read of dynamic variable {e[1]}")
            typ = self.types[cn]
            if s.const is not None:
                cv = int(s.const) if typ == "i" else float(s.const)
                return _lit(cv), typ, cv
            self.static_vars.add(cn)
            return cn, typ, None
        if k == "un":
            a, ta, ca = self.gs(e[2])
            op = e[1]
            if op == "-":
                return f"(-{a})", ta, (None if ca is None else -ca)
            if op == "!":
                return f"(!{a})", "i", (None if ca is None else int(not ca))
            if op == "~":
                return f"(~{self._as_int(a, ta)})", "i", None
        if k == "bin":
            op = e[1]
            a, ta, ca = self.gs(e[2])
            b, tb, cb = self.gs(e[3])
            return self._bin_static(op, a, ta, ca, b, tb, cb)
        if k == "cond":
            c, tc, cc = self.gs(e[1])
            if cc is not None:
                return self.gs(e[2] if cc else e[3])
            a, ta, _ = self.gs(e[2])
            b, tb, _ = self.gs(e[3])
            typ = "i" if ta == "i" and tb == "i" else "r"
            return f"({c} ? {self._cast(a, ta, typ)} : {self._cast(b, tb, typ)})", typ, None
        if k == "call":
            return self._call_static(e)
        if k == "probe":
            raise VACompileError("internal: probe in static expression")
        raise VACompileError(f"unsupported expression node {k}")

    @staticmethod
    def _cast(c: str, frm: str, to: str) -> str:
        if frm == to:
            return c
        if to == "r":
            return f"((double){c})"
        return f"((int)round({c}))"   # vaconvert: ties away from zero (src/va_env.jl:107)

    def _as_int(self, c, t):
        return c if t == "i" else f"((int)round({c}))"

    def _bin_static(self, op, a, ta, ca, b, tb, cb):
        both_int = ta == "i" and tb == "i"
        const = None
        if ca is not None and cb is not None:
            try:
                const = self._fold(op, ca, cb, both_int)
            except (ZeroDivisionError, OverflowError, ValueError):
                const = None
        if op in ("+", "-", "*"):
            typ = "i" if both_int else "r"
            if const is not None:
                return _lit(const), typ, const
            return f"({a} {op} {b})", typ, None
        if op == "/":
            typ = "i" if both_int else "r"
            if const is not None:
                return _lit(const), typ, const
            if both_int:
                return f"({a} / {b})", "i", None
            return f"({self._cast(a, ta, 'r')} / {self._cast(b, tb, 'r')})", "r", None
        if op == "%":
            if both_int:
                return f"({a} % {b})", "i", const
            return f"fmod({a}, {b})", "r", const
        if op == "**":
            return f"pow({self._cast(a, ta, 'r')}, {self._cast(b, tb, 'r')})", "r", const
        if op in ("==", "!=", "<", "<=", ">", ">="):
            if const is not None:
                return _lit(int(const)), "i", int(const)
            return f"({a} {op} {b})", "i", None
        if op in ("&&", "||"):
            # the reference lowers these to non-short-circuit ops (src/vasim.jl:225-228); same value
            if const is not None:
                return _lit(int(const)), "i", int(const)
            if op == "&&" and ((ca is not None and not ca) or (cb is not None and not cb)):
                return "0", "i", 0
            if op == "||" and ((ca is not None and ca) or (cb is not None and cb)):
                return "1", "i", 1
            return f"(({a} != 0) {op} ({b} != 0))", "i", None
        if op in ("&", "|", "^", "<<", ">>"):
            return f"({self._as_int(a, ta)} {op} {self._as_int(b, tb)})", "i", None
        if op in ("~^", "^~"):
            return f"(~({self._as_int(a, ta)} ^ {self._as_int(b, tb)}))", "i", None
        raise VACompileError(f"unsupported operator {op}")

    @staticmethod
    def _fold(op, a, b, both_int):
        if op == "+": return a + b
        if op == "-": return a - b
        if op == "*": return a * b
        if op == "/":
            if both_int:
                q = abs(a) // abs(b)
                return q if (a >= 0) == (b >= 0) else -q
            return a / b
        if op == "%": return math.fmod(a, b) if not both_int else int(math.fmod(a, b))
        if op == "**": return float(a) ** float(b)
        if op == "==": return int(a == b)
        if op == "!=": return int(a != b)
        if op == "<": return int(a < b)
        if op == "<=": return int(a <= b)
        if op == ">": return int(a > b)
        if op == ">=": return int(a >= b)
        if op == "&&": return int(bool(a) and bool(b))
        if op == "||": return int(bool(a) or bool(b))
        raise ValueError(op)

    def _call_static(self, e):
        fn, args = e[1], e[2]
        if fn == "$temperature":
            return "TEMP_K", "r", None
        if fn == "$vt":
            if args:
                a, ta, _ = self.gs(args[0])
                return f"({_lit(P_K / P_Q)} * {a})", "r", None
            return f"({_lit(P_K / P_Q)} * TEMP_K)", "r", None
        if fn == "$param_given":
            p = self.params.get(args[0][1]) if args and args[0][0] == "var" else None
            if p is None:
                raise VACompileError("$param_given needs a parameter name")
            if p.name in getattr(self, "given_const", {}):
                g = self.given_const[p.name]
                return str(g), "i", g
            return f"GIVEN({p.index})", "i", None
        if fn == "$simparam":
            key = args[0][1] if args and args[0][0] == "str" else None
            if key == "gmin":
                return "GMIN_V", "r", None
            if len(args) > 1:
                return self.gs(args[1])
            raise VACompileError(f"$simparam({key!r}) without default")
        if fn in ("$mfactor",):
            return "1.0", "r", 1.0
        if fn in ("$abstime", "$realtime"):
            return "0.0", "r", 0.0
        if fn in NOISE_FUNCS:
            return "0.0", "r", 0.0
        if fn in ("ddt", "idt", "ddx"):
            return "0.0", "r", 0.0   # derivative of a bias-independent quantity
        if fn in self.mod.functions:
            f = self.mod.functions[fn]
            if len(args) != len(f.args):
                raise VACompileError(f"wrong number of arguments to function {fn}")
            if all(kind == "input" for _, kind in f.args):
                consts = [self.gs(a)[2] for a in args]
                if all(c is not None for c in consts):
                    try:
                        cv = _ConstEval(self.mod.functions).call(fn, list(consts))
                        if isinstance(cv, int) or math.isfinite(cv):
                            return _lit(cv), ("i" if f.type == "integer" else "r"), cv
                    except _NotConst:
                        pass
            cargs = []
            for (an, kind), a in zip(f.args, args):
                if kind == "input":
                    c, t, _ = self.gs(a)
                    cargs.append(self._cast(c, t, "i" if f.var_types.get(an) == "integer" else "r"))
                else:
                    if a[0] != "var":
                        raise VACompileError(f"output argument {an} of {fn} must be a variable")
                    cn = self.resolve(a[1])
                    s = self.vs(cn)
                    s.const, s.slot = None, None
                    self.static_vars.add(cn)
                    cargs.append("&" + cn)
            self.used_funcs.add(fn)
            return f"f_{self.name}_{fn}({', '.join(cargs)})", ("i" if f.type == "integer" else "r"), None
        if fn in MATH1 or fn in MATH2:
            cs = [self.gs(a) for a in args]
            cstr = [self._cast(c, t, "r") for c, t, _ in cs]
            consts = [c[2] for c in cs]
            cname = {"ln": "log", "log": "log10", "abs": "fabs", "min": "fmin", "max": "fmax",
                     "limexp": "va_limexp"}.get(fn, fn)
            typ = "r"
            if fn in ("abs", "min", "max") and all(t == "i" for _, t, _ in cs):
                typ = "i"
                cstr = [c for c, _, _ in cs]
                if fn == "abs":
                    return f"abs({cstr[0]})", "i", (None if consts[0] is None else abs(consts[0]))
                o = "<" if fn == "min" else ">"
                return f"(({cstr[0]}) {o} ({cstr[1]}) ? ({cstr[0]}) : ({cstr[1]}))", "i", None
            const = None
            if all(c is not None for c in consts):
                try:
                    pyf = _PYF.get(fn)
                    if pyf is not None:
                        const = float(pyf(*[float(c) for c in consts]))
                except (ValueError, OverflowError, ZeroDivisionError):
                    const = None
            if const is not None and math.isfinite(const):
                return _lit(const), "r", const
            return f"{cname}({', '.join(cstr)})", typ, None
        raise VACompileError(f"unsupported function {fn}")

    # ---- staging helpers -----------------------------------------------------------------
    def is_static_expr(self, e) -> bool:
        k = e[0]
        if k in ("num", "str"):
            return True
        if k == "var":
            return not self.vs(self.resolve(e[1])).dyn
        if k == "probe":
            return False
        if k == "un":
            return self.is_static_expr(e[2])
        if k == "bin":
            return self.is_static_expr(e[2]) and self.is_static_expr(e[3])
        if k == "cond":
            return all(self.is_static_expr(s) for s in e[1:])
        if k == "call":
            if e[1] in NOISE_FUNCS or e[1].startswith("$"):
                return True
            fn = self.mod.functions.get(e[1])
            if fn is not None:
                for (an, kind), a in zip(fn.args, e[2]):
                    if kind == "input" and not self.is_static_expr(a):
                        return False
                    if kind != "input" and self.dynctl:
                        return False
                return True
            if e[1] in ("ddt", "ddx", "idt"):
                return all(self.is_static_expr(a) for a in e[2][:1])
            return all(self.is_static_expr(a) for a in e[2])
        return False

    def new_slot(self, cexpr: str) -> int:
        slot = self.nslot
        self.nslot += 1
        self.S.append(f"CACHE_ST({slot}, {cexpr});")
        return slot

    def static_read(self, cname: str) -> Val:
        """Read a static variable from the eval stream (through its cache slot)."""
        s = self.vs(cname)
        typ = self.types[cname]
        if s.const is not None:
            cv = int(s.const) if typ == "i" else float(s.const)
            return Val(typ, _lit(cv), {}, cv)
        if s.slot is None:
            self.static_vars.add(cname)
            s.slot = self.new_slot(cname)
        ld = f"CACHE_LD({s.slot})"
        return Val(typ, f"((int){ld})" if typ == "i" else ld, {})

    def hoist(self, e) -> Val:
        """Evaluate a static expression in setup and read it back in eval."""
        c, t, const = self.gs(e)
        if const is not None:
            return Val(t, _lit(const), {}, const)
        if e[0] == "var":
            return self.static_read(self.resolve(e[1]))
        slot = self.new_slot(c)
        ld = f"CACHE_LD({slot})"
        return Val(t, f"((int){ld})" if t == "i" else ld, {})

    # ---- dynamic (eval) expression generation with derivatives --------------------------
    def emit_val(self, typ: str, cexpr: str, derivs: Dict[int, str]) -> Val:
        t = self.tmp()
        self.E.append(f"const {'int' if typ == 'i' else 'double'} {t} = {cexpr};")
        d = {}
        for k, de in derivs.items():
            if isinstance(de, (int, float)):
                if de != 0:
                    d[k] = float(de)
                continue
            nm = f"{t}_d{k}"
            self.E.append(f"const double {nm} = {de};")
            d[k] = nm
        return Val(typ, t, d)

    @staticmethod
    def _datom(x) -> str:
        return _lit(float(x)) if isinstance(x, (int, float)) else x

    def _dmul(self, coef: str, x) -> object:
        """coef * x for a derivative atom x (python float or C atom)"""
        if isinstance(x, (int, float)):
            if x == 0:
                return 0.0
            if x == 1:
                return coef
            if x == -1:
                return f"(-{coef})"
            return f"({coef} * {_lit(float(x))})"
        return f"({coef} * {x})"

    def gd(self, e):
        """Generate e in the dynamic stream; returns Val or Lazy (maximal static subtree)."""
        k = e[0]
        if k == "num":
            return Lazy(e, "i" if e[2] else "r", e[1])
        if k == "str":
            return Lazy(e, "s")
        if self.is_static_expr(e):
            if k == "var":
                cn = self.resolve(e[1])
                s = self.vs(cn)
                return Lazy(e, self.types[cn], s.const)
            return Lazy(e, "?")
        if k == "var":
            cn = self.resolve(e[1])
            s = self.vs(cn)
            typ = self.types[cn]
            self.dyn_vars.add(cn)
            return Val(typ, cn, {kk: f"{cn}__d{kk}" for kk in sorted(s.deps)} if typ == "r" else {})
        if k == "probe":
            return self._probe(e)
        if k == "un":
            a = self.force(self.gd(e[2]))
            op = e[1]
            if op == "-":
                if a.typ == "i":
                    return self.emit_val("i", f"-{a.c}", {})
                self.count("add", 1 + len(a.d))
                d = {kk: (-v if isinstance(v, (int, float)) else f"-{v}") for kk, v in a.d.items()}
                return self.emit_val("r", f"-{a.c}", d)
            if op == "!":
                return self.emit_val("i", f"!{a.c}", {})
            if op == "~":
                return self.emit_val("i", f"~{self._as_int(a.c, a.typ)}", {})
        if k == "bin":
            return self._bin_dyn(e)
        if k == "cond":
            c = self.force(self.gd(e[1]))
            a = self.force(self.gd(e[2]))
            b = self.force(self.gd(e[3]))
            typ = "i" if a.typ == "i" and b.typ == "i" else "r"
            ac, bc = self._cast(a.c, a.typ, typ), self._cast(b.c, b.typ, typ)
            d = {}
            for kk in sorted(set(a.d) | set(b.d)):
                d[kk] = f"({c.c} ? {self._datom(a.d.get(kk, 0.0))} : {self._datom(b.d.get(kk, 0.0))})"
            return self.emit_val(typ, f"({c.c} ? {ac} : {bc})", d)
        if k == "call":
            return self._call_dyn(e)
        raise VACompileError(f"unsupported expression node {k}")

    def force(self, x) -> Val:
        if isinstance(x, Val):
            return x
        if x.const is not None and x.typ in ("r", "i"):
            cv = int(x.const) if x.typ == "i" else float(x.const)
            return Val(x.typ, _lit(cv), {}, cv)
        return self.hoist(x.e)

    def _probe(self, e) -> Val:
        acc, nodes = e[1], e[2]
        if acc in ("I", "flow"):
            bro = self._branch_of(nodes)
            if bro is None:
                raise VACompileError(f"probe {acc}({', '.join(nodes)}): branch has no current unknown")
            br, sign, _ = bro
            return self.emit_val("r", f"VT({br})" if sign > 0 else f"-VT({br})", {} if self.no_deriv else {br: sign})
        if acc not in ("V", "potential"):
            raise VACompileError(f"probe {acc}({', '.join(nodes)}) is not supported")
        if len(nodes) == 1 and nodes[0] in self.mod.branches:
            nodes = list(self.mod.branches[nodes[0]])
        idx = []
        for n in nodes:
            if n in ("0", "gnd") or n not in self.tindex:
                if n in ("0", "gnd"):
                    continue
                raise VACompileError(f"unknown net {n}")
            idx.append(self.tindex[n])
        if len(nodes) == 2 and len(idx) == 2:
            a, b = idx
            if a == b:
                return Val("r", "0.0", {}, 0.0)
            self.count("add")
            d = {} if self.no_deriv else {a: 1.0, b: -1.0}
            d.pop(getattr(self, "drop_seed", None), None)
            return self.emit_val("r", f"VT({a}) - VT({b})", d)
        if len(idx) == 1:
            sign = 1.0 if (len(nodes) == 1 or nodes[0] not in ("0", "gnd")) else -1.0
            a = idx[0]
            return self.emit_val("r", f"VT({a})" if sign > 0 else f"-VT({a})", {} if self.no_deriv else {a: sign})
        return Val("r", "0.0", {}, 0.0)

    def _bin_dyn(self, e) -> Val:
        op = e[1]
        a = self.force(self.gd(e[2]))
        b = self.force(self.gd(e[3]))
        both_int = a.typ == "i" and b.typ == "i"
        if op in ("==", "!=", "<", "<=", ">", ">="):
            return self.emit_val("i", f"{a.c} {op} {b.c}", {})
        if op in ("&&", "||"):
            return self.emit_val("i", f"({a.c} != 0) {op} ({b.c} != 0)", {})
        if op in ("&", "|", "^", "<<", ">>"):
            return self.emit_val("i", f"{self._as_int(a.c, a.typ)} {op} {self._as_int(b.c, b.typ)}", {})
        if op in ("~^", "^~"):
            return self.emit_val("i", f"~({self._as_int(a.c, a.typ)} ^ {self._as_int(b.c, b.typ)})", {})
        if both_int:
            if op in ("+", "-", "*", "/", "%"):
                return self.emit_val("i", f"{a.c} {op} {b.c}", {})
        ac, bc = self._cast(a.c, a.typ, "r"), self._cast(b.c, b.typ, "r")
        keys = sorted(set(a.d) | set(b.d))
        if op in ("+", "-"):
            self.count("add", 1 + len(keys))
            d = {}
            for kk in keys:
                da, db = a.d.get(kk), b.d.get(kk)
                if db is None:
                    d[kk] = da
                elif da is None:
                    d[kk] = db if op == "+" else (-db if isinstance(db, (int, float)) else f"-{db}")
                elif isinstance(da, (int, float)) and isinstance(db, (int, float)):
                    d[kk] = da + db if op == "+" else da - db
                else:
                    d[kk] = f"{self._datom(da)} {op} {self._datom(db)}"
            return self.emit_val("r", f"{ac} {op} {bc}", d)
        if op == "*":
            d = {}
            for kk in keys:
                da, db = a.d.get(kk), b.d.get(kk)
                terms = []
                if da is not None:
                    terms.append(self._dmul(bc, da))
                if db is not None:
                    terms.append(self._dmul(ac, db))
                terms = [t for t in terms if not (isinstance(t, float) and t == 0.0)]
                d[kk] = " + ".join(self._datom(t) for t in terms) if terms else 0.0
                self.count("mul", len(terms)); self.count("add", max(0, len(terms) - 1))
            self.count("mul")
            return self.emit_val("r", f"{ac} * {bc}", d)
        if op == "/":
            if b.const is not None and not b.d and b.typ in ("r", "i") and float(b.const) != 0.0 and FOLD_CONST_DIV:
                # division by a compile-time constant (a folded model-card value): multiply by its reciprocal, computed
                # here in IEEE double.  On the GPU a division is an out-of-line call (csrc/va_prelude.h) that the compiler
                # cannot fold; both targets compile the same text, so oracle and engine stay on the same arithmetic.
                r = 1.0 / float(b.const)
                if math.isfinite(r) and r != 0.0:
                    if r == 1.0:
                        return a if a.typ == "r" else self.emit_val("r", ac, {})
                    self.count("mul", 1 + len(a.d))
                    rl = _lit(r)
                    return self.emit_val("r", f"{ac} * {rl}", {kk: self._dmul(rl, v) for kk, v in a.d.items()})
            self.count("div")
            if not b.d:
                if not a.d:
                    return self.emit_val("r", f"{ac} * VA_RCP({bc})", {})
                inv = self.emit_val("r", f"VA_RCP({bc})", {})
                d = {kk: self._dmul(inv.c, v) for kk, v in a.d.items()}
                self.count("mul", 1 + len(d))
                return self.emit_val("r", f"{ac} * {inv.c}", d)
            inv = self.emit_val("r", f"VA_RCP({bc})", {})
            q = self.emit_val("r", f"{ac} * {inv.c}", {})
            d = {}
            for kk in keys:
                da, db = a.d.get(kk), b.d.get(kk)
                if db is None:
                    d[kk] = self._dmul(inv.c, da)
                    self.count("mul")
                elif da is None:
                    d[kk] = f"-({q.c} * {self._datom(db)}) * {inv.c}"
                    self.count("mul", 2)
                else:
                    d[kk] = f"({self._datom(da)} - {q.c} * {self._datom(db)}) * {inv.c}"
                    self.count("mul", 2); self.count("add")
            return self.emit_val("r", q.c, d)
        if op == "%":
            return self.emit_val("r", f"fmod({ac}, {bc})", dict(a.d))
        if op == "**":
            return self._pow(a, b)
        raise VACompileError(f"unsupported operator {op}")

    def _pow(self, a: Val, b: Val) -> Val:
        ac, bc = self._cast(a.c, a.typ, "r"), self._cast(b.c, b.typ, "r")
        # strength reduction for compile-time exponents (card-specialised code has many pow(x, 2|3|4))
        if b.const is not None and not b.d:
            e = float(b.const)
            if e == 0.0:
                return Val("r", "1.0", {}, 1.0)
            if e == 1.0:
                return Val("r", ac, dict(a.d)) if a.typ == "r" else self.emit_val("r", ac, {})
            if e == 0.5:
                return self._unary_math("sqrt", a)
            if e == float(int(e)) and 2 <= abs(e) <= 8:
                n = int(abs(e))
                base = self.emit_val("r", ac, {})
                # x^(n-1) by repeated multiplication, then value and derivative n*x^(n-1)
                pm1 = base.c
                for _ in range(n - 2):
                    pm1 = self.emit_val("r", f"{pm1} * {base.c}", {}).c
                    self.count("mul")
                val = self.emit_val("r", f"{pm1} * {base.c}", {})
                self.count("mul")
                if e > 0:
                    if not a.d:
                        return val
                    g = self.emit_val("r", f"{float(n)!r} * {pm1}", {})
                    self.count("mul", 1 + len(a.d))
                    return self.emit_val("r", val.c, {kk: self._dmul(g.c, x) for kk, x in a.d.items()})
                inv = self.emit_val("r", f"VA_RCP({val.c})", {})
                self.count("div")
                if not a.d:
                    return inv
                # d/dx x^-n = -n x^-n / x
                g = self.emit_val("r", f"{-float(n)!r} * {inv.c} / {base.c}", {})
                self.count("div"); self.count("mul", 1 + len(a.d))
                return self.emit_val("r", inv.c, {kk: self._dmul(g.c, x) for kk, x in a.d.items()})
        self.count("pow")
        p = self.emit_val("r", f"pow({ac}, {bc})", {})
        keys = sorted(set(a.d) | set(b.d))
        if not keys:
            return p
        d = {}
        dpa = None
        if a.d:
            # d/da a^b = b a^(b-1)  (ForwardDiff's rule; finite at a == 0 for b >= 1)
            # (b == 0 -> 0 avoids 0*inf = NaN at a == 0, e.g. DVTP0*pow(vdsx, DVTP1) with DVTP1 = 0)
            # (VA_DPOW: one out-of-line body on the GPU, see csrc/va_prelude.h)
            dpa = self.emit_val("r", f"VA_DPOW({ac}, {bc}, {p.c})", {})
            self.count("div"); self.count("mul")
        dpb = None
        if b.d:
            dpb = self.emit_val("r", f"({p.c} == 0.0 ? 0.0 : {p.c} * log({ac}))", {})
            self.count("log"); self.count("mul")
        for kk in keys:
            terms = []
            if kk in a.d:
                terms.append(self._dmul(dpa.c, a.d[kk]))
            if kk in b.d:
                terms.append(self._dmul(dpb.c, b.d[kk]))
            self.count("mul", len(terms))
            d[kk] = " + ".join(self._datom(t) for t in terms)
        return self.emit_val("r", p.c, d)

    def _unary_math(self, fn: str, a: Val) -> Val:
        ac = self._cast(a.c, a.typ, "r")
        has_d = bool(a.d)

        def chain(v: Val, dcoef: Optional[str]) -> Val:
            if not has_d or dcoef is None:
                return v
            g = self.emit_val("r", dcoef, {})
            self.count("mul", len(a.d))
            return self.emit_val("r", v.c, {kk: self._dmul(g.c, x) for kk, x in a.d.items()})

        if fn == "exp":
            self.count("exp")
            v = self.emit_val("r", f"exp({ac})", {})
            if not has_d:
                return v
            self.count("mul", len(a.d))
            return self.emit_val("r", v.c, {kk: self._dmul(v.c, x) for kk, x in a.d.items()})
        if fn == "limexp":
            self.count("exp")
            v = self.emit_val("r", f"va_limexp({ac})", {})
            return chain(v, f"va_dlimexp({ac})")
        if fn == "ln":
            self.count("log"); self.count("div", 1 if has_d else 0)
            return chain(self.emit_val("r", f"log({ac})", {}), f"VA_RCP({ac})")
        if fn == "log":
            self.count("log"); self.count("div", 1 if has_d else 0)
            return chain(self.emit_val("r", f"log10({ac})", {}), f"{_lit(1.0 / math.log(10.0))} * VA_RCP({ac})")
        if fn == "sqrt":
            self.count("sqrt"); self.count("div", 1 if has_d else 0)
            v = self.emit_val("r", f"VA_SQRT({ac})", {})
            return chain(v, f"0.5 * VA_RCP({v.c})")
        if fn == "sin":
            self.count("trig", 2 if has_d else 1)
            return chain(self.emit_val("r", f"sin({ac})", {}), f"cos({ac})")
        if fn == "cos":
            self.count("trig", 2 if has_d else 1)
            return chain(self.emit_val("r", f"cos({ac})", {}), f"-sin({ac})")
        if fn == "tan":
            self.count("trig")
            v = self.emit_val("r", f"tan({ac})", {})
            return chain(v, f"1.0 + {v.c} * {v.c}")
        if fn == "asin":
            self.count("trig")
            return chain(self.emit_val("r", f"asin({ac})", {}), f"VA_RCP(VA_SQRT(1.0 - {ac} * {ac}))")
        if fn == "acos":
            self.count("trig")
            return chain(self.emit_val("r", f"acos({ac})", {}), f"-VA_RCP(VA_SQRT(1.0 - {ac} * {ac}))")
        if fn == "atan":
            self.count("trig"); self.count("div", 1 if has_d else 0)
            return chain(self.emit_val("r", f"atan({ac})", {}), f"VA_RCP(1.0 + {ac} * {ac})")
        if fn == "sinh":
            self.count("exp", 2)
            return chain(self.emit_val("r", f"sinh({ac})", {}), f"cosh({ac})")
        if fn == "cosh":
            self.count("exp", 2)
            return chain(self.emit_val("r", f"cosh({ac})", {}), f"sinh({ac})")
        if fn == "tanh":
            self.count("exp"); self.count("div")
            v = self.emit_val("r", f"tanh({ac})", {})
            return chain(v, f"1.0 - {v.c} * {v.c}")
        if fn == "asinh":
            self.count("log")
            return chain(self.emit_val("r", f"asinh({ac})", {}), f"VA_RCP(VA_SQRT({ac} * {ac} + 1.0))")
        if fn == "acosh":
            self.count("log")
            return chain(self.emit_val("r", f"acosh({ac})", {}), f"VA_RCP(VA_SQRT({ac} * {ac} - 1.0))")
        if fn == "atanh":
            self.count("log")
            return chain(self.emit_val("r", f"atanh({ac})", {}), f"VA_RCP(1.0 - {ac} * {ac})")
        if fn == "abs":
            if a.typ == "i":
                return self.emit_val("i", f"abs({a.c})", {})
            v = self.emit_val("r", f"fabs({ac})", {})
            if not has_d:
                return v
            s = self.emit_val("r", f"({ac} < 0.0 ? -1.0 : 1.0)", {})
            return self.emit_val("r", v.c, {kk: self._dmul(s.c, x) for kk, x in a.d.items()})
        if fn in ("floor", "ceil"):
            return self.emit_val("r", f"{fn}({ac})", {})
        raise VACompileError(f"unsupported function {fn}")

    def _call_dyn(self, e):
        fn, args = e[1], e[2]
        if fn in NOISE_FUNCS:
            return Val("r", "0.0", {}, 0.0)
        if fn == "ddx" and self.no_deriv:
            raise VACompileError("ddx() needs derivatives: no value-only variant for this module")
        if fn == "ddx":
            v = self.force(self.gd(args[0]))
            pr = args[1]
            if pr[0] != "probe" or pr[1] not in ("V", "potential"):
                raise VACompileError("ddx: second argument must be a V() probe")
            idx = [self.tindex[n] for n in pr[2] if n in self.tindex]

            def dwrt(node):
                t_ = self.tindex.get(node, -1)
                if t_ >= 0 and t_ == getattr(self, "drop_seed", None):   # recovered by invariance
                    terms = [self._datom(x) for x in v.d.values()]
                    return "(-(" + " + ".join(terms) + "))" if terms else "0.0"
                return self._datom(v.d.get(t_, 0.0))

            if len(pr[2]) == 1:
                return self.emit_val("r", dwrt(pr[2][0]) if idx else "0.0", {})
            # two-node probe: (dx1 - dx2)/2 as the reference does (src/vasim.jl:398-411)
            d1 = dwrt(pr[2][0])
            d2 = dwrt(pr[2][1])
            return self.emit_val("r", f"({d1} - {d2}) * 0.5", {})
        if fn == "ddt":
            raise VACompileError("ddt() is only supported as a linear term of a contribution")
        if fn in self.mod.functions:
            return self._inline(self.mod.functions[fn], args)
        if fn in MATH1:
            return self._unary_math(fn, self.force(self.gd(args[0])))
        if fn == "pow":
            return self._pow(self.force(self.gd(args[0])), self.force(self.gd(args[1])))
        if fn in ("min", "max"):
            a = self.force(self.gd(args[0]))
            b = self.force(self.gd(args[1]))
            typ = "i" if a.typ == "i" and b.typ == "i" else "r"
            ac, bc = self._cast(a.c, a.typ, typ), self._cast(b.c, b.typ, typ)
            o = "<" if fn == "min" else ">"
            sel = self.emit_val("i", f"{ac} {o} {bc}", {})
            d = {}
            for kk in sorted(set(a.d) | set(b.d)):
                d[kk] = f"({sel.c} ? {self._datom(a.d.get(kk, 0.0))} : {self._datom(b.d.get(kk, 0.0))})"
            return self.emit_val(typ, f"({sel.c} ? {ac} : {bc})", d)
        if fn == "atan2":
            y = self.force(self.gd(args[0]))
            x = self.force(self.gd(args[1]))
            v = self.emit_val("r", f"atan2({y.c}, {x.c})", {})
            if not (y.d or x.d):
                return v
            den = self.emit_val("r", f"VA_RCP({x.c} * {x.c} + {y.c} * {y.c})", {})
            d = {}
            for kk in sorted(set(y.d) | set(x.d)):
                d[kk] = (f"({x.c} * {self._datom(y.d.get(kk, 0.0))} - {y.c} * {self._datom(x.d.get(kk, 0.0))})"
                         f" * {den.c}")
            return self.emit_val("r", v.c, d)
        if fn == "hypot":
            a = self.force(self.gd(args[0]))
            b = self.force(self.gd(args[1]))
            v = self.emit_val("r", f"hypot({a.c}, {b.c})", {})
            d = {}
            for kk in sorted(set(a.d) | set(b.d)):
                d[kk] = (f"({a.c} * {self._datom(a.d.get(kk, 0.0))} + {b.c} * {self._datom(b.d.get(kk, 0.0))})"
                         f" / {v.c}")
            return self.emit_val("r", v.c, d)
        raise VACompileError(f"unsupported function {fn} in bias-dependent code")

    def _inline(self, f: Function, args) -> Val:
        """Inline a user analog function into the dynamic stream (AD goes through its body)."""
        if len(args) != len(f.args):
            raise VACompileError(f"wrong number of arguments to function {f.name}")
        self.ninl += 1
        pre = f"f{self.ninl}_"
        scope = {}
        for vn, vt in f.var_types.items():
            cn = pre + vn
            scope[vn] = cn
            self.types[cn] = "i" if vt == "integer" else "r"
        ins = []
        for (an, kind), a in zip(f.args, args):
            if kind in ("input", "inout"):
                ins.append((an, self.gd(a)))
        outs = [(an, a) for (an, kind), a in zip(f.args, args) if kind in ("output", "inout")]
        for an, a in outs:
            if a[0] != "var":
                raise VACompileError(f"output argument {an} of {f.name} must be a variable")
        self.scopes.append(scope)
        for vn in f.var_types:
            self.state[scope[vn]] = VS()
        for an, v in ins:
            self._store(scope[an], v)
        self.stmt(f.body)
        self.scopes.pop()
        self.dead_locals.update(scope.values())
        for an, a in outs:
            self._store(self.resolve(a[1]), self._load(scope[an]))
        return self._load(scope[f.name])

    def _load(self, cn: str):
        s = self.vs(cn)
        typ = self.types[cn]
        if not s.dyn:
            if s.const is not None:
                return Lazy(("num", s.const, typ == "i"), typ, s.const)
            return self.static_read(cn)
        self.dyn_vars.add(cn)
        return Val(typ, cn, {kk: f"{cn}__d{kk}" for kk in sorted(s.deps)} if typ == "r" else {})

    def _store(self, cn: str, v):
        """Assign a generated value (Val or Lazy) to variable cn, in the right stream."""
        typ = self.types[cn]
        if isinstance(v, Lazy) and not self.dynctl:
            c, t, const = self.gs(v.e)
            s = VS(False, frozenset(), None, None)
            if const is not None:
                if typ == "i":   # vaconvert: round half away from zero (src/va_env.jl:107)
                    s.const = int(math.floor(abs(const) + 0.5)) * (1 if const >= 0 else -1) if t == "r" else int(const)
                else:
                    s.const = float(const)
            self.S.append(f"{cn} = {self._cast(c, t, typ)};")
            self.static_vars.add(cn)
            self.state[cn] = s
            return
        v = self.force(v)
        self.dyn_vars.add(cn)
        self.E.append(f"{cn} = {self._cast(v.c, v.typ, typ)};")
        deps = set()
        if typ == "r":
            for kk, x in v.d.items():
                self.E.append(f"{cn}__d{kk} = {self._datom(x)};")
                deps.add(kk)
            if cn in self.forced_all:
                for kk in range(len(self.terms)):
                    if kk not in deps:
                        self.E.append(f"{cn}__d{kk} = 0.0;")
                deps = set(range(len(self.terms)))
            self.decl_deps.setdefault(cn, set()).update(deps)
        self.state[cn] = VS(True, frozenset(deps), None, None)

    # ---- statements -----------------------------------------------------------------------
    dynctl = False

    def stmt(self, st):
        k = st[0]
        if k in ("nop", "task"):
            return
        if k == "block":
            scope = {}
            for vn, vt in st[3].items():
                self.ninl += 1
                cn = f"b{self.ninl}_{vn}"
                scope[vn] = cn
                self.types[cn] = "i" if vt == "integer" else "r"
                self.state[cn] = VS()
            self.scopes.append(scope)
            for s in st[2]:
                self.stmt(s)
            self.scopes.pop()
            self.dead_locals.update(scope.values())
            return
        if k == "assign":
            cn = self.resolve(st[1])
            if cn.startswith("p_"):
                raise VACompileError(f"assignment to parameter {st[1]}")
            self._store(cn, self.gd(st[2]))
            return
        if k == "contrib":
            return self.contrib(st)
        if k == "if":
            return self.if_stmt(st[1], st[2], st[3])
        if k == "case":
            chain = st[3]
            for vals, body in reversed(st[2]):
                cond = None
                for v in vals:
                    c = ("bin", "==", st[1], v)
                    cond = c if cond is None else ("bin", "||", cond, c)
                chain = ("if", cond, body, chain)
            if chain is not None:
                self.stmt(chain)
            return
        if k == "for":
            self.stmt(st[1])
            return self.loop(st[2], ("block", None, [st[4], st[3]], {}))
        if k == "while":
            return self.loop(st[1], st[2])
        if k == "repeat":
            self.ninl += 1
            cn = f"rep{self.ninl}"
            self.types[cn] = "i"
            self.scopes.append({cn: cn})
            self.stmt(("assign", cn, st[1]))
            body = ("block", None, [st[2], ("assign", cn, ("bin", "-", ("var", cn), ("num", 1, True)))], {})
            self.loop(("bin", ">", ("var", cn), ("num", 0, True)), body)
            self.scopes.pop()
            return
        raise VACompileError(f"unsupported statement {k}")

    # -- contributions (src/vasim.jl:128-180) --
    def _split_ddt(self, e):
        """e = resistive + sum coef_i * ddt(q_i).  Returns (resistive or None, [(coef, q)])."""
        k = e[0]
        if k == "call" and e[1] == "ddt":
            return None, [(("num", 1.0, False), e[2][0])]
        if not self._has_ddt(e):
            return e, []
        if k == "bin" and e[1] in ("+", "-"):
            ra, qa = self._split_ddt(e[2])
            rb, qb = self._split_ddt(e[3])
            if e[1] == "-":
                rb = None if rb is None else ("un", "-", rb)
                qb = [(("un", "-", c), q) for c, q in qb]
            r = ra if rb is None else (rb if ra is None else ("bin", "+", ra, rb))
            return r, qa + qb
        if k == "bin" and e[1] == "*":
            if self._has_ddt(e[2]) and not self._has_ddt(e[3]):
                r, q = self._split_ddt(e[2])
                other = e[3]
            elif self._has_ddt(e[3]) and not self._has_ddt(e[2]):
                r, q = self._split_ddt(e[3])
                other = e[2]
            else:
                raise VACompileError("product of two ddt() terms")
            r = None if r is None else ("bin", "*", other, r)
            return r, [(("bin", "*", other, c), qq) for c, qq in q]
        if k == "bin" and e[1] == "/" and not self._has_ddt(e[3]):
            r, q = self._split_ddt(e[2])
            r = None if r is None else ("bin", "/", r, e[3])
            return r, [(("bin", "/", c, e[3]), qq) for c, qq in q]
        if k == "un" and e[1] == "-":
            r, q = self._split_ddt(e[2])
            return (None if r is None else ("un", "-", r)), [(("un", "-", c), qq) for c, qq in q]
        raise VACompileError("ddt() must appear linearly in a contribution")

    def _strip_part(self, st):
        """Half of the model: every contribution keeps only its resistive part ("I") or only its ddt() part ("Q"); dead-code
        elimination then drops what only the other half needs."""
        k = st[0]
        if k == "contrib":
            e, noises = self._split_noise(st[3]) if not self.noise else (st[3], [])
            if e is None:
                return ("nop",)
            res, qs = self._split_ddt(e)
            if self.part == "I":
                return ("contrib", st[1], st[2], res) if res is not None else ("nop",)
            new = None
            for c, q in qs:
                term = ("bin", "*", c, ("call", "ddt", [q]))
                new = term if new is None else ("bin", "+", new, term)
            return ("contrib", st[1], st[2], new) if new is not None else ("nop",)
        if k == "block":
            return ("block", st[1], [self._strip_part(x) for x in st[2]], st[3])
        if k == "if":
            return ("if", st[1], self._strip_part(st[2]), self._strip_part(st[3]) if st[3] else None)
        if k == "case":
            return ("case", st[1], [(vals, self._strip_part(x)) for vals, x in st[2]], self._strip_part(st[3]) if st[3] else None)
        if k == "for":
            return ("for", st[1], st[2], st[3], self._strip_part(st[4]))
        if k in ("while", "repeat"):
            return (k, st[1], self._strip_part(st[2]))
        return st

    def _has_ddt(self, e) -> bool:
        k = e[0]
        if k == "call":
            return e[1] == "ddt" or any(self._has_ddt(a) for a in e[2])
        if k == "bin":
            return self._has_ddt(e[2]) or self._has_ddt(e[3])
        if k == "un":
            return self._has_ddt(e[2])
        if k == "cond":
            return any(self._has_ddt(s) for s in e[1:])
        return False

    # -- noise sources (reference: src/va_env.jl:82-90 makes every white_noise / flicker_noise call an epsilon
    #    variable that is 0 in DC / transient and a unit-PSD input of the small-signal system in noise!(),
    #    src/ac.jl:100-160: output PSD = sum_k |H_k(jw)|^2 pwr_k / f^exp_k) --
    def _has_noise(self, e) -> bool:
        k = e[0]
        if k == "call":
            return e[1] in NOISE_FUNCS or any(self._has_noise(a) for a in e[2])
        if k == "bin":
            return self._has_noise(e[2]) or self._has_noise(e[3])
        if k == "un":
            return self._has_noise(e[2])
        if k == "cond":
            return any(self._has_noise(s) for s in e[1:])
        return False

    def _split_noise(self, e):
        """e = rest + sum coef_i * noise_call_i.  Returns (rest or None, [(coef, call)])."""
        k = e[0]
        one = ("num", 1.0, False)
        if k == "call" and e[1] in NOISE_FUNCS:
            return None, [(one, e)]
        if not self._has_noise(e):
            return e, []
        if k == "bin" and e[1] in ("+", "-"):
            ra, na = self._split_noise(e[2])
            rb, nb = self._split_noise(e[3])
            if e[1] == "-":
                rb = None if rb is None else ("un", "-", rb)
            r = ra if rb is None else (rb if ra is None else ("bin", "+", ra, rb))
            return r, na + nb          # the sign of a noise term does not matter (powers)
        if k == "bin" and e[1] == "*":
            if self._has_noise(e[2]) and not self._has_noise(e[3]):
                r, n = self._split_noise(e[2])
                other = e[3]
            elif self._has_noise(e[3]) and not self._has_noise(e[2]):
                r, n = self._split_noise(e[3])
                other = e[2]
            else:
                raise VACompileError("product of two noise sources")
            r = None if r is None else ("bin", "*", other, r)
            return r, [(("bin", "*", other, c), q) for c, q in n]
        if k == "bin" and e[1] == "/" and not self._has_noise(e[3]):
            r, n = self._split_noise(e[2])
            r = None if r is None else ("bin", "/", r, e[3])
            return r, [(("bin", "/", c, e[3]), q) for c, q in n]
        if k == "un" and e[1] == "-":
            r, n = self._split_noise(e[2])
            return (None if r is None else ("un", "-", r)), n
        raise VACompileError("noise sources must appear linearly in a contribution")

    def _noise_contrib(self, pos, neg, noises):
        for coef, call in noises:
            fn, args = call[1], call[2]
            if fn == "noise_table":
                raise VACompileError("noise_table() is not supported")
            k = len(self.noise_sources)
            label = args[-1][1] if args and args[-1][0] == "str" else ""
            self.noise_sources.append((-1 if pos is None else pos, -1 if neg is None else neg,
                                       "flicker" if fn == "flicker_noise" else "white", str(label)))
            pwr = ("bin", "*", ("bin", "*", coef, coef), args[0])
            terms = [("N", pwr)]
            if fn == "flicker_noise":
                terms.append(("NE", args[1]))
            for kind, expr in terms:
                v = self.force(self.gd(expr))
                c = self._cast(v.c, v.typ, "r")
                an = f"acc{kind}_{k}"
                self.types[an] = "r"
                self.E.append(f"{an} = {c};")
                self.dyn_vars.add(an)
                self.decl_deps.setdefault(an, set())
                self.state[an] = VS(True, frozenset(), None, None)

    # ---- branches with a current unknown ---------------------------------------------------
    def _branch_nodes(self, nodes) -> Tuple[str, str]:
        nodes = list(nodes)
        if len(nodes) == 1 and nodes[0] in self.mod.branches:
            nodes = list(self.mod.branches[nodes[0]])
        g = lambda n: "0" if n in ("0", "gnd") else n
        return (g(nodes[0]), g(nodes[1]) if len(nodes) > 1 else "0")

    def _scan_branches(self, x):
        if isinstance(x, tuple):
            if len(x) >= 3 and x[0] == "contrib" and x[1] in ("V", "potential"):
                a, b = self._branch_nodes(x[2])
                if (b, a) in self.branch_kind:
                    self.branch_kind[(b, a)] = "V"
                else:
                    self.branch_kind[(a, b)] = "V"
            elif len(x) >= 3 and x[0] == "probe" and x[1] in ("I", "flow"):
                a, b = self._branch_nodes(x[2])
                if (a, b) not in self.branch_kind and (b, a) not in self.branch_kind:
                    self.branch_kind[(a, b)] = "P"
            for y in x:
                self._scan_branches(y)
        elif isinstance(x, list):
            for y in x:
                self._scan_branches(y)
        elif isinstance(x, dict):
            for y in x.values():
                self._scan_branches(y)

    def _branch_of(self, nodes):
        """(terminal index of the branch current, orientation sign, kind) or None for an ordinary branch."""
        a, b = self._branch_nodes(nodes)
        if (a, b) in self.branch_kind:
            return self.tindex[f"I({a},{b})"], 1.0, self.branch_kind[(a, b)]
        if (b, a) in self.branch_kind:
            return self.tindex[f"I({b},{a})"], -1.0, self.branch_kind[(b, a)]
        return None

    def _acc(self, kind: str, node: int, sign: float, v: Val):
        """acc<kind>_<node> += sign * v  (value and partials), in the eval stream"""
        an = f"acc{kind}_{node}"
        self.types[an] = "r"
        s = self.vs(an)
        op = "+=" if sign > 0 else "-="
        self.E.append(f"{an} {op} {v.c};")
        deps = set(s.deps) if s.dyn else set()
        for kk, x in v.d.items():
            self.E.append(f"{an}__d{kk} {op} {self._datom(x)};")
            deps.add(kk)
        self.count("add", 1 + len(v.d))
        self.dyn_vars.add(an)
        self.decl_deps.setdefault(an, set()).update(deps)
        self.state[an] = VS(True, frozenset(deps), None, None)

    def _open_branches(self):
        """Base terms of every branch-current unknown, emitted once at the top of the eval stream."""
        for (a, b), kind in self.branch_kind.items():
            br = self.tindex[f"I({a},{b})"]
            x = self.emit_val("r", f"VT({br})", {} if self.no_deriv else {br: 1.0})
            ia = self.tindex.get(a) if a != "0" else None
            ib = self.tindex.get(b) if b != "0" else None
            if ia is not None:
                self._acc("I", ia, 1.0, x)
            if ib is not None:
                self._acc("I", ib, -1.0, x)
            if kind == "V":
                self._acc("I", br, 1.0, self.force(self._probe(("probe", "V", [n for n in (a, b)]))))
            else:
                self._acc("I", br, 1.0, x)

    def contrib(self, st):
        acc, nodes, e = st[1], st[2], st[3]
        if len(nodes) == 1 and nodes[0] in self.mod.branches:
            nodes = list(self.mod.branches[nodes[0]])
        if acc not in ("I", "flow", "V", "potential"):
            raise VACompileError(f"{acc}() contributions are not supported")
        bro = self._branch_of(nodes)
        is_v = acc in ("V", "potential")
        if bro is not None and (bro[2] == "V") != is_v:
            # the reference resets the accumulator when the kind of contribution to a branch changes
            # (src/vasim.jl:149-154,172-177: a switch branch); not needed by any model on the sweep path
            raise VACompileError(f"branch ({', '.join(nodes)}) receives both I() and V() contributions (switch branch); unsupported")
        pos = self.tindex.get(nodes[0]) if nodes[0] not in ("0", "gnd") else None
        neg = None
        if len(nodes) > 1 and nodes[1] not in ("0", "gnd"):
            neg = self.tindex.get(nodes[1])
        if pos is not None and pos == neg:
            return
        e, noises = self._split_noise(e)
        if self.noise:
            self._noise_contrib(pos, neg, noises)
            return   # the noise variant outputs source powers only
        if e is None:
            return
        res, qs = self._split_ddt(e)
        saved = self.dynctl
        for c, _q in qs:
            if not self.is_static_expr(c):
                raise VACompileError("bias-dependent coefficient on ddt() is not charge-conserving; unsupported")
        for kind, expr in ([("I", res)] if res is not None else []) + [("Q", ("bin", "*", c, q)) for c, q in qs]:
            v = self.gd(expr)
            if isinstance(v, Lazy) and v.const is not None and v.const == 0:
                continue
            v = self.force(v)
            if v.typ == "i":
                v = Val("r", self._cast(v.c, "i", "r"), {})
            if bro is not None:
                # the branch has its own unknown: its row is  (V(a,b) | I_br) - sum(contributions) = 0, the KCL flows
                # were emitted by _open_branches; a reversed pair contributes with the opposite sign
                self._acc(kind, bro[0], -bro[1], v)
                continue
            for node, sign in ((pos, 1.0), (neg, -1.0)):
                if node is None:
                    continue
                self._acc(kind, node, sign, v)
        self.dynctl = saved

    # -- control flow --
    def _snapshot_state(self) -> Dict[str, VS]:
        return {k: VS(v.dyn, v.deps, v.slot, v.const) for k, v in self.state.items()}

    def if_stmt(self, cond, then, other):
        cv = self.gd(cond)
        if isinstance(cv, Lazy):
            c, t, const = self.gs(cv.e)
            if const is not None:   # compile-time branch
                taken = then if const else other
                if taken is not None:
                    self.stmt(taken)
                return
            if not self.dynctl:
                return self._if_static(c, then, other)
            cv = self.force(cv)
        return self._if_dynamic(cv, then, other)

    def _run_branch(self, st, start_state, dyn):
        save_S, save_E, save_state, save_dyn = self.S, self.E, self.state, self.dynctl
        self.S = save_S if dyn else []
        self.E = []
        self.state = start_state
        self.dynctl = dyn or save_dyn
        if st is not None:
            self.stmt(st)
        self.flush_ops(self.E)
        out = (self.S, self.E, self.state)
        self.S, self.E, self.state, self.dynctl = save_S, save_E, save_state, save_dyn
        return out

    def _merge(self, s0, sa, sb, Ea, Eb, Sa, Sb, shared_setup: bool):
        """Merge branch states; appends fix-ups to the branch code lists."""
        merged: Dict[str, VS] = {}
        for name in set(sa) | set(sb):
            a = sa.get(name) or VS()
            b = sb.get(name) or VS()
            if name not in self.types or name in self.dead_locals:
                continue
            if not a.dyn and not b.dyn:
                base = s0.get(name)
                if shared_setup:
                    # setup has no branch here: a snapshot taken in either arm ran unconditionally
                    slot = a.slot if a.slot is not None else b.slot
                else:
                    # only a snapshot that predates the `if` and survived both arms is still valid
                    slot = a.slot if (base is not None and a.slot == b.slot == base.slot) else None
                const = a.const if (a.const is not None and a.const == b.const) else None
                merged[name] = VS(False, frozenset(), slot, const)
                continue
            deps = frozenset((a.deps if a.dyn else frozenset()) | (b.deps if b.dyn else frozenset()))
            if name in self.forced_all:
                deps = frozenset(range(len(self.terms)))
            if self.types[name] != "r":
                deps = frozenset()   # integer variables carry no derivatives
            for br, E_, S_ in ((a, Ea, Sa), (b, Eb, Sb)):
                if br.dyn:
                    for kk in sorted(deps - br.deps):
                        E_.append(f"{name}__d{kk} = 0.0;")
                else:
                    # materialise the static value in this branch
                    save_S, save_E, save_state = self.S, self.E, self.state
                    self.S, self.E = S_, E_
                    self.state = {name: VS(False, frozenset(), br.slot, br.const)}
                    v = self.static_read(name)
                    self.S, self.E, self.state = save_S, save_E, save_state
                    E_.append(f"{name} = {v.c};")
                    if self.types[name] == "r":
                        for kk in sorted(deps):
                            E_.append(f"{name}__d{kk} = 0.0;")
            self.dyn_vars.add(name)
            if self.types[name] == "r":
                self.decl_deps.setdefault(name, set()).update(deps)
            merged[name] = VS(True, deps if self.types[name] == "r" else frozenset(), None, None)
        return merged

    def _if_static(self, c: str, then, other):
        s0 = self.state
        self.ntemp += 1
        cvar = f"c{self.ntemp}"
        self.S.append(f"const int {cvar} = ({c}) != 0;")
        marker = len(self.S)
        Sa, Ea, sa = self._run_branch(then, self._snapshot_state(), False)
        Sb, Eb, sb = self._run_branch(other, self._snapshot_state(), False)
        self.state = self._merge(s0, sa, sb, Ea, Eb, Sa, Sb, shared_setup=False)
        if Sa or Sb:
            self.S.append(f"if ({cvar}) {{")
            self.S.extend("    " + l for l in Sa)
            if Sb:
                self.S.append("} else {")
                self.S.extend("    " + l for l in Sb)
            self.S.append("}")
        if Ea or Eb:
            slot = self.nslot
            self.nslot += 1
            self.S.insert(marker, f"CACHE_ST({slot}, {cvar});")
            if Ea:
                self.E.append(f"if (CACHE_LD({slot}) != 0.0) {{")
                self.E.extend("    " + l for l in Ea)
                if Eb:
                    self.E.append("} else {")
                    self.E.extend("    " + l for l in Eb)
            else:
                self.E.append(f"if (CACHE_LD({slot}) == 0.0) {{")
                self.E.extend("    " + l for l in Eb)
            self.E.append("}")

    def _if_dynamic(self, cv: Val, then, other):
        s0 = self.state
        _, Ea, sa = self._run_branch(then, self._snapshot_state(), True)
        # slots created while generating `then` stay valid (setup has no branch here)
        start_b = self._snapshot_state()
        for name, a in sa.items():
            b = start_b.get(name)
            if b is not None and not a.dyn and not b.dyn and b.slot is None and a.slot is not None:
                b.slot = a.slot
        _, Eb, sb = self._run_branch(other, start_b, True)
        self.state = self._merge(s0, sa, sb, Ea, Eb, self.S, self.S, shared_setup=True)
        if Ea or Eb:
            if Ea:
                self.E.append(f"if ({cv.c}) {{")
                self.E.extend("    " + l for l in Ea)
                if Eb:
                    self.E.append("} else {")
                    self.E.extend("    " + l for l in Eb)
            else:
                self.E.append(f"if (!({cv.c})) {{")
                self.E.extend("    " + l for l in Eb)
            self.E.append("}")

    def loop(self, cond, body):
        asg: Set[str] = set()
        _assigned(body, asg, self.mod.functions)
        names = [self.resolve(n) for n in asg]
        # trial run on a scratch copy to classify the loop
        saved_ops = {k: list(v) for k, v in self.block_ops.items()}
        saved = (self.S, self.E, self.state, self.ntemp, self.nslot, self.ninl, dict(self.census),
                 {k: set(v) for k, v in self.decl_deps.items()}, set(self.dyn_vars), set(self.static_vars))
        self.S, self.E, self.state = [], [], self._snapshot_state()
        for cn in names:
            s = self.vs(cn)
            s.const, s.slot = None, None
        static_ok = not self.dynctl and self.is_static_expr(cond)
        if static_ok:
            self.stmt(body)
            static_ok = not self.E and self.is_static_expr(cond)
        (self.S, self.E, self.state, self.ntemp, self.nslot, self.ninl, self.census, self.decl_deps,
         self.dyn_vars, self.static_vars) = saved
        self.block_ops = saved_ops
        if static_ok:
            for cn in names:
                s = self.vs(cn)
                s.const, s.slot = None, None
                self.static_vars.add(cn)
            c, _, _ = self.gs(cond)
            outer = self.S
            self.S = []
            self.stmt(body)
            inner = self.S
            self.S = outer
            self.S.append(f"while ({c}) {{")
            self.S.extend("    " + l for l in inner)
            self.S.append("}")
            for cn in names:
                s = self.vs(cn)
                s.const, s.slot = None, None
            return
        # dynamic loop: every variable assigned inside carries all seeds
        allseeds = frozenset(range(len(self.terms)))
        for cn in names:
            v = self._load(cn)
            self.forced_all.add(cn)
            save = self.dynctl
            self.dynctl = True
            self._store(cn, v)
            self.dynctl = save
        save_dyn, outer = self.dynctl, self.E
        self.dynctl = True
        self.E = []
        cv = self.force(self.gd(cond))
        self.E.append(f"if (!({cv.c})) break;")
        self.stmt(body)
        self.flush_ops(self.E)
        inner = self.E
        self.E = outer
        self.dynctl = save_dyn
        self.E.append("for (;;) {")
        self.E.extend("    " + l for l in inner)
        self.E.append("}")
        for cn in names:
            if self.types[cn] == "r":
                self.state[cn] = VS(True, allseeds, None, None)

    # ---- driver ---------------------------------------------------------------------------------
    def compile(self) -> CompiledModel:
        mod = self.mod
        self.used_funcs: Set[str] = set()
        self.given_const: Dict[str, int] = {}
        # parameters: value if given, else default expression (may reference earlier parameters)
        for p in mod.params:
            cn = "p_" + p.name
            if p.type == "string":
                continue
            self.state[cn] = VS(False, frozenset(), None, None)
            ctyp = "i" if p.type == "integer" else "r"
            raw = f"PAR({p.index})"
            if self.const_params is not None:
                pu = p.name.upper()
                if pu in self.runtime_params:
                    self.given_const[p.name] = 1
                    self.S.append(f"{cn} = {self._cast(raw, 'r', ctyp)};")
                    self.static_vars.add(cn)
                    continue
                self.given_const[p.name] = 1 if pu in self.const_params else 0
                if pu in self.const_params:
                    v = float(self.const_params[pu])
                    cv = _vaconvert_int(v) if ctyp == "i" else v
                    self.state[cn] = VS(False, frozenset(), None, cv)
                    continue
                dflt, t, dc = self.gs(p.default)
                if dc is not None:
                    cv = (_vaconvert_int(dc) if t == "r" else int(dc)) if ctyp == "i" else float(dc)
                    self.state[cn] = VS(False, frozenset(), None, cv)
                else:
                    self.S.append(f"{cn} = {self._cast(dflt, t, ctyp)};")
                    self.static_vars.add(cn)
                continue
            dflt, t, _ = self.gs(p.default)
            self.S.append(f"{cn} = GIVEN({p.index}) ? {self._cast(raw, 'r', ctyp)} : {self._cast(dflt, t, ctyp)};")
            self.static_vars.add(cn)
        if self.const_params is not None:
            known = {p.name.upper() for p in mod.params}
            bad = [k for k in list(self.const_params) + list(self.runtime_params) if k not in known]
            if bad:
                raise VACompileError(f"module {mod.name} has no parameter(s) {bad}")
        body = ("block", None, list(mod.analog), {})
        if self.part in ("I", "Q"):
            body = self._strip_part(body)
        # (translational invariance holds for node voltages only, not for branch-current unknowns)
        self.drop_seed = None if (self.no_deriv or self.branch_terms) else self._pick_drop_seed(body)
        if not self.noise:
            self._open_branches()
        _NOISE_LIVE[0] = self.noise
        try:
            pruned, _ = prune_dead(body, set(), mod.functions)
        finally:
            _NOISE_LIVE[0] = False
        if pruned is not None:
            self.stmt(pruned)
        self.flush_ops(self.E)
        self._stage_cache()
        nt = len(self.terms)
        # outputs
        jrow, jcol = [], []
        out_lines = []
        for k in range(len(self.noise_sources) if self.noise else 0):
            out_lines.append(f"OUT_N({k}, accN_{k});")
            out_lines.append(f"OUT_NE({k}, {'accNE_%d' % k if ('accNE_%d' % k) in self.dyn_vars else '0.0'});")
        for kk in range(0 if self.noise else nt):
            si, sq = self.state.get(f"accI_{kk}"), self.state.get(f"accQ_{kk}")
            out_lines.append(f"OUT_I({kk}, {'accI_%d' % kk if si else '0.0'});")
            out_lines.append(f"OUT_Q({kk}, {'accQ_%d' % kk if sq else '0.0'});")
            di = si.deps if si else frozenset()
            dq = sq.deps if sq else frozenset()
            cols = set(di | dq)
            if self.drop_seed is not None and cols:
                cols.add(self.drop_seed)
            for ll in sorted(cols):
                if ll == self.drop_seed:
                    g = "-(" + " + ".join(f"accI_{kk}__d{m}" for m in sorted(di)) + ")" if di else "0.0"
                    cq = "-(" + " + ".join(f"accQ_{kk}__d{m}" for m in sorted(dq)) + ")" if dq else "0.0"
                else:
                    g = f"accI_{kk}__d{ll}" if ll in di else "0.0"
                    cq = f"accQ_{kk}__d{ll}" if ll in dq else "0.0"
                out_lines.append(f"OUT_J({len(jrow)}, {kk}, {ll}, {g}, {cq});")
                jrow.append(kk)
                jcol.append(ll)
        src = self._render(out_lines)
        linear = self._jacobian_is_static(out_lines)
        ptypes = [p.type for p in mod.params]
        defaults: Dict[str, float] = {}
        ce = _ConstEval(mod.functions)
        for p in mod.params:
            if p.type == "string":
                continue
            try:
                v = ce.expr(p.default, defaults)
                defaults[p.name] = float(v)
            except (_NotConst, TypeError, ValueError):
                pass
        return CompiledModel(self.name, mod.name, list(self.terms), len(mod.ports), [p.name for p in mod.params],
                             ptypes, self.nslot, jrow, jcol, src, len(self.E), len(self.S), dict(self.census),
                             param_defaults=defaults, branch_terms=list(self.branch_terms), linear=linear, nuni=self.nuni)

    def _jacobian_is_static(self, out_lines: List[str]) -> bool:
        """True when no derivative output depends on a terminal value: taint every eval-stream name assigned from an
        expression that reads VT(...) or a tainted name (one forward pass over the straight-line code; conditionals only
        add assignments), then look at the derivative arguments of OUT_J."""
        ident = re.compile(r"[A-Za-z_][A-Za-z0-9_]*")
        tainted: Set[str] = set()
        asg = re.compile(r"^\s*(?:const\s+)?(?:double|int)?\s*([A-Za-z_][A-Za-z0-9_]*)\s*(?:=|\+=|-=|\*=)\s*(.*);\s*$")
        cond = re.compile(r"^\s*(?:\}\s*else\s+)?if\s*\((.*)\)\s*\{\s*$")
        depth_taint: List[bool] = []      # inside a conditional whose condition is bias-dependent, every assignment is
        for l in self.E:
            m = cond.match(l)
            if m:
                t = "VT(" in m.group(1) or any(n in tainted for n in ident.findall(m.group(1)))
                if l.lstrip().startswith("}"):
                    if depth_taint:
                        depth_taint[-1] = depth_taint[-1] or t
                else:
                    depth_taint.append(t)
                continue
            if l.strip().startswith("} else"):
                continue
            if l.strip() == "}":
                if depth_taint:
                    depth_taint.pop()
                continue
            m = asg.match(l)
            if not m:
                continue
            name, rhs = m.group(1), m.group(2)
            if any(depth_taint) or "VT(" in rhs or any(n in tainted for n in ident.findall(rhs)):
                tainted.add(name)
        for l in out_lines:
            if l.startswith("OUT_J("):
                args = l[len("OUT_J("):].rsplit(")", 1)[0].split(",", 3)[3]
                if any(n in tainted for n in ident.findall(args)):
                    return False
        return True

    def _stage_cache(self):
        """Re-lays the cache out as a *stream* in the order the eval function consumes it.

        The eval code is long straight-line code whose cache reads are scattered through it; on the
        GPU every read is an HBM access of ~1 us latency, and the register allocator can only hoist
        a few of them.  Ordering the slots by first use turns the cache into a sequential stream that
        the CUDA prelude moves through a shared-memory ring with per-thread `cp.async` copies, a
        fixed number of chunks ahead of the consumer (see VA_CHUNK in csrc/va_prelude.h).

        Guarantee given to the prelude: in the top-level statement that follows VA_CHUNK(k) markers
        up to chunk k, every CACHE_LD(p) satisfies  need - VA_WINDOW < p <= need,  need = highest
        stream position used so far.  A value whose uses are further apart than the window gets a
        second stream position (the setup function stores it to all of them), so nothing is held in
        registers across the ring.  Markers are only placed between top-level statements: the
        sequence of copy groups is the same on every control path.
        """
        E, R, W = self.E, CACHE_CHUNK_ROWS, CACHE_WINDOW
        self.nuni = 0
        if UNIFORM_SLOTS:
            self._split_uniform_slots()
        ld = re.compile(r"CACHE_LD\((\d+)\)")
        blocks, depth, start = [], 0, 0
        for i, l in enumerate(E):
            if depth == 0:
                start = i
            depth += l.count("{") - l.count("}")
            if depth == 0:
                blocks.append((start, i + 1))
        stream: List[int] = []          # stream position -> old slot
        latest: Dict[int, int] = {}     # old slot -> its most recent stream position
        positions: Dict[int, List[int]] = {}
        out: List[str] = []
        for (a, b) in blocks:
            used: List[int] = []
            for l in E[a:b]:
                for m in ld.finditer(l):
                    k = int(m.group(1))
                    if k not in used:
                        used.append(k)
            direct = len(used) > W   # too many for the window: this statement reads HBM directly
            if used:
                while True:
                    new = [x for x in used if x not in latest]
                    need = len(stream) + len(new)
                    stale = [] if direct else [x for x in used if x in latest and latest[x] < need - W]
                    if not stale:
                        break
                    for x in stale:
                        del latest[x]
                for x in new:
                    if len(stream) % R == 0:
                        out.append(f"VA_CHUNK({len(stream) // R})")
                    latest[x] = len(stream)
                    positions.setdefault(x, []).append(len(stream))
                    stream.append(x)
            for l in E[a:b]:
                mac = "CACHE_LDG" if direct else "CACHE_LD"
                out.append(ld.sub(lambda m: f"{mac}({latest[int(m.group(1))]})", l) if used else l)
        self.E[:] = out
        st = re.compile(r"^(\s*)CACHE_ST\((\d+), (.*)\);$")
        S2: List[str] = []
        for l in self.S:
            m = st.match(l)
            if not m:
                S2.append(l)
                continue
            pos = positions.get(int(m.group(2)), [])
            if len(pos) == 1:
                S2.append(f"{m.group(1)}CACHE_ST({pos[0]}, {m.group(3)});")
            elif pos:
                S2.append(f"{m.group(1)}{{ const double cs_ = (double)({m.group(3)}); "
                          + " ".join(f"CACHE_ST({q}, cs_);" for q in pos) + " }")
        self.S[:] = S2
        self.nslot = len(stream)
        self.nchunk = (len(stream) + R - 1) // R

    def _split_uniform_slots(self):
        """Cache slots whose value does not depend on an instance parameter (PAR / GIVEN reads) -- only on the folded model
        card, the temperature and gmin -- are the same for every device of the model: they move from the per-instance
        cache rows (HBM, streamed by every evaluation) to one small table per model (CACHE_STU / CACHE_LDU).  BSIM-CMG on
        an ASAP7 card with L and NFIN as instance parameters: 84 of 256 stream positions.
        One forward taint pass over the setup stream; a name once tainted stays tainted (conservative), a store inside a
        conditional whose condition is tainted is tainted."""
        ident = re.compile(r"[A-Za-z_][A-Za-z0-9_]*")
        st = re.compile(r"CACHE_ST\((\d+), (.*)\);\s*$")
        asg = re.compile(r"^\s*(?:const\s+)?(?:double|int)?\s*([A-Za-z_][A-Za-z0-9_]*)\s*(?:=|\+=|-=|\*=|/=)\s*(.*);\s*$")
        ctl = re.compile(r"^\s*(?:\}\s*else\s+)?(?:if|while|for)\s*\((.*)\)\s*\{\s*$")
        tainted: Set[str] = set()
        stack: List[bool] = []
        dep: Dict[int, bool] = {}

        def is_t(text: str) -> bool:
            return "PAR(" in text or "GIVEN(" in text or any(n in tainted for n in ident.findall(text))

        for l in self.S:
            m = ctl.match(l)
            if m:
                t = is_t(m.group(1))
                if l.lstrip().startswith("}"):
                    if stack:
                        stack[-1] = stack[-1] or t
                else:
                    stack.append(t)
                continue
            ls = l.strip()
            if ls.startswith("} else"):
                continue
            if ls == "}":
                if stack:
                    stack.pop()
                continue
            if ls == "{":
                stack.append(False)
                continue
            m = st.search(l)
            if m:
                k = int(m.group(1))
                dep[k] = dep.get(k, False) or any(stack) or is_t(m.group(2))
                continue
            t = any(stack) or is_t(l)
            m = asg.match(l)
            if m and t:
                tainted.add(m.group(1))
            if t:
                for out in re.findall(r"&\s*([A-Za-z_][A-Za-z0-9_]*)", l):   # output arguments of analog functions
                    tainted.add(out)
            elif not m and "(" in l and "=" not in l:
                pass
        used = set(int(k) for l in self.E for k in re.findall(r"CACHE_LD\((\d+)\)", l))
        uni = sorted(k for k in used if k in dep and not dep[k])
        if not uni:
            return
        umap = {k: i for i, k in enumerate(uni)}
        self.nuni = len(uni)
        ld = re.compile(r"CACHE_LD\((\d+)\)")
        self.E[:] = [ld.sub(lambda m: f"CACHE_LDU({umap[int(m.group(1))]})" if int(m.group(1)) in umap else m.group(0), l) for l in self.E]
        sst = re.compile(r"CACHE_ST\((\d+), ")
        self.S[:] = [sst.sub(lambda m: f"CACHE_STU({umap[int(m.group(1))]}, " if int(m.group(1)) in umap else m.group(0), l) for l in self.S]

    def _pick_drop_seed(self, body) -> Optional[int]:
        """Translational invariance: when every probe is a difference V(a,b) of two terminals, every
        value v satisfies sum_k dv/dV_k = 0, so the derivative w.r.t. one terminal never has to be
        propagated -- it is recovered at the end as minus the sum of the others.  Returns the index of
        the terminal that appears in the most probes (the cheapest one to drop), or None."""
        counts: Dict[int, int] = {}
        ok = [True]

        def ex(e):
            k = e[0]
            if k == "probe":
                nodes = list(e[2])
                if len(nodes) == 1 and nodes[0] in self.mod.branches:
                    nodes = list(self.mod.branches[nodes[0]])
                if len(nodes) != 2 or any(n not in self.tindex for n in nodes):
                    ok[0] = False
                    return
                for n in nodes:
                    counts[self.tindex[n]] = counts.get(self.tindex[n], 0) + 1
            elif k == "bin":
                ex(e[2]); ex(e[3])
            elif k == "un":
                ex(e[2])
            elif k == "cond":
                for s_ in e[1:]:
                    ex(s_)
            elif k == "call":
                for a in e[2]:
                    ex(a)

        def st(s_):
            k = s_[0]
            if k == "assign":
                ex(s_[2])
            elif k == "contrib":
                ex(s_[3])
            elif k == "block":
                for x in s_[2]:
                    st(x)
            elif k == "if":
                ex(s_[1]); st(s_[2])
                if s_[3]:
                    st(s_[3])
            elif k == "case":
                ex(s_[1])
                for vals, x in s_[2]:
                    for v in vals:
                        ex(v)
                    st(x)
                if s_[3]:
                    st(s_[3])
            elif k == "for":
                st(s_[1]); ex(s_[2]); st(s_[3]); st(s_[4])
            elif k in ("while", "repeat"):
                ex(s_[1]); st(s_[2])

        st(body)
        for f in self.mod.functions.values():
            st(f.body)
        if not ok[0] or len(counts) < 2:
            return None
        return max(counts, key=lambda kk: (counts[kk], kk))

    def _c_function(self, f: Function) -> str:
        """value-only C translation of an analog function (used by the setup stream)"""
        sub = _Compiler.__new__(_Compiler)
        sub.__dict__.update(self.__dict__)
        sub.S, sub.E, sub.state, sub.scopes = [], [], {}, [dict((v, "l_" + v) for v in f.var_types)]
        sub.types = dict(self.types)
        sub.dynctl = False
        sub.static_vars = set()
        for v, t in f.var_types.items():
            sub.types["l_" + v] = "i" if t == "integer" else "r"
            sub.state["l_" + v] = VS(False, frozenset(), None, None)
        outs = {a for a, kind in f.args if kind != "input"}
        sig = []
        for a, kind in f.args:
            ct = "int" if f.var_types.get(a) == "integer" else "double"
            sig.append(f"{ct}{'*' if kind != 'input' else ''} {'o_' if kind != 'input' else 'l_'}{a}")
        lines = []
        rt = "int" if f.type == "integer" else "double"
        lines.append(f"VA_FN {rt} f_{self.name}_{f.name}({', '.join(sig)}) {{")
        for v, t in f.var_types.items():
            if v in [a for a, k in f.args if k == "input"]:
                continue
            init = f"*o_{v}" if v in outs and dict(f.args)[v] == "inout" else "0"
            lines.append(f"    {'int' if t == 'integer' else 'double'} l_{v} = {init};")
        sub.stmt(f.body)
        if sub.E:
            raise VACompileError(f"function {f.name} touches bias-dependent state")
        lines.extend("    " + l for l in sub.S)
        for v in outs:
            lines.append(f"    *o_{v} = l_{v};")
        lines.append(f"    return l_{f.name};")
        lines.append("}")
        self.used_funcs |= sub.used_funcs
        return "\n".join(lines)

    def _render(self, out_lines: List[str]) -> str:
        nt = len(self.terms)
        L: List[str] = []
        L.append(f"// generated by cedarsim.jl_b200.va.compiler from Verilog-A module '{self.mod.name}'")
        L.append(f"// terminals: {', '.join(self.terms)}   cache slots: {self.nslot}")
        L.append("#undef VA_CHUNK_ROWS\n#undef VA_WINDOW\n#undef VA_NCHUNK")
        L.append(f"#define VA_CHUNK_ROWS {CACHE_CHUNK_ROWS}\n#define VA_WINDOW {CACHE_WINDOW}\n#define VA_NCHUNK {self.nchunk}")
        L.append(f"#undef VA_NSTREAM\n#undef NUNI\n#define VA_NSTREAM {self.nslot}\n#define NUNI {max(1, self.nuni)}")
        # helper functions used by setup (emit in dependency-safe order: iterate to closure)
        done: Dict[str, str] = {}
        pending = set(self.used_funcs)
        while pending:
            fn = pending.pop()
            if fn in done:
                continue
            before = set(self.used_funcs)
            done[fn] = self._c_function(self.mod.functions[fn])
            pending |= (self.used_funcs - before) - set(done)
        order = [fn for fn in self.mod.functions if fn in done]  # declaration order
        self.defined_funcs = list(order)
        for fn in order:
            if fn not in self.skip_funcs:
                L.append(done[fn])
        V = "N" if self.noise else ("V" if self.no_deriv else "")
        L.append(f"VA_SETUP{V}_BEGIN({self.name})")
        for cn in sorted(self.static_vars):
            ct = "int" if self.types.get(cn) == "i" else "double"
            L.append(f"    {ct} {cn} = 0;")
        L.extend("    " + l for l in self.S)
        L.append(f"VA_SETUP{V}_END({self.name})")
        L.append(f"VA_EVAL{V}_BEGIN({self.name})")
        for cn in sorted(self.dyn_vars):
            if self.types.get(cn) == "i":
                L.append(f"    int {cn} = 0;")
            else:
                L.append(f"    double {cn} = 0.0;")
                for kk in sorted(self.decl_deps.get(cn, ())):
                    L.append(f"    double {cn}__d{kk} = 0.0;")
        L.extend("    " + l for l in self.E)
        L.extend("    " + l for l in out_lines)
        L.append(f"VA_EVAL{V}_END({self.name})")
        return "\n".join(L) + "\n"


TIME_PORT = "time__"   # hidden port of a module that reads $abstime / $realtime (see _lower_abstime)


def _lower_abstime(mod: Module) -> Module:
    """`$abstime` / `$realtime` inside a module: the device code takes no time argument, so simulation time enters the way
    it does for behavioural sources (netlist._behavioural) -- as the voltage of a net driven by V(t) = t.  A module that
    reads the time gets one extra, hidden port TIME_PORT behind its own ports and every `$abstime` becomes the probe
    V(time__); the netlist flattener connects the port to the circuit's time net (0 V in the DC solve, like `$abstime`
    there).  Returns `mod` itself when it does not read the time."""
    import copy
    hit = [False]

    def walk(x):
        if isinstance(x, tuple):
            if len(x) >= 2 and x[0] == "call" and x[1] in ("$abstime", "$realtime"):
                hit[0] = True
                return ("probe", "V", [TIME_PORT])
            return tuple(walk(y) for y in x)
        if isinstance(x, list):
            return [walk(y) for y in x]
        return x

    analog = walk(mod.analog)
    if not hit[0]:
        return mod
    for f in mod.functions.values():
        probe = [False]

        def has(x):
            if isinstance(x, tuple) and len(x) >= 2 and x[0] == "call" and x[1] in ("$abstime", "$realtime"):
                probe[0] = True
            if isinstance(x, (tuple, list)):
                for y in x:
                    has(y)
        has(f.body)
        if probe[0]:
            raise VACompileError(f"$abstime inside analog function {f.name} is not supported")
    if TIME_PORT in mod.nets:
        raise VACompileError(f"net name {TIME_PORT} is reserved")
    m2 = copy.copy(mod)
    m2.analog = analog
    m2.ports = list(mod.ports) + [TIME_PORT]
    m2.nets = list(m2.ports) + [n for n in mod.nets if n not in mod.ports]
    return m2


def _lower_switch_branches(mod: Module) -> Module:
    """A branch that receives BOTH `V(a,b) <+` and `I(a,b) <+` contributions (a switch branch): the kind of the LAST
    contribution executed decides what the branch is, and a contribution of the other kind discards what was
    accumulated before (the reference resets its accumulator when the kind changes, src/vasim.jl:149-154, 172-177).
    Lowered here, before code generation, to constructs the generator already has: per switch branch an integer mode
    (0 none, 1 voltage, 2 current) and two real accumulators are ordinary module variables, every contribution becomes

        if (mode != KIND) begin acc_KIND = 0; mode = KIND; end   acc_KIND = acc_KIND + (expr);

    and one contribution at the end of the analog block closes the branch through its own current unknown:

        I(a,b) <+ I(a,b) - ((mode == 1) ? V(a,b) - acc_v : I(a,b) - acc_i);

    i.e. the row of the branch current reads  V(a,b) - acc_v = 0  in voltage mode and  I(a,b) - acc_i = 0  otherwise
    (I = 0 when no contribution ran, the LRM's unassigned branch).  The mode may depend on run-time parameters -- one
    sweep can hold points of both kinds.  ddt() inside a contribution to a switch branch is not supported."""
    import copy

    def nodes_of(nodes):
        nodes = list(nodes)
        if len(nodes) == 1 and nodes[0] in mod.branches:
            nodes = list(mod.branches[nodes[0]])
        g = lambda n: "0" if n in ("0", "gnd") else n
        return (g(nodes[0]), g(nodes[1]) if len(nodes) > 1 else "0")

    kinds: Dict[Tuple[str, str], Set[str]] = {}

    def scan(x):
        if isinstance(x, tuple):
            if len(x) >= 4 and x[0] == "contrib" and x[1] in ("V", "potential", "I", "flow"):
                a, b = nodes_of(x[2])
                key = (a, b) if (a, b) in kinds or (b, a) not in kinds else (b, a)
                kinds.setdefault(key, set()).add("V" if x[1] in ("V", "potential") else "I")
            for y in x:
                scan(y)
        elif isinstance(x, list):
            for y in x:
                scan(y)
        elif isinstance(x, dict):
            for y in x.values():
                scan(y)

    scan(list(mod.analog))
    switches = {k: i for i, k in enumerate(sorted(k for k, v in kinds.items() if len(v) == 2))}
    if not switches:
        return mod

    def has_ddt(e) -> bool:
        if isinstance(e, tuple):
            if len(e) >= 2 and e[0] == "call" and e[1] in ("ddt", "idt"):
                return True
            return any(has_ddt(y) for y in e)
        if isinstance(e, list):
            return any(has_ddt(y) for y in e)
        return False

    num = lambda v: ("num", v, isinstance(v, int))

    def rewrite(x):
        if isinstance(x, tuple):
            if len(x) >= 4 and x[0] == "contrib" and x[1] in ("V", "potential", "I", "flow"):
                a, b = nodes_of(x[2])
                key, sign = ((a, b), 1.0) if (a, b) in switches else (((b, a), -1.0) if (b, a) in switches else (None, 1.0))
                if key is not None:
                    if has_ddt(x[3]):
                        raise VACompileError(f"ddt() in a contribution to the switch branch ({a}, {b}) is not supported")
                    k = switches[key]
                    is_v = x[1] in ("V", "potential")
                    mode, acc = f"sw{k}__m", f"sw{k}__{'v' if is_v else 'i'}"
                    e = x[3] if sign > 0 else ("un", "-", x[3])
                    kind = 1 if is_v else 2
                    return ("block", None, [
                        ("if", ("bin", "!=", ("var", mode), num(kind)),
                         ("block", None, [("assign", acc, num(0.0)), ("assign", mode, num(kind))], {}), None),
                        ("assign", acc, ("bin", "+", ("var", acc), e))], {})
            return tuple(rewrite(y) for y in x)
        if isinstance(x, list):
            return [rewrite(y) for y in x]
        if isinstance(x, dict):
            return {kk: rewrite(v) for kk, v in x.items()}
        return x

    m2 = copy.copy(mod)
    m2.var_types = dict(mod.var_types)
    head, tail = [], []
    for (a, b), k in switches.items():
        for nm, ty in ((f"sw{k}__m", "integer"), (f"sw{k}__v", "real"), (f"sw{k}__i", "real")):
            if nm in m2.var_types:
                raise VACompileError(f"variable name {nm} is reserved")
            m2.var_types[nm] = ty
        head += [("assign", f"sw{k}__m", num(0)), ("assign", f"sw{k}__v", num(0.0)), ("assign", f"sw{k}__i", num(0.0))]
        nodes = [a] if b == "0" else [a, b]
        ibr, vbr = ("probe", "I", list(nodes)), ("probe", "V", list(nodes))
        row = ("cond", ("bin", "==", ("var", f"sw{k}__m"), num(1)), ("bin", "-", vbr, ("var", f"sw{k}__v")),
               ("bin", "-", ibr, ("var", f"sw{k}__i")))
        tail.append(("contrib", "I", list(nodes), ("bin", "-", ibr, row)))
    m2.analog = head + rewrite(list(mod.analog)) + tail
    return m2


def compile_module(mod: Module, name: Optional[str] = None, const_params=None, runtime_params=None,
                   probe_branches=()) -> CompiledModel:
    mod = _lower_switch_branches(_lower_abstime(mod))
    full = _Compiler(mod, name or mod.name, const_params, runtime_params, probe_branches=probe_branches)
    cm = full.compile()
    try:
        vc = _Compiler(mod, name or mod.name, const_params, runtime_params, no_deriv=True, skip_funcs=full.defined_funcs,
                       probe_branches=probe_branches)
        cv = vc.compile()
        cm.source_v, cm.ncache_v, cm.nuni_v = cv.source, cv.ncache, cv.nuni
    except VACompileError:
        pass
    if _module_has_noise(mod):
        nc = _Compiler(mod, name or mod.name, const_params, runtime_params, noise=True, skip_funcs=full.defined_funcs,
                       probe_branches=probe_branches)
        cn = nc.compile()
        cm.source_n, cm.ncache_n, cm.noise_sources, cm.nuni_n = cn.source, cn.ncache, list(nc.noise_sources), cn.nuni
    cm.gen_version = GEN_VERSION
    return cm


def _module_has_noise(mod: Module) -> bool:
    def walk(x) -> bool:
        if isinstance(x, tuple):
            if len(x) >= 2 and x[0] == "call" and x[1] in NOISE_FUNCS:
                return True
            return any(walk(y) for y in x)
        if isinstance(x, (list, dict)):
            return any(walk(y) for y in (x.values() if isinstance(x, dict) else x))
        return False
    return walk(list(mod.analog))


def compile_va_file(path: str, module: Optional[str] = None, name: Optional[str] = None,
                    include_paths: Sequence[str] = (), defines=None, suppress_defines=(),
                    const_params=None, runtime_params=None, probe_branches=()) -> CompiledModel:
    import time
    t0 = time.perf_counter()
    pp = Preprocessor(include_paths, defines, suppress_defines)
    text = pp.process_file(path)
    cm = compile_va_text(text, module, name, preprocessed=True, const_params=const_params,
                         runtime_params=runtime_params, probe_branches=probe_branches)
    cm.codegen_seconds = time.perf_counter() - t0
    return cm


def compile_va_text(text: str, module: Optional[str] = None, name: Optional[str] = None,
                    preprocessed: bool = False, defines=None, const_params=None, runtime_params=None,
                    probe_branches=()) -> CompiledModel:
    if not preprocessed:
        text = Preprocessor((), defines).process_text(text)
    mods = parse(text)
    if not mods:
        raise VACompileError("no module found")
    mod = mods[-1] if module is None else next(m for m in mods if m.name == module)
    return compile_module(mod, name, const_params, runtime_params, probe_branches=probe_branches)
