"""SPICE / Spectre numbers and parameter expressions.

Numbers follow the reference's magnitude table (src/spectre.jl:383-455): case-insensitive SI
suffixes t g meg k m mil u n p f a, trailing unit letters ignored.  The reference scales with
decimal-exact `Dec64` arithmetic so that `0.22u === 0.22e-6` (test/basic.jl:609-638); we get
the same result by folding the suffix into the decimal exponent before the float conversion.

Expressions (`'...'`, `{...}`, `.param` right-hand sides) are parsed to a small AST and can be
evaluated with scalars or numpy arrays (one value per sweep point), which is how swept
parameters propagate to device parameters on the host.
"""
from __future__ import annotations

import math
import re
from typing import Callable, Dict, Mapping, Union

import numpy as np

_SUFFIX_EXP = {"t": 12, "g": 9, "meg": 6, "k": 3, "m": -3, "u": -6, "n": -9, "p": -12, "f": -15, "a": -18}
_NUM = re.compile(r"^([+-]?(?:\d+\.?\d*|\.\d+))(?:[eE]([+-]?\d+))?(.*)$")


class ExprError(Exception):
    pass


def parse_number(tok: str) -> float:
    s = tok.strip().lower().rstrip(",")
    m = _NUM.match(s)
    if not m:
        raise ExprError(f"not a number: {tok!r}")
    mant, exp, rest = m.group(1), int(m.group(2) or 0), m.group(3).lstrip("_")
    if rest.startswith("mil"):
        return float(f"{mant}e{exp}") * 25.4e-6
    if rest.startswith("meg"):
        exp += 6
    elif rest and rest[0] in _SUFFIX_EXP and not rest.startswith("am"):
        # `1Amp` is one ampere, not one atto-"mp" (SPICE lexer.jl:377-395 excludes ('A','M'))
        exp += _SUFFIX_EXP[rest[0]]
    return float(f"{mant}e{exp}")


_TOK = re.compile(r"\s*(?:(\d+\.?\d*(?:[eE][+-]?\d+)?[A-Za-z_]*|\.\d+(?:[eE][+-]?\d+)?[A-Za-z_]*)"
                  r"|([A-Za-z_$][A-Za-z0-9_.$]*)|(\*\*|==|!=|<=|>=|&&|\|\||~\^|\^~|<<|>>|[-+*/^(),<>?:!&|~]))")

_FUNCS: Dict[str, Callable] = {
    "sqrt": np.sqrt, "exp": np.exp, "ln": np.log, "log": np.log, "log10": np.log10, "abs": np.abs,
    "sin": np.sin, "cos": np.cos, "tan": np.tan, "atan": np.arctan, "arctan": np.arctan, "sinh": np.sinh,
    "cosh": np.cosh, "tanh": np.tanh, "asinh": np.arcsinh, "acosh": np.arccosh, "atanh": np.arctanh,
    "min": np.minimum, "max": np.maximum, "pow": np.power, "pwr": np.power,
    "floor": np.floor, "ceil": np.ceil, "int": np.trunc, "nint": np.rint,
    "agauss": lambda nom, avar, sigma: nom,   # rng disabled in the reference (src/spectre_env.jl:178-187)
    "gauss": lambda nom, rvar, sigma: nom,
}
_CONSTS = {"pi": math.pi, "e": math.e, "true": 1.0, "false": 0.0,
           "m_1_pi": 1 / math.pi,   # the one Spectre constant the reference defines (src/spectre_env.jl:142)
           "m_pi": math.pi, "m_e": math.e, "m_two_pi": 2 * math.pi, "m_sqrt2": math.sqrt(2.0)}


def tokenize(text: str):
    pos, out = 0, []
    text = text.strip()
    while pos < len(text):
        m = _TOK.match(text, pos)
        if not m or m.end() == pos:
            raise ExprError(f"bad expression near {text[pos:pos + 20]!r}")
        pos = m.end()
        if m.group(1) is not None:
            out.append(("num", parse_number(m.group(1))))
        elif m.group(2) is not None:
            out.append(("id", m.group(2).lower()))
        else:
            out.append(("op", m.group(3)))
    return out


class _P:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("eof", None)

    def eat(self, op=None):
        tok = self.peek()
        if op is not None and tok != ("op", op):
            raise ExprError(f"expected {op!r}, got {tok!r}")
        self.i += 1
        return tok

    def ternary(self):
        c = self.binary(0)
        if self.peek() == ("op", "?"):
            self.eat()
            a = self.ternary()
            self.eat(":")
            b = self.ternary()
            return ("cond", c, a, b)
        return c

    # C-like precedence; the bitwise operators are those of Spectre expressions (test/spectre_expr.jl:11: `1&2~^3`)
    LEVELS = [("||",), ("&&",), ("|",), ("~^", "^~"), ("&",), ("==", "!="), ("<", "<=", ">", ">="), ("<<", ">>"),
              ("+", "-"), ("*", "/"), ]

    def binary(self, lvl):
        if lvl == len(self.LEVELS):
            return self.unary()
        lhs = self.binary(lvl + 1)
        while self.peek()[0] == "op" and self.peek()[1] in self.LEVELS[lvl]:
            op = self.eat()[1]
            lhs = ("bin", op, lhs, self.binary(lvl + 1))
        return lhs

    def unary(self):
        tok = self.peek()
        if tok == ("op", "-"):
            self.eat()
            return ("neg", self.unary())
        if tok == ("op", "+"):
            self.eat()
            return self.unary()
        if tok == ("op", "!"):
            self.eat()
            return ("not", self.unary())
        if tok == ("op", "~"):
            self.eat()
            return ("inv", self.unary())
        return self.power()

    def power(self):
        base = self.atom()
        if self.peek() in (("op", "**"), ("op", "^")):
            self.eat()
            return ("bin", "**", base, self.unary())
        return base

    def atom(self):
        kind, val = self.peek()
        if kind == "num":
            self.eat()
            return ("num", val)
        if kind == "id":
            self.eat()
            if self.peek() == ("op", "("):
                self.eat()
                args = []
                if self.peek() != ("op", ")"):
                    args.append(self.ternary())
                    while self.peek() == ("op", ","):
                        self.eat()
                        args.append(self.ternary())
                self.eat(")")
                return ("call", val, args)
            return ("var", val)
        if (kind, val) == ("op", "("):
            self.eat()
            e = self.ternary()
            self.eat(")")
            return e
        raise ExprError(f"unexpected token {val!r}")


def parse_expr(text: str):
    text = text.strip()
    if len(text) >= 2 and text[0] in "'{\"" and text[-1] in "'}\"":
        text = text[1:-1]
    p = _P(tokenize(text))
    e = p.ternary()
    if p.peek()[0] != "eof":
        raise ExprError(f"trailing input in expression {text!r}")
    return e


def free_vars(e, out=None):
    out = set() if out is None else out
    k = e[0]
    if k == "var":
        if e[1] not in _CONSTS:
            out.add(e[1])
    elif k in ("neg", "not", "inv"):
        free_vars(e[1], out)
    elif k == "bin":
        free_vars(e[2], out)
        free_vars(e[3], out)
    elif k == "cond":
        for s in e[1:]:
            free_vars(s, out)
    elif k == "call":
        for a in e[2]:
            free_vars(a, out)
    return out


Number = Union[float, np.ndarray]


def evaluate(e, env: Mapping[str, Number]) -> Number:
    k = e[0]
    if k == "num":
        return e[1]
    if k == "var":
        if e[1] in env:
            return env[e[1]]
        if e[1] in _CONSTS:
            return _CONSTS[e[1]]
        raise ExprError(f"undefined parameter {e[1]!r}")
    if k == "neg":
        return -evaluate(e[1], env)
    if k == "not":
        v = evaluate(e[1], env)
        return np.where(v != 0, 0.0, 1.0) if isinstance(v, np.ndarray) else float(not v)
    if k == "inv":
        v = evaluate(e[1], env)
        return (~np.asarray(v).astype(np.int64)).astype(float) if isinstance(v, np.ndarray) else float(~int(v))
    if k == "bin":
        a, b = evaluate(e[2], env), evaluate(e[3], env)
        op = e[1]
        if op in ("&", "|", "~^", "^~", "<<", ">>"):   # integer operators
            ia, ib = np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64)
            r = {"&": lambda: ia & ib, "|": lambda: ia | ib, "~^": lambda: ~(ia ^ ib), "^~": lambda: ~(ia ^ ib),
                 "<<": lambda: ia << ib, ">>": lambda: ia >> ib}[op]()
            return r.astype(float) if r.ndim > 0 else float(r)
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            return a / b
        if op == "**":
            return a ** b
        r = {"==": np.equal, "!=": np.not_equal, "<": np.less, "<=": np.less_equal, ">": np.greater,
             ">=": np.greater_equal, "&&": lambda x, y: np.logical_and(x != 0, y != 0),
             "||": lambda x, y: np.logical_or(x != 0, y != 0)}[op](a, b)
        return r.astype(float) if isinstance(r, np.ndarray) and r.ndim > 0 else float(r)
    if k == "cond":
        c = evaluate(e[1], env)
        a, b = evaluate(e[2], env), evaluate(e[3], env)
        if isinstance(c, np.ndarray):
            return np.where(c != 0, a, b)
        return a if c else b
    if k == "call":
        fn = _FUNCS.get(e[1])
        if fn is None:
            raise ExprError(f"unknown function {e[1]!r}")
        r = fn(*[evaluate(a, env) for a in e[2]])
        return r if isinstance(r, np.ndarray) and r.ndim > 0 else float(r)
    raise ExprError(f"bad expression node {k}")
