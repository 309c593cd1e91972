"""Verilog-A lexer and recursive-descent parser for the analog subset compact models use.

Produces a small tuple-based AST (see the node list below).  Behavioural counterpart of the
reference's VerilogAParser.jl (forms.jl / parse.jl), scoped to what src/vasim.jl lowers:
modules, parameters, variables, analog functions, analog blocks with if/case/for/while/
repeat, assignments, contributions, system tasks.

Expression nodes
    ('num', value, is_int)            ('str', text)              ('var', name)
    ('bin', op, a, b)                 ('un', op, a)              ('cond', c, a, b)
    ('call', name, [args])            ('probe', 'V'|'I'|..., [node names])
Statement nodes
    ('assign', name, expr)            ('contrib', access, [nodes], expr)
    ('if', cond, then, else|None)     ('block', label|None, [stmts], [decls])
    ('case', expr, [([values], stmt)], default|None)
    ('for', init, cond, step, body)   ('while', cond, body)      ('repeat', count, body)
    ('task', name, [args])            ('nop',)
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


class VAParseError(Exception):
    pass


SCALE = {"T": 1e12, "G": 1e9, "M": 1e6, "K": 1e3, "k": 1e3, "m": 1e-3, "u": 1e-6, "n": 1e-9,
         "p": 1e-12, "f": 1e-15, "a": 1e-18}

_TOKEN = re.compile(r"""
    (?P<ws>\s+)
  | (?P<attr>\(\*(?!\s*\)).*?\*\))
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+|[TGMKkmunpfa](?![A-Za-z0-9_]))?)
  | (?P<id>[A-Za-z_$][A-Za-z0-9_$]*)
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<op><\+|\*\*|==|!=|<=|>=|&&|\|\||<<|>>|~\^|\^~|[-+*/%<>!~&|^?:=(){}\[\],;.@\#'])
""", re.X | re.S)


@dataclass
class Tok:
    kind: str
    text: str
    pos: int


def tokenize(text: str) -> List[Tok]:
    toks: List[Tok] = []
    pos, n = 0, len(text)
    while pos < n:
        m = _TOKEN.match(text, pos)
        if not m:
            line = text.count("\n", 0, pos) + 1
            raise VAParseError(f"unexpected character {text[pos]!r} at line {line}")
        kind = m.lastgroup
        if kind not in ("ws", "attr"):
            toks.append(Tok(kind, m.group(0), pos))
        pos = m.end()
    toks.append(Tok("eof", "", n))
    return toks


@dataclass
class Param:
    name: str
    type: str  # 'real' | 'integer' | 'string'
    default: tuple
    index: int = -1


@dataclass
class Function:
    name: str
    type: str
    args: List[Tuple[str, str]]  # (name, 'input'|'output'|'inout') in declaration order
    var_types: Dict[str, str]
    body: tuple


@dataclass
class Module:
    name: str
    ports: List[str]
    nets: List[str] = field(default_factory=list)  # declaration order, ports first
    params: List[Param] = field(default_factory=list)
    var_types: Dict[str, str] = field(default_factory=dict)
    functions: Dict[str, Function] = field(default_factory=dict)
    branches: Dict[str, Tuple[str, str]] = field(default_factory=dict)
    analog: List[tuple] = field(default_factory=list)

    @property
    def internal_nodes(self):
        return [n for n in self.nets if n not in self.ports]


_BINPREC = [  # low -> high
    ("||",), ("&&",), ("|",), ("^", "~^", "^~"), ("&",), ("==", "!="), ("<", "<=", ">", ">="),
    ("<<", ">>"), ("+", "-"), ("*", "/", "%"), ("**",),
]
_ACCESS = {"V", "I", "Temp", "Pwr", "potential", "flow"}
_TYPES = ("real", "integer", "string")


class Parser:
    def __init__(self, text: str):
        self.text = text
        self.toks = tokenize(text)
        self.i = 0
        self.disciplines = {"electrical", "thermal", "voltage", "current", "magnetic", "kinematic",
                            "kinematic_v", "rotational", "rotational_omega"}

    # ---- token helpers -----------------------------------------------------------
    @property
    def tok(self) -> Tok:
        return self.toks[self.i]

    def peek(self, k=1) -> Tok:
        return self.toks[min(self.i + k, len(self.toks) - 1)]

    def error(self, msg):
        line = self.text.count("\n", 0, self.tok.pos) + 1
        raise VAParseError(f"{msg} at line {line} near {self.tok.text!r}")

    def at(self, text) -> bool:
        return self.tok.text == text and self.tok.kind in ("op", "id")

    def accept(self, text) -> bool:
        if self.at(text):
            self.i += 1
            return True
        return False

    def expect(self, text) -> Tok:
        if not self.at(text):
            self.error(f"expected {text!r}")
        self.i += 1
        return self.toks[self.i - 1]

    def ident(self) -> str:
        if self.tok.kind != "id":
            self.error("expected identifier")
        self.i += 1
        return self.toks[self.i - 1].text

    # ---- top level ---------------------------------------------------------------
    def parse(self) -> List[Module]:
        mods = []
        while self.tok.kind != "eof":
            if self.at("module") or self.at("macromodule"):
                mods.append(self.module())
            elif self.at("discipline"):
                self.i += 1
                self.disciplines.add(self.ident())
                while not self.accept("enddiscipline"):
                    self.i += 1
            elif self.at("nature"):
                while not self.accept("endnature"):
                    self.i += 1
            else:
                self.error("expected module")
        return mods

    def module(self) -> Module:
        self.i += 1
        name = self.ident()
        ports = []
        if self.accept("("):
            while not self.accept(")"):
                ports.append(self.ident())
                self.accept(",")
        self.expect(";")
        mod = Module(name, ports, nets=list(ports))
        while not self.accept("endmodule"):
            self.module_item(mod)
        for k, p in enumerate(mod.params):
            p.index = k
        return mod

    def _add_nets(self, mod, names):
        for n in names:
            if n not in mod.nets:
                mod.nets.append(n)

    def id_list(self) -> List[str]:
        names = [self.ident()]
        while self.accept(","):
            names.append(self.ident())
        return names

    def module_item(self, mod: Module):
        t = self.tok.text
        if t in ("inout", "input", "output"):
            self.i += 1
            if self.tok.text in self.disciplines:
                self.i += 1
            self.id_list()
            self.expect(";")
        elif t in self.disciplines or t == "ground":
            self.i += 1
            self._add_nets(mod, self.id_list())
            self.expect(";")
        elif t == "branch":
            self.i += 1
            self.expect("(")
            a = self.ident()
            b = self.ident() if self.accept(",") else "0"
            self.expect(")")
            for nm in self.id_list():
                mod.branches[nm] = (a, b)
            self.expect(";")
        elif t in ("parameter", "localparam", "aliasparam"):
            self.parameter(mod)
        elif t in _TYPES:
            self.var_decl(mod.var_types)
        elif t == "analog":
            self.i += 1
            if self.at("function"):
                fn = self.function()
                mod.functions[fn.name] = fn
            else:
                self.accept("initial")
                mod.analog.append(self.statement())
        else:
            self.error("unsupported module item")

    def parameter(self, mod: Module):
        kind = self.ident()
        if kind == "aliasparam":
            while not self.accept(";"):
                self.i += 1
            return
        ptype = "real"
        if self.tok.text in _TYPES:
            ptype = self.ident()
        first = True
        while first or self.accept(","):
            first = False
            name = self.ident()
            self.expect("=")
            default = self.expr()
            # optional ranges: from [a:b) / exclude x
            while self.at("from") or self.at("exclude"):
                if self.accept("from"):
                    if self.tok.text in ("[", "("):
                        self.i += 1
                        self.range_bound()
                        self.expect(":")
                        self.range_bound()
                        if self.tok.text not in ("]", ")"):
                            self.error("bad range")
                        self.i += 1
                    else:  # from '{' list '}'
                        self.expect("{")
                        while not self.accept("}"):
                            self.i += 1
                else:
                    self.i += 1
                    if self.tok.text in ("[", "("):
                        self.i += 1
                        self.range_bound()
                        self.expect(":")
                        self.range_bound()
                        self.i += 1
                    else:
                        self.expr()
            if ptype == "real" and default[0] == "str":
                ptype = "string"
            mod.params.append(Param(name, ptype, default))
        self.expect(";")

    def range_bound(self):
        if self.accept("-"):
            if self.accept("inf"):
                return
            self.i -= 1
        if self.accept("+"):
            pass
        if self.accept("inf"):
            return
        self.expr()

    def var_decl(self, types: Dict[str, str]):
        vtype = self.ident()
        first = True
        while first or self.accept(","):
            first = False
            name = self.ident()
            if self.accept("["):  # arrays are not used by compact models we target
                self.error("array variables are not supported")
            if self.accept("="):
                self.expr()  # initialisers are ignored (variables start at 0)
            types[name] = vtype
        self.expect(";")

    def function(self) -> Function:
        self.expect("function")
        ftype = "real"
        if self.tok.text in ("real", "integer"):
            ftype = self.ident()
        name = self.ident()
        self.expect(";")
        args: List[Tuple[str, str]] = []
        vtypes: Dict[str, str] = {}
        while self.tok.text in ("input", "output", "inout", "real", "integer"):
            t = self.tok.text
            if t in ("input", "output", "inout"):
                self.i += 1
                for nm in self.id_list():
                    args.append((nm, t))
                self.expect(";")
            else:
                self.var_decl(vtypes)
        body = self.statement()
        self.expect("endfunction")
        vtypes.setdefault(name, ftype)
        for a, _ in args:
            vtypes.setdefault(a, "real")
        return Function(name, ftype, args, vtypes, body)

    # ---- statements ----------------------------------------------------------------
    def statement(self) -> tuple:
        t = self.tok
        if t.text == ";":
            self.i += 1
            return ("nop",)
        if t.text == "begin":
            self.i += 1
            label = None
            decls: Dict[str, str] = {}
            if self.accept(":"):
                label = self.ident()
            stmts = []
            while not self.accept("end"):
                if self.tok.text in ("real", "integer") and self.peek().kind == "id":
                    self.var_decl(decls)
                else:
                    stmts.append(self.statement())
            return ("block", label, stmts, decls)
        if t.text == "if":
            self.i += 1
            self.expect("(")
            cond = self.expr()
            self.expect(")")
            then = self.statement()
            other = self.statement() if self.accept("else") else None
            return ("if", cond, then, other)
        if t.text == "case":
            self.i += 1
            self.expect("(")
            sel = self.expr()
            self.expect(")")
            items, default = [], None
            while not self.accept("endcase"):
                if self.accept("default"):
                    self.accept(":")
                    default = self.statement()
                else:
                    vals = [self.expr()]
                    while self.accept(","):
                        vals.append(self.expr())
                    self.expect(":")
                    items.append((vals, self.statement()))
            return ("case", sel, items, default)
        if t.text == "for":
            self.i += 1
            self.expect("(")
            init = self.assignment()
            self.expect(";")
            cond = self.expr()
            self.expect(";")
            step = self.assignment()
            self.expect(")")
            return ("for", init, cond, step, self.statement())
        if t.text == "while":
            self.i += 1
            self.expect("(")
            cond = self.expr()
            self.expect(")")
            return ("while", cond, self.statement())
        if t.text == "repeat":
            self.i += 1
            self.expect("(")
            cnt = self.expr()
            self.expect(")")
            return ("repeat", cnt, self.statement())
        if t.text == "@":  # event control: @(initial_step) stmt -- the statement is kept
            self.i += 1
            self.expect("(")
            depth = 1
            while depth:
                if self.at("("):
                    depth += 1
                elif self.at(")"):
                    depth -= 1
                self.i += 1
            return self.statement()
        if t.kind == "id" and t.text.startswith("$"):
            name = self.ident()
            args = []
            if self.accept("("):
                while not self.accept(")"):
                    args.append(self.expr())
                    self.accept(",")
            self.expect(";")
            return ("task", name, args)
        if t.kind == "id":
            # contribution:  access ( nodes ) <+ expr ;
            if self.peek().text == "(":
                save = self.i
                name = self.ident()
                self.expect("(")
                nodes = []
                ok = True
                while not self.accept(")"):
                    if self.tok.kind != "id":
                        ok = False
                        break
                    nodes.append(self.ident())
                    self.accept(",")
                if ok and self.accept("<+"):
                    e = self.expr()
                    self.expect(";")
                    return ("contrib", name, nodes, e)
                self.i = save
            st = self.assignment()
            self.expect(";")
            return st
        self.error("unsupported statement")

    def assignment(self) -> tuple:
        name = self.ident()
        self.expect("=")
        return ("assign", name, self.expr())

    # ---- expressions -----------------------------------------------------------------
    def expr(self) -> tuple:
        c = self.binary(0)
        if self.accept("?"):
            a = self.expr()
            self.expect(":")
            b = self.expr()
            return ("cond", c, a, b)
        return c

    def binary(self, level: int) -> tuple:
        if level == len(_BINPREC):
            return self.unary()
        ops = _BINPREC[level]
        if ops == ("**",):  # right associative
            lhs = self.binary(level + 1)
            if self.tok.kind == "op" and self.tok.text == "**":
                self.i += 1
                return ("bin", "**", lhs, self.binary(level))
            return lhs
        lhs = self.binary(level + 1)
        while self.tok.kind == "op" and self.tok.text in ops:
            op = self.tok.text
            self.i += 1
            lhs = ("bin", op, lhs, self.binary(level + 1))
        return lhs

    def unary(self) -> tuple:
        t = self.tok
        if t.kind == "op" and t.text in ("-", "+", "!", "~"):
            self.i += 1
            operand = self.unary_pow()
            if t.text == "+":
                return operand
            if t.text == "-" and operand[0] == "num":
                return ("num", -operand[1], operand[2])
            return ("un", t.text, operand)
        return self.primary()

    def unary_pow(self) -> tuple:
        # unary operators bind tighter than '**' per the LRM precedence table
        return self.unary()

    def primary(self) -> tuple:
        t = self.tok
        if t.kind == "num":
            self.i += 1
            txt = t.text
            if txt[-1] in SCALE and not txt[-1].isdigit():
                return ("num", float(txt[:-1]) * SCALE[txt[-1]], False)
            if re.fullmatch(r"\d+", txt):
                return ("num", int(txt), True)
            return ("num", float(txt), False)
        if t.kind == "str":
            self.i += 1
            return ("str", t.text[1:-1])
        if t.text == "(":
            self.i += 1
            e = self.expr()
            self.expect(")")
            return e
        if t.kind == "id":
            name = self.ident()
            if self.at("("):
                self.i += 1
                args = []
                while not self.accept(")"):
                    args.append(self.expr())
                    self.accept(",")
                if name in _ACCESS:
                    nodes = []
                    for a in args:
                        if a[0] != "var":
                            self.error("access function arguments must be nets or branches")
                        nodes.append(a[1])
                    return ("probe", name, nodes)
                return ("call", name, args)
            if name == "inf":
                return ("num", float("inf"), False)
            if name.startswith("$"):
                return ("call", name, [])
            return ("var", name)
        self.error("unexpected token in expression")


def parse(text: str) -> List[Module]:
    return Parser(text).parse()
