"""Spectre-language subset reader: the same `Netlist` structure as the SPICE reader (netlist.py), so that the
flattener, the sweep API and the engine see no difference.

What the sweep path and the reference's own Spectre-syntax tests need (test/basic.jl:168-205, :265-278, :353-367):
instances `name (n1 n2 ...) master k=v ...` of the primitives `resistor capacitor inductor vsource isource vcvs vccs`,
of subcircuits, of model cards and of Verilog-A modules; `subckt ... ends` with `parameters`; `model`; `include`,
`ahdl_include`; `type=pwl wave=[...]`, `type=sine`, `type=pulse` sources; `\\` continuations, `//` and `*` comments.
Names are kept as the flattener wants them (lower case; results are addressed case-insensitively).
Behavioural sources (`bsource v=` / `i=` over V() probes, parameters and `$time`) become the same cards the SPICE
reader makes of B sources; analyses are skipped and `simulator lang=spice` sections are outside the subset and raise.

Device / parameter names follow the reference's lowering (src/spectre.jl:999-1071: `r`, `c`, `l`, `gain`, `gm`;
source parameters src/spectre_env.jl:144-176).
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Sequence

from .modelcard import parse_model_cards
from .netlist import Card, Netlist, NetlistError, Subckt, _include, _resolve

_PRIMS = {"resistor": "r", "capacitor": "c", "inductor": "l", "vsource": "v", "isource": "i", "vcvs": "e", "vccs": "g"}
_TOK = re.compile(r"\[[^\]]*\]|\"[^\"]*\"|'[^']*'|[()=]|[^\s()=\[\]]+")


def _strip_comment(raw: str) -> str:
    """`// ...` comments, but not inside a quoted string (include "jlpkg://...")"""
    quote = None
    for i, ch in enumerate(raw):
        if quote:
            if ch == quote:
                quote = None
        elif ch in "\"'":
            quote = ch
        elif ch == "/" and raw[i:i + 2] == "//":
            return raw[:i]
    return raw


def _logical_lines(text: str) -> List[str]:
    out: List[str] = []
    cur = ""
    for raw in text.splitlines():
        line = _strip_comment(raw).rstrip()
        if not cur and line.lstrip().startswith("*"):
            continue
        if line.endswith("\\"):
            cur += line[:-1] + " "
            continue
        cur += line
        if cur.count("[") > cur.count("]"):   # a vector that runs over the line end
            cur += " "
            continue
        if cur.strip():
            out.append(cur.strip())
        cur = ""
    if cur.strip():
        out.append(cur.strip())
    return out


def _kv(toks: List[str]):
    pos, kv, i = [], {}, 0
    while i < len(toks):
        if i + 2 < len(toks) and toks[i + 1] == "=":
            kv[toks[i].lower()] = toks[i + 2]
            i += 3
        else:
            pos.append(toks[i])
            i += 1
    return pos, kv


_ASSIGN = re.compile(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?<![=!<>])=(?!=)")


def _assignments(text: str) -> Dict[str, str]:
    """`a=1 b = x*2 c = p ? 1 : 2` -> {a: '1', b: 'x*2', c: 'p ? 1 : 2'}: a value runs up to the next `name =`"""
    ms = list(_ASSIGN.finditer(text))
    out: Dict[str, str] = {}
    for k, m in enumerate(ms):
        end = ms[k + 1].start() if k + 1 < len(ms) else len(text)
        out[m.group(1).lower()] = text[m.end():end].strip()
    return out


def _source(kv: Dict[str, str]) -> dict:
    """vsource / isource parameters -> the source dict of the SPICE reader (dc / ac / tran)."""
    src = {"dc": kv.get("dc"), "ac": kv.get("mag"), "tran": None}
    typ = kv.get("type", "dc").lower().strip("\"'")
    if typ == "pwl":
        wave = kv.get("wave", "[]").strip("[]").replace(",", " ").split()
        src["tran"] = ("pwl", wave)
    elif typ in ("sine", "sin"):
        # SIN(vo va freq td theta phase), src/spectre_env.jl:169-176
        src["tran"] = ("sin", [kv.get("sinedc", kv.get("dc", "0")), kv.get("ampl", "0"), kv.get("freq", "0"), kv.get("delay", "0"),
                               kv.get("damp", "0"), kv.get("sinephase", "0")])
    elif typ == "pulse":
        # PULSE(v1 v2 td tr tf pw per), src/spectre_env.jl:153-166
        src["tran"] = ("pulse", [kv.get("val0", "0"), kv.get("val1", "0"), kv.get("delay", "0"), kv.get("rise", "0"),
                                 kv.get("fall", "0"), kv.get("width", "0"), kv.get("period", "0")])
    elif typ != "dc":
        raise NetlistError(f"source type {typ!r} is outside the Spectre subset")
    return src


def parse_spectre(text: str, path: Optional[str] = None, include_dirs: Optional[Sequence[str]] = None,
                  _nl: Optional[Netlist] = None) -> Netlist:
    nl = _nl or Netlist()
    if include_dirs:
        nl.include_dirs.extend(include_dirs)
    base = os.path.dirname(path) if path else "."
    stack: List[Subckt] = [nl.top]
    for line in _logical_lines(text):
        toks = _TOK.findall(line)
        head = toks[0].lower()
        cur = stack[-1]
        if head == "simulator":
            if re.search(r"lang\s*=\s*spice", line, re.I):
                raise NetlistError("`simulator lang=spice` sections are outside the Spectre subset; use the SPICE reader")
            continue
        if head == "parameters":
            kv = _assignments(line[len(toks[0]):])
            (cur.local_params if cur is not nl.top else cur.params).update(kv)
            continue
        if head in ("subckt", "inline"):
            t = [x for x in toks[1:] if x not in ("(", ")", "subckt")]
            sub = Subckt(t[0].lower(), [p.lower() for p in t[1:]])
            cur.subckts[sub.name] = sub
            stack.append(sub)
            continue
        if head == "ends":
            if len(stack) > 1:
                stack.pop()
            continue
        if head == "model":
            nl.cards.update(parse_model_cards(line))
            continue
        if head in ("include", "ahdl_include"):
            fname = toks[1].strip("\"'")
            if head == "ahdl_include" or fname.endswith((".va", ".vams")):
                vpath = _resolve(nl, fname, base)
                with open(vpath, "r", errors="replace") as f:
                    for mname in re.findall(r"^\s*module\s+([A-Za-z_][A-Za-z0-9_$]*)", f.read(), re.M):
                        nl.va_modules[mname.lower()] = vpath
            else:
                _include(nl, fname, base, None)
            continue
        if head in ("global", "save", "ic", "nodeset", "options") or (len(toks) > 1 and toks[1].lower() in ("options", "tran", "dc", "ac", "noise", "info")):
            continue   # analyses and controls: the sweep API decides what is run
        # ---- instance: name (nodes) master k=v ...   (parentheses optional)
        name = toks[0].lower()
        rest = toks[1:]
        if rest and rest[0] == "(":
            close = rest.index(")")
            nodes, rest = rest[1:close], rest[close + 1:]
            pos, kv = _kv(rest)
            master = pos[0] if pos else ""
        else:
            pos, kv = _kv(rest)
            nodes, master = pos[:-1], pos[-1] if pos else ""
        if master:   # values may be whole expressions (`r=(p1+p2)/p3`): re-read the assignments from the raw text
            tail = line[line.index(master, len(toks[0])) + len(master):]
            kv = _assignments(tail) or kv
        nodes = [n.lower() for n in nodes]
        m = master.lower()
        if m in _PRIMS:
            kind = _PRIMS[m]
            if kind in "vi":
                cur.cards.append(Card(kind, name, nodes[:2], source=_source(kv)))
            elif kind in "eg":
                val = kv.pop("gain", None) if kind == "e" else kv.pop("gm", kv.pop("gain", None))
                cur.cards.append(Card(kind, name, nodes[:4], None, kv, val))
            else:
                cur.cards.append(Card(kind, name, nodes[:2], None, kv, None))
        elif m == "bsource":   # `B5 (0 5) bsource v=$time*V(3)`: same behavioural card the SPICE reader makes of `B5 0 5 v=...`
            key = next((k for k in kv if k.lower() in ("v", "i")), None)
            if key is None:
                raise NetlistError(f"{toks[0]}: bsource needs v=<expr> or i=<expr>")
            cur.cards.append(Card("b", name, nodes[:2], None, {key.lower(): str(kv[key])}))
        elif m in nl.cards or any(k.startswith(m + ".") for k in nl.cards):
            card = nl.cards.get(m)
            kind = "m" if (card is None or card.master.startswith(("bsim", "nmos", "pmos"))) else card.master[:1]
            cur.cards.append(Card(kind, name, nodes, m, kv))
        else:
            cur.cards.append(Card("x", name, nodes, m, kv))   # subcircuit or Verilog-A module, resolved by the flattener
    return nl


def parse_spectre_file(path: str, include_dirs: Optional[Sequence[str]] = None) -> Netlist:
    with open(path, "r", errors="replace") as f:
        return parse_spectre(f.read(), path, include_dirs=list(include_dirs or []) + [os.path.dirname(path)])
