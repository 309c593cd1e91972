"""Reader for model cards (`model NAME MASTER k=v ...` in Spectre syntax, `.model NAME TYPE k=v ...`
in SPICE syntax) -- only what the sweep path needs to resolve transistor parameters.

Parameter-name handling follows the reference's `.model` lowering (src/spectre.jl:558-566,
:630-641): names are upper-cased, `LEVEL`/`VERSION` are dropped, `type=n|p` becomes
`DEVTYPE=1|0` for BSIM-CMG.  SPICE cards of type `nmos` / `pmos` pick their device family from
`level` (+ `version`) as `spice_select_device` does (src/spectre.jl:596-617): 14 / 54 -> bsim4,
17 / 72 -> bsimcmg107.  Cards named `<base>.<N>` are bins of the binned model `<base>`
(src/spectre.jl:675,716-720); `find_bin` restates the half-open window test of
src/spectre.jl:1160-1170.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict

from .expr import evaluate, parse_expr, parse_number


@dataclass
class ModelCard:
    name: str
    master: str
    params: Dict[str, float] = field(default_factory=dict)
    exprs: Dict[str, str] = field(default_factory=dict)   # values that need the netlist's parameter scope


class NoBinException(ValueError):
    """No bin of a binned model covers (scale*l, scale*w) -- the reference's NoBinExpection (src/spectre.jl:1152-1157)."""


_BIN_RX = re.compile(r"^(.*)\.([0-9]+)$")   # binning_rx, src/spectre.jl:675


def bins_of(cards: Dict[str, "ModelCard"], base: str):
    """Bins `<base>.<N>` in card order (the reference keeps definition order, src/spectre.jl:716-720)."""
    base = base.lower()
    out = []
    for name, c in cards.items():
        m = _BIN_RX.match(name)
        if m and m.group(1) == base:
            out.append(c)
    return out


def find_bin(cards: Dict[str, "ModelCard"], base: str, l: float, w: float, scale: float = 1.0) -> "ModelCard":
    """LMIN <= scale*l < LMAX and WMIN <= scale*w < WMAX, first match wins (src/spectre.jl:1160-1170)."""
    bins = bins_of(cards, base)
    if not bins:
        raise KeyError(base)
    ls, ws = scale * l, scale * w
    for c in bins:
        try:
            lmin, lmax, wmin, wmax = (c.params[k] for k in ("LMIN", "LMAX", "WMIN", "WMAX"))
        except KeyError as e:
            raise NoBinException(f"bin {c.name!r} of binned model {base!r} lacks {e.args[0]}") from None
        if lmin <= ls < lmax and wmin <= ws < wmax:
            return c
    raise NoBinException(f"NoBinExpection: no bin for BinnedModel {base} of size (l={ls}, w={ws}).")


def _logical_lines(text: str):
    cur = None
    for raw in text.splitlines():
        line = raw.split("//")[0].rstrip()
        s = line.strip()
        if not s or s.startswith("*"):
            continue
        if s.startswith("+"):
            if cur is not None:
                cur += " " + s[1:]
            continue
        if cur is not None:
            yield cur
        cur = s
    if cur is not None:
        yield cur


def parse_model_cards(text: str) -> Dict[str, ModelCard]:
    cards: Dict[str, ModelCard] = {}
    for line in _logical_lines(text):
        m = re.match(r"\.?model\s+(\S+)\s+(\S+)\s*(.*)$", line, re.I)
        if not m:
            continue
        name, master, rest = m.group(1).lower(), m.group(2).lower(), m.group(3)
        rest = rest.replace("(", " ").replace(")", " ")
        card = ModelCard(name, master)
        level = version = None
        mos = master if master in ("nmos", "pmos") else None
        for k, v in re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*=\s*('[^']*'|\{[^}]*\}|[^\s=]+)", rest):
            ku = k.upper()
            if ku == "LEVEL":
                level = int(float(v))
                continue
            if ku == "VERSION":
                version = float(v)
                continue
            if ku == "TYPE":
                mos = "pmos" if v.lower().strip("'\"").startswith("p") else "nmos"
                continue
            if v[0] in "'{":   # quoted expression: constant-fold now, else defer to the netlist's parameter scope
                inner = v.strip("'{}")
                try:
                    card.params[ku] = float(evaluate(parse_expr(inner), {}))
                except Exception:
                    card.exprs[ku] = inner
                continue
            try:
                card.params[ku] = parse_number(v)
            except Exception:
                card.exprs[ku] = v
        if master in ("nmos", "pmos") and level is not None:
            if level in (17, 72) and version in (None, 107.0):
                card.master = "bsimcmg107"
            elif level in (14, 54):
                card.master = "bsim4"
        if mos is not None:   # devtype_param, src/spectre.jl:630-641
            if card.master == "bsim4":
                card.params["TYPE"] = -1.0 if mos == "pmos" else 1.0
            else:
                card.params["DEVTYPE"] = 0.0 if mos == "pmos" else 1.0
        cards[name] = card
    return cards


def load_model_cards(path: str) -> Dict[str, ModelCard]:
    with open(path, "r", errors="replace") as f:
        return parse_model_cards(f.read())
