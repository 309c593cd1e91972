"""Reader for model cards (`model NAME MASTER k=v ...` in Spectre syntax, `.model NAME TYPE k=v ...`
in SPICE syntax) -- only what the sweep path needs to resolve transistor parameters.

Parameter-name handling follows the reference's `.model` lowering (src/spectre.jl:558-566,
:630-641): names are upper-cased, `LEVEL`/`VERSION` are dropped, `type=n|p` becomes
`DEVTYPE=1|0` for BSIM-CMG.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict

from .expr import parse_number


@dataclass
class ModelCard:
    name: str
    master: str
    params: Dict[str, float] = field(default_factory=dict)


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
        for k, v in re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*=\s*([^\s=]+)", rest):
            ku = k.upper()
            if ku in ("LEVEL", "VERSION"):
                continue
            if ku == "TYPE":
                card.params["DEVTYPE"] = 0.0 if v.lower().startswith("p") else 1.0
                continue
            card.params[ku] = parse_number(v)
        cards[name] = card
    return cards


def load_model_cards(path: str) -> Dict[str, ModelCard]:
    with open(path, "r", errors="replace") as f:
        return parse_model_cards(f.read())
