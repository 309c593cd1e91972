"""Registry of compiled device models and model cards.

The Verilog-A sources and foundry cards are *inputs* that live outside this repository (the
reference vendors BSIM-CMG 107 under VerilogAParser.jl/cmc_models/bsimcmg107 and the ASAP7 cards
under SpectreNetlistParser.jl/test/examples/7nm_TT.scs).  They are compiled where they lie; the
generated artefacts are cached under `cedarsim.jl_b200/_gen/` (git-ignored, travels to the GPU
box with the built .so files), so no GPU-side code ever reads /root/reference.
"""
from __future__ import annotations

import dataclasses
import json
import os
from typing import Dict, Optional

from .modelcard import ModelCard, load_model_cards
from .va.compiler import GEN_VERSION, CompiledModel, compile_va_file

_HERE = os.path.dirname(os.path.abspath(__file__))
GEN_DIR = os.path.join(_HERE, os.environ.get("CB_GEN_DIR", "_gen"))   # CB_GEN_DIR: experiment variants keep their own cache
REFERENCE_ROOT = os.environ.get("CEDAR_REFERENCE_ROOT", "/root/reference")
BSIMCMG_VA = os.path.join(REFERENCE_ROOT, "VerilogAParser.jl/cmc_models/bsimcmg107/bsimcmg.va")
ASAP7_CARDS = os.path.join(REFERENCE_ROOT, "SpectreNetlistParser.jl/test/examples/7nm_TT.scs")

_cache: Dict[str, CompiledModel] = {}


def save_model(cm: CompiledModel, path: str):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(f"{path}.{os.getpid()}.tmp", "w") as f:
        json.dump(dataclasses.asdict(cm), f)
    os.replace(f"{path}.{os.getpid()}.tmp", path)


def load_model(path: str) -> CompiledModel:
    with open(path) as f:
        return CompiledModel(**json.load(f))


def _count_executed_ops(cm: CompiledModel, probe) -> list:
    """FP64 operation count of one evaluation on the path executed at a representative bias (the
    algorithmic flops per device evaluation used by bench.py's roofline).  exp/log/sqrt/pow/div
    count 1 each, so fractions of peak derived from it are conservative."""
    import numpy as np
    from .va.build import build_host
    hm = build_host(cm, count_ops=True)
    cache = hm.run_setup(probe["params"])
    return list(hm.executed_ops(cache, np.asarray(probe["v"], dtype=float)))


def compiled_model(name: str, va_path: Optional[str] = None, rebuild: bool = False, probe=None, **kw) -> CompiledModel:
    """Compiled model by name, from the _gen cache or by compiling `va_path`."""
    if name in _cache and not rebuild:
        return _cache[name]
    path = os.path.join(GEN_DIR, f"{name}.model.json")
    cm = None
    if os.path.exists(path) and not rebuild:
        cm = load_model(path)
        if cm.gen_version != GEN_VERSION and va_path is not None and os.path.exists(va_path):
            cm = None   # written by another version of the generator and the source is at hand: regenerate
    if cm is None:
        if va_path is None or not os.path.exists(va_path):
            raise FileNotFoundError(f"no cached model {path} and no Verilog-A source {va_path!r}; run build() where the source exists")
        cm = compile_va_file(va_path, name=name, **kw)
        if probe is not None:
            cm.exec_ops = _count_executed_ops(cm, probe)
        save_model(cm, path)
    _cache[name] = cm
    return cm


def bsimcmg107(rebuild: bool = False) -> CompiledModel:
    # __OPINFO__ only computes operating-point report variables; suppressing it does not change
    # any contribution (bsimcmg_main.va:62).
    return compiled_model("bsimcmg107", BSIMCMG_VA, rebuild, suppress_defines=["__OPINFO__"])


def bsimcmg107_card(card: str, runtime=("L", "NFIN"), rebuild: bool = False) -> CompiledModel:
    """BSIM-CMG 107 specialised on one ASAP7 model card: every card parameter is folded at code
    generation time (circuit-specialised CUDA C, as the north star asks); only `runtime`
    instance parameters are read per device / per sweep point."""
    params = asap7_cards()[card].params
    name = f"bsimcmg107_{card}"
    vdd = 0.7 if params.get("DEVTYPE", 1.0) else -0.7
    probe = {"params": {"L": 21e-9, "NFIN": 3.0}, "v": [0.5 * vdd, 0.5 * vdd, 0.0, 0.0, 0.5 * vdd, 0.0]}
    return compiled_model(name, BSIMCMG_VA, rebuild, probe=probe, suppress_defines=["__OPINFO__"], const_params=params,
                          runtime_params=list(runtime))


def specialized_model(card: ModelCard, runtime=("L", "NFIN"), rebuild: bool = False) -> CompiledModel:
    """BSIM-CMG 107 specialised on an arbitrary model card (netlist `.model` or included card file)."""
    import hashlib
    if not card.master.startswith("bsimcmg"):
        raise ValueError(f"no Verilog-A source for device family {card.master!r}")
    runtime = tuple(sorted(k.upper() for k in runtime))
    const = {k: v for k, v in card.params.items() if k.upper() not in runtime}
    key = hashlib.sha1(repr((sorted(const.items()), runtime)).encode()).hexdigest()[:10]
    try:
        ref = asap7_cards().get(card.name)
    except FileNotFoundError:
        ref = None
    if ref is not None and ref.params == card.params and runtime == ("L", "NFIN"):
        return bsimcmg107_card(card.name, rebuild=rebuild)
    vdd = 0.7 if const.get("DEVTYPE", 1.0) else -0.7
    probe = {"params": {k: (21e-9 if k == "L" else 3.0 if k == "NFIN" else card.params.get(k, 0.0)) for k in runtime},
             "v": [0.5 * vdd, 0.5 * vdd, 0.0, 0.0, 0.5 * vdd, 0.0]}
    ident = "".join(ch if ch.isalnum() else "_" for ch in card.name)   # bins are named `<base>.<N>`
    return compiled_model(f"bsimcmg107_{ident}_{key}", BSIMCMG_VA, rebuild, probe=probe, suppress_defines=["__OPINFO__"],
                          const_params=const, runtime_params=list(runtime))


def param_default(cm: CompiledModel, name: str) -> float:
    for k, v in cm.param_defaults.items():
        if k.upper() == name.upper():
            return v
    raise KeyError(f"parameter {name} of {cm.name} has no constant default")


def asap7_cards(rebuild: bool = False) -> Dict[str, ModelCard]:
    path = os.path.join(GEN_DIR, "asap7_cards.json")
    if os.path.exists(path) and not rebuild:
        with open(path) as f:
            return {k: ModelCard(**v) for k, v in json.load(f).items()}
    if not os.path.exists(ASAP7_CARDS):
        raise FileNotFoundError(f"no cached cards {path} and no card file {ASAP7_CARDS}")
    cards = load_model_cards(ASAP7_CARDS)
    os.makedirs(GEN_DIR, exist_ok=True)
    with open(path, "w") as f:
        json.dump({k: dataclasses.asdict(v) for k, v in cards.items()}, f)
    return cards


def available() -> bool:
    return (os.path.exists(os.path.join(GEN_DIR, "bsimcmg107_nmos_lvt.model.json")) or os.path.exists(BSIMCMG_VA)) and \
           (os.path.exists(os.path.join(GEN_DIR, "asap7_cards.json")) or os.path.exists(ASAP7_CARDS))
