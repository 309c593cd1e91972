"""Programmatic builders of the benchmark / parity workloads.

BSIM4 and the GF180 / sky130 cards are not in the reference tree, so the
transistor-level workloads use BSIM-CMG 107 with the ASAP7 cards, same topologies:
  * fet_iv        : config 4, single nFET DC I-V sweep over (vg, vd)
  * inverter      : configs 1/2, CMOS inverter driven by a PWL source, sweep vdd x nfin x l
  * dff           : config 3, the 30-FET D flip-flop of test/DFF/*.ngspice (topology restated
                    here with ASAP7 devices, 0.7 V supply and time axis scaled to FinFET speeds)
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from . import models
from .flat import FlatCircuit, VAModelShape, Wave, W_PWL, W_DC, shape_of


def _model(fc: FlatCircuit, card: str, host: bool, used: list) -> int:
    """Register BSIM-CMG specialised on `card`; with host=True the shape also carries the addresses
    of the host-compiled setup/eval functions (needed only by the CPU oracle in tests)."""
    cm = models.bsimcmg107_card(card)
    if cm not in used:
        used.append(cm)
    if host:
        from .va.build import build_host
        # host="fast:<cpu tag>": the -O3 -march=native build of bench.py's cpu_fast baseline arm (not the checker's)
        shape = build_host(cm, fast_tag=host.split(":", 1)[1] if isinstance(host, str) and host.startswith("fast:") else None).shape()
    else:
        shape = shape_of(cm)
    return fc.va_model(shape)


def fet_iv(host: bool = False, card: str = "nmos_lvt", nfin: float = 3.0, l: float = 20e-9):
    fc, used = FlatCircuit(), []
    m = _model(fc, card, host, used)
    fc.vsource("Vg", "g", "0", fc.param("vg.dc"))
    fc.vsource("Vd", "d", "0", fc.param("vd.dc"))
    fc.va_instance("m1", m, ["d", "g", "0", "0"], dict(L=l, NFIN=nfin))
    fc.set_outputs(["vd.i", "m1.di", "m1.si"])
    return fc, used


def inverter(host: bool = False, vdd=0.7, sweep: bool = True, cload: float = 1e-15, tscale: float = 1.0):
    """CMOS inverter, topology of benchmarks/benchmark_common.jl:82-106 (Xneg/Xpos, VVDD, VD PWL,
    CQ on the output).  Swept columns: vvdd.dc, xneg.nfin, xneg.l (FinFET analogue of W, L)."""
    fc, used = FlatCircuit(), []
    mn = _model(fc, "nmos_lvt", host, used)
    mp = _model(fc, "pmos_lvt", host, used)
    vd = fc.param("vvdd.dc") if sweep else vdd
    nfin = fc.param("xneg.nfin") if sweep else 3.0
    ln = fc.param("xneg.l") if sweep else 21e-9
    fc.vsource("VVDD", "vdd", "0", vd)
    fc.vsource("VVSS", "vss", "0", 0.0)
    # 8-point PWL on D with 10 ns edges (scaled), amplitude follows the nominal supply
    t = np.array([0, 50, 60, 150, 160, 250, 260, 400]) * 1e-9 * tscale
    y = [0.0, 0.0, vdd, vdd, 0.0, 0.0, vdd, vdd]
    fc.vsource("VD", "d", "0", Wave(W_PWL, t=list(t), y=y))
    fc.va_instance("xneg", mn, ["q", "d", "vss", "vss"], dict(L=ln, NFIN=nfin))
    fc.va_instance("xpos", mp, ["q", "d", "vdd", "vdd"], dict(L=21e-9, NFIN=3.0))
    fc.capacitor("CQ", "q", "0", cload)
    fc.set_outputs(["q", "d"])
    return fc, used


# (name, drain, gate, source, bulk, kind, nfin): topology of
# test/DFF/gf180mcu_fd_sc_mcu7t5v0__dffnq_4.ngspice:4-58, widths mapped to fin counts
_DFF_FETS: List[Tuple[str, str, str, str, str, str, int]] = [
    ("x_tn10", "vss", "d", "d_neg", "vpw", "n", 2), ("x_tp10", "vdd", "d", "d_neg", "vnw", "p", 3),
    ("x_tn11", "d_neg", "cki", "d_neg_clked", "vpw", "n", 2), ("x_tp11", "d_neg_clked", "ncki", "d_neg", "vnw", "p", 3),
    ("x_tn15", "q_internal", "d_neg_clked", "vss", "vpw", "n", 2), ("x_tp15", "q_internal", "d_neg_clked", "vdd", "vnw", "p", 3),
    ("x_tn0", "d_neg_clked", "ncki", "net11", "vpw", "n", 2), ("x_tp0", "net4", "cki", "d_neg_clked", "vnw", "p", 3),
    ("x_tn1", "vss", "q_internal", "net11", "vpw", "n", 2), ("x_tp1", "vdd", "q_internal", "net4", "vnw", "p", 3),
    ("x_tn2", "net0", "ncki", "q_internal", "vpw", "n", 2), ("x_tp7", "net0", "cki", "q_internal", "vnw", "p", 3),
    ("x_tn3", "net7", "cki", "net0", "vpw", "n", 2), ("x_tp6", "net7", "ncki", "net0", "vnw", "p", 3),
    ("x_tn5", "q_neg", "net0", "vss", "vpw", "n", 5), ("x_tp3", "q_neg", "net0", "vdd", "vnw", "p", 6),
    ("x_tn4", "vss", "q_neg", "net7", "vpw", "n", 5), ("x_tp2", "vdd", "q_neg", "net7", "vnw", "p", 6),
    ("x_tn6_7", "q", "q_neg", "vss", "vpw", "n", 4), ("x_tn6", "q", "q_neg", "vss", "vpw", "n", 4),
    ("x_tn6_7_61", "q", "q_neg", "vss", "vpw", "n", 4), ("x_tn6_49", "q", "q_neg", "vss", "vpw", "n", 4),
    ("x_tp4_13", "q", "q_neg", "vdd", "vnw", "p", 6), ("x_tp4", "q", "q_neg", "vdd", "vnw", "p", 6),
    ("x_tp4_13_64", "q", "q_neg", "vdd", "vnw", "p", 6), ("x_tp4_55", "q", "q_neg", "vdd", "vnw", "p", 6),
    ("x_tn9", "ncki", "clkn", "vss", "vpw", "n", 3), ("x_tp9", "ncki", "clkn", "vdd", "vnw", "p", 5),
    ("x_tn16", "cki", "ncki", "vss", "vpw", "n", 3), ("x_tp16", "cki", "ncki", "vdd", "vnw", "p", 5),
]
DFF_FET_NAMES = [f[0] for f in _DFF_FETS]


def dff(host: bool = False, vdd: float = 0.7, sweep: bool = True, tscale: float = 1.0, cq: float = 2e-15):
    """30-FET DFF + 7 V sources + load cap, deck shape of test/DFF/DFF_cap_all.cir.  With `sweep`
    every FET gets two swept columns `<inst>.l` and `<inst>.nfin` (P = 60, Monte-Carlo draws are
    supplied by the caller as a TandemSweep of pre-drawn values)."""
    fc, used = FlatCircuit(), []
    mn = _model(fc, "nmos_lvt", host, used)
    mp = _model(fc, "pmos_lvt", host, used)
    fc.vsource("VVDD", "vdd", "0", vdd)
    fc.vsource("VVSS", "vss", "0", 0.0)
    for name, d, g, s, b, kind, nfin in _DFF_FETS:
        ln = fc.param(f"{name}.l") if sweep else 21e-9
        nf = fc.param(f"{name}.nfin") if sweep else float(nfin)
        fc.va_instance(name, mn if kind == "n" else mp, [d, g, s, b], dict(L=ln, NFIN=nf))
    fc.capacitor("CQ", "q_tmp", "0", cq)
    fc.vsource("VQ", "q", "q_tmp", 0.0)
    fc.vsource("VNW", "vnw", "vdd", 0.0)
    fc.vsource("VPW", "vpw", "vss", 0.0)
    ps = 1e-12 * tscale
    tc = np.array([0, 50000, 51020, 100000, 101020, 400000, 401020, 500000, 501020, 600000, 601020, 700000]) * ps
    yc = [vdd, vdd, 0, 0, vdd, vdd, 0, 0, vdd, vdd, 0, 0]
    td = np.array([0, 200000, 201020, 300000, 301020, 380000, 381020, 600000]) * ps  # 3rd edge 20 ns early: no D/CLKN race
    yd = [0, 0, vdd, vdd, 0, 0, vdd, vdd]
    fc.vsource("VCLKN", "clkn", "0", Wave(W_PWL, t=list(tc), y=yc))
    fc.vsource("VD", "d", "0", Wave(W_PWL, t=list(td), y=yd))
    fc.set_outputs(["q", "d"])
    return fc, used


def dff_nominal_params() -> Dict[str, float]:
    out = {}
    for name, *_rest, nfin in _DFF_FETS:
        out[f"{name}.l"] = 21e-9
        out[f"{name}.nfin"] = float(nfin)
    return out


def dff_mc_params(fc: FlatCircuit, B: int, seed: int = 20240607, sigma: float = 0.02) -> np.ndarray:
    """Pre-drawn Monte-Carlo values, draw order [instance][device][l, nfin], z ~ N(0,1) truncated
    at +-3."""
    rng = np.random.default_rng(seed)
    z = np.clip(rng.standard_normal((B, len(_DFF_FETS), 2)), -3.0, 3.0)
    nom = dff_nominal_params()
    P = np.zeros((len(fc.param_names), B))
    for k, (name, *_rest) in enumerate(_DFF_FETS):
        P[fc.param_names.index(f"{name}.l")] = nom[f"{name}.l"] * (1.0 + sigma * z[:, k, 0])
        P[fc.param_names.index(f"{name}.nfin")] = nom[f"{name}.nfin"] * (1.0 + sigma * z[:, k, 1])
    return P


def two_resistor():
    """test/sweep.jl:326-340: V = 1 V across R1 + R2, sweep R1 x R2, I(V) = -1/(R1+R2)."""
    fc = FlatCircuit()
    fc.vsource("V", "vcc", "0", 1.0)
    fc.resistor("R1", "vcc", "out", fc.param("R1"))
    fc.resistor("R2", "out", "0", fc.param("R2"))
    fc.set_outputs(["v.i", "out"])
    return fc


# ---------------------------------------------------------------- decks of the small-signal parity tests
# test/bsimcmg/inverter_cmg_cedar.cir, the circuit behind the reference's ngspice noise table (test/ac.jl:161-237)
BSIMCMG_INVERTER_DECK = """** Test circuit
.include "jlpkg://ASAP7PDK/7nm_TT.pm"

* built-in
mneg Q D VSS VSS nmos_lvt
mpos Q D VDD VDD pmos_lvt

VVDD VDD 0 1.0
VVSS VSS 0 0.0
CQ D 0 1e-15
VD D 0 AC 1 SIN (0.5 0.01 1e7)

.TRAN 1e-9 4.0e-7

.END
"""
# the same inverter with the input bias as a parameter (swept in tests/test_gpu_ac_noise.py)
BSIMCMG_INVERTER_VIN_DECK = (BSIMCMG_INVERTER_DECK.replace("VD D 0 AC 1 SIN (0.5 0.01 1e7)", "VD D 0 DC 'vin' AC 1")
                             .replace("* built-in", ".param vin=0.5"))


# BASELINE config 5 stand-in: corner x temperature x mismatch transient sweep of one cell.  sky130 /
# BSIM4 are not in the reference tree, so the cell is the CMOS inverter of configs 1/2 on BSIM-CMG 107 + ASAP7 cards;
# a process corner is a pair of threshold shifts and mobility factors through the model's own variability handles
# (DELVTRAND, U0MULT: bsimcmg_body.include:133-137), temperature is SimSpec.temp, mismatch a draw of both gate lengths.
CONFIG5_DECK = """* corner x temperature x mismatch sweep of an inverter (config 5 stand-in)
.include "jlpkg://ASAP7PDK/7nm_TT.pm"
.param dvtn=0 dvtp=0 u0n=1 u0p=1 ln=21n lp=21n
mneg q d vss vss nmos_lvt l='ln' nfin=3 delvtrand='dvtn' u0mult='u0n'
mpos q d vdd vdd pmos_lvt l='lp' nfin=3 delvtrand='dvtp' u0mult='u0p'
VVDD vdd 0 0.7
VVSS vss 0 0.0
VD d 0 PWL(0 0 0.5n 0 0.6n 0.7 1.5n 0.7 1.6n 0 2.5n 0)
CQ q 0 1e-15
.TRAN 0.01n 2.5n
"""
CONFIG5_CORNERS = {   # tt, ff, ss, fs (nFET fast / pFET slow) as (dvtn, dvtp, u0n, u0p).  In BSIM-CMG 107 DELVTRAND adds to the gate
    # drive of either polarity (bsimcmg_body.include:2332), so a positive value is the fast (low-threshold) direction
    "dvtn": [0.0, 0.03, -0.03, 0.03], "dvtp": [0.0, 0.03, -0.03, -0.03], "u0n": [1.0, 1.08, 0.92, 1.08], "u0p": [1.0, 1.08, 0.92, 0.92]}


def config5_sweep(n_temp: int = 32, n_mismatch: int = 1024, seed: int = 130, sigma: float = 0.02):
    """ProductSweep(corner (4, tandem) x temp (-40 .. 125 C) x mismatch (tandem of pre-drawn gate lengths)),
    size (4, n_temp, n_mismatch); 4 x 32 x 1024 = 131 072 instances."""
    from .sweeps import ProductSweep, Sweep, TandemSweep
    rng = np.random.default_rng(seed)
    z = np.clip(rng.standard_normal((n_mismatch, 2)), -3.0, 3.0)
    return ProductSweep(TandemSweep(**CONFIG5_CORNERS), Sweep(temp=np.linspace(-40.0, 125.0, n_temp)),
                        TandemSweep(ln=21e-9 * (1.0 + sigma * z[:, 0]), lp=21e-9 * (1.0 + sigma * z[:, 1])))


# test/bsimcmg/asap7_inv.scs restated for the SPICE reader (the reference reads the cards from a Spectre file that is
# not in its tree): the gate is driven by a behavioural source in `$time`
ASAP7_INV_TIME_DECK = """* asap7_inv.scs: inverter on the ASAP7 cards, gate = 1.8*(1-sin(2 pi 1e7 t))
.include "jlpkg://ASAP7PDK/7nm_TT.pm"
.param vcc=1.8
M1p Vout Vgate VDD VDD pmos_lvt
M1n Vout Vgate 0 0 nmos_lvt
R1 Vout 0 10k
VScc VDD 0 'vcc'
Bgate Vgate 0 v=vcc*(1-sin(10.0**7*2*pi*$time))
"""


def small_signal_decks():
    """(deck text, swept columns) of the deck-based GPU tests; build() compiles them so that the GPU box finds
    their cubins in the cache."""
    one = np.array([1.0])
    return [(BSIMCMG_INVERTER_DECK, None),
            (BSIMCMG_INVERTER_VIN_DECK, {"vin": np.array([0.5]), "mneg.nfin": one}),
            (ASAP7_INV_TIME_DECK, {"vcc": 1.8 * one}),
            (CONFIG5_DECK, {"dvtn": 0 * one, "dvtp": 0 * one, "u0n": one, "u0p": one, "temp": 27 * one, "ln": 21e-9 * one,
                            "lp": 21e-9 * one})]
