"""SPICE-subset reader / flattener semantics (reference src/spectre.jl, see netlist.py)."""
import os

import numpy as np
import pytest

from cedarsim.jl_b200 import modelcard, netlist
from cedarsim.jl_b200.expr import evaluate, parse_expr, parse_number
from cedarsim.jl_b200.flat import Col
from cedarsim.jl_b200.sweeps import CircuitSweep, ProductSweep


def test_numbers_and_magnitudes():   # test/basic.jl:609-638
    assert parse_number("0.22u") == parse_number("0.22e-6") == 0.22e-6
    assert parse_number("1MegQux") == 1e6 and parse_number("1Mil") == 25.4e-6 and parse_number("-1mAmp") == -1e-3
    assert parse_number("1Amp") == 1.0 and parse_number("1.8_V") == 1.8 and parse_number("10k") == 1e4


def test_expressions():   # functions int/nint/floor/ceil/pow/ln: test/basic.jl:651-684
    env = {"a": 2.0}
    assert evaluate(parse_expr("'a*3+pow(a,3)'"), env) == 14.0
    assert evaluate(parse_expr("{int(2.7)+nint(2.5)+floor(-0.5)+ceil(0.2)}"), env) == 2 + 2 - 1 + 1
    assert evaluate(parse_expr("a > 1 ? 10 : 20"), env) == 10
    assert abs(evaluate(parse_expr("ln(exp(1.5))"), env) - 1.5) < 1e-15
    v = evaluate(parse_expr("2*x"), {"x": np.array([1.0, 2.0])})
    assert v.tolist() == [2.0, 4.0]


def test_comments_continuations_case():
    nl = netlist.parse_netlist("""title line is ignored
* comment
R1 A 0 1k $ trailing
V1 a 0
+ DC 2 ; other trailing
.END
""")
    fl = netlist.flatten(nl)
    assert fl.fc.node_names == ["a"] and [d.name for d in fl.fc.devices] == ["r1", "v1"]
    assert fl.fc.devices[0].value == 1000.0 and fl.fc.waves[0].dc == 2.0


def test_param_scoping_and_sweep_columns():   # test/basic.jl:382-532 flavour; src/circuitodesystem.jl:66-97
    text = """* scoping
.param top=2 r_a='top*100'
.subckt div a b rr=10
r1 a m 'rr'
r2 m b 'rr*top'
.ends
x1 in 0 div rr='r_a'
x2 in 0 div
v1 in 0 'top'
"""
    nl = netlist.parse_netlist(text)
    fl = netlist.flatten(nl)
    vals = {d.name: d.value for d in fl.fc.devices}
    assert vals["x1.r1"] == 200.0 and vals["x1.r2"] == 400.0 and vals["x2.r1"] == 10.0 and vals["x2.r2"] == 20.0
    # sweeping `top` turns exactly the dependent values into per-point columns
    fl = netlist.flatten(nl, {"top": np.array([1.0, 2.0, 3.0])})
    vals = {d.name: d.value for d in fl.fc.devices}
    assert isinstance(vals["x1.r1"], Col) and isinstance(vals["x2.r2"], Col) and vals["x2.r1"] == 10.0
    assert fl.params.shape[1] == 3
    assert fl.params[vals["x1.r2"].index].tolist() == [100.0, 400.0, 900.0]
    with pytest.raises(netlist.NetlistError):
        netlist.flatten(nl, {"nosuch": np.array([1.0])})


def test_sources():
    nl = netlist.parse_netlist("""* sources
V1 a 0 PWL(0 0 1n 1 2n 1)
V2 b 0 DC 0.5 PULSE(0 1 1n 0.1n 0.1n 2n 5n)
I3 c 0 SIN(0 1m 1meg)
R1 a 0 1
R2 b 0 1
R3 c 0 1
""")
    fl = netlist.flatten(nl)
    w = fl.fc.waves
    assert w[0].kind == 1 and list(w[0].t) == [0.0, 1e-9, 2e-9] and w[0].dc is None
    assert w[1].kind == 2 and w[1].dc == 0.5 and list(w[1].v)[:2] == [0.0, 1.0]
    assert w[2].kind == 3 and list(w[2].v)[:3] == [0.0, 1e-3, 1e6]


def test_circuit_sweep_shapes_without_gpu():   # test/sweep.jl:244-319 (construction only)
    text = "* two r\nv1 vcc 0 1\nr1 vcc out 'r1v'\nr2 out 0 'r2v'\n.param r1v=100 r2v=100\n"
    cs = CircuitSweep(text, ProductSweep(r1v=np.arange(100.0, 2001, 100), r2v=np.arange(100.0, 2001, 100)))
    assert len(cs) == 400 and cs.size() == (20, 20) and cs.size(1) == 20 and cs.sweepvars() == {"r1v", "r2v"}
    assert cs.flat.params.shape == (2, 400)
    first = next(iter(cs))
    assert first == (("r1v", 100.0), ("r2v", 100.0))


def test_bsimcmg_deck(host_bsimcmg):   # test/bsimcmg/inverter_cmg_cedar.cir shape
    nl = netlist.parse_netlist("""** Test circuit
.include "jlpkg://ASAP7PDK/7nm_TT.pm"
mneg Q D VSS VSS nmos_lvt
mpos Q D VDD VDD pmos_lvt nfin=2
VVDD VDD 0 1.0
VVSS VSS 0 0.0
CQ D 0 1e-15
VD D 0 AC 1 SIN (0.5 0.01 1e7)
.TRAN 1e-9 4.0e-7
""")
    fl = netlist.flatten(nl, {"mneg.nfin": np.array([1.0, 2.0])})
    assert fl.tran == (1e-9, 4e-7) and len(fl.fc.va_insts) == 2 and fl.fc.n_nodes == 8
    assert "mneg.nfin" in fl.fc.param_names


BINNED_DECK = """* binned FinFET cards
.option scale={scale}
.param vth_shift=0.01
.model nf.0 nmos level=72 lmin=10n lmax=25n wmin=10n wmax=1u eot=1n dvt0='0.05+vth_shift'
.model nf.1 nmos level=72 lmin=25n lmax=100n wmin=10n wmax=1u eot=1.2n
.model pf.0 pmos level=17 version=107 lmin=10n lmax=100n wmin=10n wmax=1u
m1 d g 0 0 nf l=21n w=50n nfin=2
vg g 0 0.5
vd d 0 0.7
"""


def test_model_level_selects_family_and_devtype():   # spice_select_device / devtype_param, src/spectre.jl:596-641
    nl = netlist.parse_netlist(BINNED_DECK.format(scale=1.0))
    assert {k: c.master for k, c in nl.cards.items()} == {"nf.0": "bsimcmg107", "nf.1": "bsimcmg107", "pf.0": "bsimcmg107"}
    assert nl.cards["nf.0"].params["DEVTYPE"] == 1.0 and nl.cards["pf.0"].params["DEVTYPE"] == 0.0
    assert nl.cards["nf.0"].exprs == {"DVT0": "0.05+vth_shift"}          # needs the `.param` scope
    b4 = modelcard.parse_model_cards(".model n4 nmos level=54 version=4.5 toxe=4n\n.model p4 pmos level=14")
    assert b4["n4"].master == "bsim4" and b4["n4"].params["TYPE"] == 1.0 and b4["p4"].params["TYPE"] == -1.0


def test_model_binning(host_bsimcmg):   # find_bin, src/spectre.jl:1160-1170; test/binning/bins.jl:20-23
    cards = netlist.parse_netlist(BINNED_DECK.format(scale=1.0)).cards
    assert modelcard.find_bin(cards, "nf", 21e-9, 50e-9).name == "nf.0"
    assert modelcard.find_bin(cards, "nf", 25e-9, 50e-9).name == "nf.1"      # lmax is exclusive, lmin inclusive
    assert modelcard.find_bin(cards, "nf", 10e-9, 10e-9).name == "nf.0"
    assert modelcard.find_bin(cards, "nf", 42e-9, 100e-9, scale=0.5).name == "nf.0"   # scale multiplies l and w
    with pytest.raises(modelcard.NoBinException):
        modelcard.find_bin(cards, "nf", 21e-9, 1e-6)                         # wmax exclusive
    # through the flattener: the instance picks its bin, the bin's card (with `.param` expressions resolved) is compiled in
    fl = netlist.flatten(netlist.parse_netlist(BINNED_DECK.format(scale=1.0)), {"m1.l": np.array([20e-9, 24e-9])})
    assert len(fl.models) == 1 and "nf_0" in fl.models[0].name and fl.fc.param_names == ["m1.l"]
    fl2 = netlist.flatten(netlist.parse_netlist(BINNED_DECK.format(scale=2.0)), {"m1.l": np.array([20e-9, 24e-9])})
    assert "nf_1" in fl2.models[0].name                                       # `.option scale` moves the lookup only
    assert fl2.params[0].tolist() == [20e-9, 24e-9]
    with pytest.raises(netlist.NetlistError, match="crosses bins"):
        netlist.flatten(netlist.parse_netlist(BINNED_DECK.format(scale=1.0)), {"m1.l": np.array([20e-9, 30e-9])})
    with pytest.raises(netlist.NetlistError, match="NoBinExpection"):
        netlist.flatten(netlist.parse_netlist(BINNED_DECK.format(scale=1.0)), {"m1.l": np.array([2e-7])})


def test_reference_bins_deck():   # test/binning/bins.cir + bins.jl:20-23 (BSIM4 cards: lookup only, the family is not in tree)
    path = "/root/reference/test/binning/bins.cir"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    cards = modelcard.load_model_cards(path)
    assert len(modelcard.bins_of(cards, "nmos_3p3")) >= 5 and cards["nmos_3p3.0"].master == "bsim4"
    assert modelcard.find_bin(cards, "nmos_3p3", 2.8e-7, 2.2e-7).name == "nmos_3p3.0"
    assert modelcard.find_bin(cards, "nmos_3p3", 5.0e-7, 2.2e-7).name == "nmos_3p3.1"


def test_if_elseif_else_chain():   # `.if/.elseif/.else/.endif` (src/spectre.jl:1445-1525): exactly one branch is live
    deck = "* chain\n.param sel={sel}\nv1 a 0 1\n.if (sel == 1)\nr1 a 0 1\n.elseif (sel == 2)\nr1 a 0 2\n.else\nr1 a 0 4\n.endif\n"
    for sel, want in ((1, 1.0), (2, 2.0), (3, 4.0)):
        nl = netlist.parse_netlist(deck.format(sel=sel))
        cards = [c for c in nl.top.cards if c.name == "r1"]
        assert len(cards) == 1 and float(cards[0].value) == want
