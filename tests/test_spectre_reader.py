"""Spectre-language subset reader (cedarsim.jl_b200/spectre.py) against the reference's own Spectre-syntax tests."""
import os

import numpy as np
import pytest

from cedarsim.jl_b200 import netlist, spectre
from cedarsim.jl_b200.sweeps import CircuitSweep, Sweep
from oracle import orc

HERE = os.path.dirname(os.path.abspath(__file__))

SOURCES = """
I1 (0 1) isource dc=2.2u
R1 (1 0) resistor r=1000

I2 (0 2) isource type=pwl wave=[0 1m .5 2m 1 1.75m]
R2 (2 0) resistor r=2k

V3 (0 3) vsource dc=1.5
R3 (3 0) resistor r=1k

V4 (0 4) vsource type=pwl wave=[ 0 1 .5 2 \\
        1 5]
R4 (4 0) resistor r=4k
"""


def test_simple_spectre_sources(tmp_path):   # test/basic.jl:168-205 (read from a file, as there; the bsource rows follow in the next test)
    f = tmp_path / "sources.scs"
    f.write_text(SOURCES)
    fl = netlist.flatten(spectre.parse_spectre_file(str(f)))
    fc = fl.fc
    ts = np.linspace(0.0, 1.0, 11)
    y, st, _ = orc.tran(fc, 0.0, 1.0, ts, opts=orc.default_options())
    assert st.max() == 0
    v = lambda n: y[fc.unknown(n), :, 0]
    assert np.allclose(v("1"), 2.2e-3) and np.allclose(v("1") / 1000, 2.2e-6)          # node_1, R1.I
    assert np.allclose(v("3"), -1.5) and np.allclose(v("3") / 1e3, -1.5e-3)            # node_3, R3.I
    assert np.isclose(v("2")[-1], 3.5) and np.isclose(v("2")[-1] / 2e3, 1.75e-3)       # node_2[end], R2.I[end]
    assert np.isclose(v("4")[-1], -5.0) and np.isclose(v("4")[-1] / 4e3, -1.25e-3)     # node_4[end], R4.I[end]


def test_spectre_bsource_with_time():   # test/basic.jl:185-186, :203: B5 (0 5) bsource v=$time*V(3) -> node_5[end] == 1.5
    fl = netlist.flatten(spectre.parse_spectre(SOURCES + "B5 (0 5) bsource v=$time*V(3)\nR5 (5 0) resistor r=1k\n"), host=True)
    fc = fl.fc
    ts = np.linspace(0.0, 1.0, 11)
    y, st, _ = orc.tran(fc, 0.0, 1.0, ts, opts=orc.default_options())
    assert st.max() == 0
    assert np.allclose(y[fc.unknown("time__"), :, 0], ts, rtol=0, atol=1e-15)        # the hidden ramp IS the time
    assert np.allclose(y[fc.unknown("5"), :, 0], 1.5 * ts, atol=1e-9)                # V(0,5) = t * V(3) = -1.5 t
    assert np.isclose(y[fc.unknown("5"), -1, 0], 1.5) and np.isclose(y[fc.unknown("3"), -1, 0], -1.5)
    # a current-mode source, nonlinear in time: I(0,1) = 2 mA * t^2 into 1 kOhm
    fl = netlist.flatten(spectre.parse_spectre("B1 (0 1) bsource i=2m*$time*$time\nR1 (1 0) resistor r=1k\n"), host=True)
    y, st, _ = orc.tran(fl.fc, 0.0, 1.0, ts, opts=orc.default_options())
    assert st.max() == 0 and np.allclose(y[fl.fc.unknown("1"), :, 0], 2.0 * ts ** 2, atol=1e-9)
    with pytest.raises(netlist.NetlistError, match="bsource needs"):
        spectre.parse_spectre(SOURCES + "B5 (0 5) bsource r=1\n")


SUBCKT = """
subckt myres vcc gnd
    parameters r=1k
    r1 (vcc gnd) resistor r=r
ends myres

x1 (vcc 0) myres r=2k
v1  (vcc 0) vsource dc=1
"""


def test_simple_spectre_subcircuit():   # test/basic.jl:265-278: sys.x1.r1.I == 0.5e-3
    nl = spectre.parse_spectre(SUBCKT)
    fl = netlist.flatten(nl)
    x, xf, st, _ = orc.dc(fl.fc, None)
    assert st.max() == 0 and abs(-xf[fl.fc.unknown("v1.i"), 0] - 0.5e-3) < 1e-15
    # the subcircuit parameter as a sweep column through the sweep API's naming (x1.r)
    cs = CircuitSweep(nl, Sweep("x1.r", np.array([1e3, 2e3, 4e3])))
    x, xf, st, _ = orc.dc(cs.flat.fc, cs.flat.params)
    assert np.allclose(-xf[cs.flat.fc.unknown("v1.i")], [1e-3, 5e-4, 2.5e-4], rtol=1e-14, atol=0)


def test_spectre_ahdl_include():   # test/basic.jl:353-367: ahdl_include + instance of the module, sys.v1.I == -1/2e3
    ckt = 'ahdl_include "va_resistor.va"\n\nx1 (vcc 0) BasicVAResistor R=2k\nv1 (vcc 0) vsource dc=1\n'
    fl = netlist.flatten(spectre.parse_spectre(ckt, include_dirs=[os.path.join(HERE, "va")]), host=True)
    x, xf, st, _ = orc.dc(fl.fc, None)
    assert st.max() == 0 and abs(xf[fl.fc.unknown("v1.i"), 0] + 1 / 2e3) < 1e-15


def test_spectre_model_cards_and_sources(host_bsimcmg):
    # the BSIM-CMG inverter of test/bsimcmg/inverter_cmg_cedar.cir written in Spectre syntax: same flat circuit as the SPICE deck
    scs = """
include "jlpkg://ASAP7PDK/7nm_TT.scs"
mneg (q d vss vss) nmos_lvt
mpos (q d vdd vdd) pmos_lvt nfin=2
vvdd (vdd 0) vsource dc=1.0
vvss (vss 0) vsource dc=0.0
cq (d 0) capacitor c=1e-15
vd (d 0) vsource type=sine sinedc=0.5 ampl=0.01 freq=1e7 mag=1
"""
    spice = """** Test circuit
.include "jlpkg://ASAP7PDK/7nm_TT.pm"
mneg Q D VSS VSS nmos_lvt
mpos Q D VDD VDD pmos_lvt nfin=2
VVDD VDD 0 1.0
VVSS VSS 0 0.0
CQ D 0 1e-15
VD D 0 AC 1 SIN (0.5 0.01 1e7)
"""
    a = netlist.flatten(spectre.parse_spectre(scs), {"mneg.nfin": np.array([1.0, 2.0])})
    b = netlist.flatten(netlist.parse_netlist(spice), {"mneg.nfin": np.array([1.0, 2.0])})
    assert a.fc.node_names == b.fc.node_names and a.fc.branch_names == b.fc.branch_names
    assert len(a.fc.va_insts) == 2 and a.fc.param_names == b.fc.param_names
    assert [(w.kind, w.ac) for w in a.fc.waves] == [(w.kind, w.ac) for w in b.fc.waves]


def test_subcircuit_port_aliases():   # test/alias.jl:5-33: sys.x1.node_pos == sys.node_vcc, sys.x1.node_neg == sys.node_0
    ckt = """
subckt myres pos neg
    parameters r=1k
    r1 (pos neg) resistor r=r
ends myres

x1 (vcc 0) myres r=2k
v1 (vcc 0) vsource dc=1
"""
    fl = netlist.flatten(spectre.parse_spectre(ckt))
    fc = fl.fc
    assert fc.aliases == {"x1.pos": "vcc", "x1.neg": "0"}                       # the reference's aliasmap
    assert fc.unknown("x1.node_pos") == fc.unknown("node_vcc")
    with pytest.raises(KeyError, match="ground"):
        fc.unknown("x1.node_neg")
    # through the result objects: a grounded port reads as 0 V
    from cedarsim.jl_b200.sweeps import _Observable
    assert _Observable(fc, "x1.node_neg").terms == [] and _Observable(fc, "x1.node_pos").terms == [(1.0, fc.unknown("vcc"))]


def test_spectre_parameter_expressions():   # test/spectre_expr.jl:10-43
    code = """
parameters p1=23pf p2=.3 p3 = 1&2~^3 p4 = true && false || true p5 = M_1_PI * 3.0
r1 (1 0) resistor r=p1      // another simple expression // fdsfdsf
r2 (1 0) resistor r=p2*p2   // a binary multiply expression
r3 (1 0) resistor r=(p1+p2)/p3      // a more complex expression
r4 (1 0) resistor r=sqrt(p1+p2)     // an algebraic function call
r5 (1 0) resistor r=3+atan(p1/p2) //a trigonometric function call
r6 (1 0) resistor r=((p1<1) ? p4+1 : p3)  // the ternary operator
"""
    import math
    nl = spectre.parse_spectre(code)
    fl = netlist.flatten(nl)
    val = {d.name: d.value for d in fl.fc.devices}
    p1, p2, p3, p4, p5 = 23e-12, 0.3, float(~((1 & 2) ^ 3)), 1.0, 3.0 / math.pi
    from cedarsim.jl_b200.expr import evaluate, parse_expr
    got = {k: evaluate(parse_expr(v), {}) for k, v in nl.top.params.items()}
    assert abs(got["p1"] - p1) < 1e-26 and got["p2"] == p2 and got["p3"] == p3 and got["p4"] == p4 and got["p5"] == p5
    assert val["r1"] == pytest.approx(p1) and val["r2"] == pytest.approx(p2 * p2) and val["r3"] == pytest.approx((p1 + p2) / p3)
    assert val["r4"] == pytest.approx(math.sqrt(p1 + p2)) and val["r5"] == pytest.approx(3 + math.atan(p1 / p2))
    assert val["r6"] == pytest.approx(p4 + 1)
