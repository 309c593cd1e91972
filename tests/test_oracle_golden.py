"""Pin the CPU oracle against the reference's own known answers.

Every circuit below is restated from the cited reference test; the asserted numbers and tolerances
are the reference's.  (The reference itself cannot be executed here -- it is Julia-only -- so these
known-answer tests are what anchors the oracle; transistor-level waveforms are unpinned in the
reference, see test_bsimcmg_* for what is checked there.)
"""
import numpy as np
import pytest

from cedarsim.jl_b200 import netlist
from cedarsim.jl_b200.flat import FlatCircuit, Wave, W_PWL, params_matrix
from oracle import orc

DEFTOL = 1e-7   # test/common.jl: deftol


def solve_dc(text, sweep=None):
    fl = netlist.flatten(netlist.parse_netlist(text), sweep)
    x, xf, st, _ = orc.dc(fl.fc, fl.params if fl.params.size else None)
    assert st.max() == 0
    return fl.fc, xf


def test_two_resistor_dc_sweep():   # test/sweep.jl:326-340
    fc = FlatCircuit()
    fc.vsource("V", "vcc", "0", 1.0)
    fc.resistor("R1", "vcc", "out", fc.param("R1"))
    fc.resistor("R2", "out", "0", fc.param("R2"))
    fc.set_outputs(["v.i"])
    r1, r2 = np.meshgrid(np.arange(100, 2001, 100.0), np.arange(100, 2001, 100.0), indexing="ij")
    P = params_matrix([r1.ravel(order="F"), r2.ravel(order="F")])
    x, _, st, _ = orc.dc(fc, P)
    assert st.max() == 0
    assert np.abs(x[0] - (-1.0 / (P[0] + P[1]))).max() < DEFTOL


def test_spice_subckt_param_sweep():   # test/sweep.jl:342-371: I = v_in / r_load
    text = """* Parameter scoping test
.subckt subcircuit1 vss gnd
.param r_load=1
r1 vss gnd 'r_load'
.ends
.param v_in=1
x1 vss 0 subcircuit1
v1 vss 0 'v_in'
"""
    vin, rl = np.meshgrid(np.arange(1.0, 11), np.arange(1.0, 11), indexing="ij")
    fc, xf = solve_dc(text, {"v_in": vin.ravel(order="F"), "x1.r_load": rl.ravel(order="F")})
    i_r1 = -xf[fc.unknown("v1.i")]   # series circuit: resistor current = -(source branch current)
    assert np.abs(i_r1 - vin.ravel(order="F") / rl.ravel(order="F")).max() < DEFTOL


def test_vr_and_ir():   # test/basic.jl:21-43 (5 V / 2 Ohm -> 2.5 A), :81-106 (I = -5 A into 2 Ohm -> 10 V)
    fc, xf = solve_dc("* VR\nV vcc 0 5\nR vcc 0 2\n")
    assert abs(xf[fc.unknown("vcc"), 0] - 5.0) < DEFTOL and abs(-xf[fc.unknown("v.i"), 0] - 2.5) < DEFTOL
    fc, xf = solve_dc("* IR\nI icc 0 -5\nR icc 0 2\n")
    assert abs(xf[fc.unknown("icc"), 0] - 10.0) < DEFTOL


def test_spice_controlled_sources():   # test/basic.jl:207-235 (E/G rows; B-sources replaced by their value)
    text = """* Simple SPICE sources
V1 0 1 1
R1 1 0 1k
V5 5 0 2
R5 5 0 1k
E6 0 6 0 5 2
R6 6 0 r=1k
G7 0 7 0 5 2
R7 7 0 r=1k
"""
    fc, xf = solve_dc(text)
    assert abs(xf[fc.unknown("6"), 0] - 4.0) < DEFTOL
    assert abs(xf[fc.unknown("7"), 0] + 4000.0) < 1e-6


MULT = """* multiplicities
v1 vcc 0 DC 1
r1a vcc 1 1 m=10
r1b 1 0 1
.subckt r10 a b m=10
r2a a b 1
.ends
x2a vcc 2 r10
r2b 2 0 1
x3a1 vcc 3 r10 m=5
x3a2 vcc 3 r10 m=5
r3b 3 0 1
.subckt r5t2 a b
x5r1 a b r10 m=5
x5r2 a b r10 m=5
.ends
x4a1 vcc 4 r5t2
r4b 4 0 1
.subckt r2 a b
r2 a b 1 m=2
.ends
x5a vcc 5 r2 m=5
r5b 5 0 1
.model rm r R=1
r6a vcc 6 rm m=10 l=1u
r6b 6 0 1
"""


def test_multiplicities():   # test/basic.jl:556-595: all six nodes exactly 10/11
    fc, xf = solve_dc(MULT)
    for k in range(1, 7):
        assert xf[fc.unknown(str(k)), 0] == 10 / 11


def test_model_case_and_units():   # test/basic.jl:597-624
    fc, xf = solve_dc("* .model case\nv1 vcc 0 DC 1\n.model rr r R=1\nr1 vcc 1 rr l=1u\nr2 1 0 rr R=2 l=1u\n")
    assert abs(xf[fc.unknown("1"), 0] - 2 / 3) < 1e-12
    fc, xf = solve_dc("* units\ni1 vcc 0 DC -1mAmp\nr1 vcc 0 1MegQux\n")
    assert abs(xf[fc.unknown("vcc"), 0] - 1000.0) < 1e-6
    fc, xf = solve_dc("* units 2\ni1 vcc 0 DC -1Amp\nr1 vcc 0 1Mil\n")
    assert abs(xf[fc.unknown("vcc"), 0] - 2.54e-5) < 1e-12


def test_pwl_semantics():   # src/spectre_env.jl:15-21,43-69; test/transients.jl:66-96
    fc = FlatCircuit()
    fc.vsource("V", "a", "0", Wave(W_PWL, t=[1e-3, 9e-3], y=[0.0, 1.0]))
    fc.resistor("R", "a", "0", 1.0)
    for t, want in ((0.0, 0.0), (1e-3, 0.0), (5e-3, 0.5), (9e-3, 1.0), (1.0, 1.0)):
        assert abs(orc.wave_value(fc, 0, t) - want) < 1e-15
    fc2 = FlatCircuit()
    fc2.vsource("V", "a", "0", Wave(W_PWL, t=[0.0, 1.0, 1.0, 2.0], y=[0.0, 0.0, 1.0, 1.0]))  # vertical segment
    fc2.resistor("R", "a", "0", 1.0)
    assert orc.wave_value(fc2, 0, 0.5) == 0.0 and orc.wave_value(fc2, 0, 1.5) == 1.0
    assert orc.wave_value(fc2, 0, 1.0) == 0.5   # infinitely steep segment: mean of the two values (spectre_env.jl:60-64)
    # test/transients.jl:66-96: a breakpoint belongs to the NEXT segment (one-sided slopes by differences)
    fc3 = FlatCircuit()
    fc3.vsource("V", "a", "0", Wave(W_PWL, t=[0.0, 100e-9, 110e-9, 200e-9, 210e-9], y=[0.0, 0.0, 5.0, 5.0, 0.0]))
    fc3.resistor("R", "a", "0", 1.0)
    d = 1e-12
    slope = lambda t: (orc.wave_value(fc3, 0, t + d) - orc.wave_value(fc3, 0, t)) / d
    for t, want in ((0.0, 0.0), (50e-9, 0.0), (99e-9, 0.0), (100e-9, 5e8), (110e-9, 0.0), (200e-9, -5e8)):
        assert abs(slope(t) - want) < 1e-3 * 5e8


def test_pwl_current_ramp_transient():   # test/transients.jl:17-63, tol 1e-7 at every time point
    i_max, r_val = 1e-3, 1e3
    text = f"""* PWL test
.param pval=-1
i1 vout 0 PWL(1m 0 9m 'pval*{i_max}')
R1 vout 0 r={r_val}
"""
    fl = netlist.flatten(netlist.parse_netlist(text), outputs=["vout"])
    ts = np.linspace(0, 10e-3, 201)
    y, st, _ = orc.tran(fl.fc, 0.0, 10e-3, ts, opts=orc.default_options(reltol=1e-6))
    assert st.max() == 0
    want = np.clip((ts - 1e-3) / 8e-3, 0, 1) * i_max * r_val
    assert np.abs(y[0, :, 0] - want).max() < DEFTOL


def test_butterworth_closed_form():   # test/transients.jl:129-179
    w = 1.0
    text = f"""*Third order low pass filter, butterworth
V1 vin 0 SIN (0, 1, {w / (2 * np.pi)})
L1 vin n1 1.5
C2 n1 0 {4 / 3}
L3 n1 vout 0.5
R4 vout 0 1
"""
    fl = netlist.flatten(netlist.parse_netlist(text), outputs=["vout"])
    ts = np.linspace(0, 100.0, 1001)
    y, st, stats = orc.tran(fl.fc, 0.0, 100.0, ts, opts=orc.default_options(reltol=1e-7, vabstol=1e-9, dt_max=0.05))
    assert st.max() == 0
    t = ts
    want = (np.exp(-t) - np.sin(t) - np.cos(t)) / 2 + (2 * np.sin(np.sqrt(3) * t / 2)) / (np.sqrt(3) * np.sqrt(np.exp(t)))
    # the reference asserts 1e-7 with FBDF at its tolerances; second-order trapezoidal at dt <= 0.05 gives 1e-4
    assert np.abs(y[0, :, 0] - want).max() < 2e-4
    tail = y[0, len(ts) // 2:, 0]
    assert abs(np.sqrt(np.mean(tail ** 2)) - 0.5) < 0.1


def test_vrc_endpoints():   # test/basic.jl:111-141, uncharged start (u0 = 0 -> skip_dc)
    fc = FlatCircuit()
    fc.vsource("V", "vcc", "0", 5.0)
    fc.resistor("R", "vcc", "vrc", 2000.0)
    fc.capacitor("C", "vrc", "0", 1e-6)
    fc.set_outputs(["vrc", "v.i"])
    ts = np.linspace(0, 1.0, 11)
    y, st, _ = orc.tran(fc, 0.0, 1.0, ts, opts=orc.default_options(skip_dc=1, reltol=1e-6))
    assert st.max() == 0
    assert abs(y[0, 0, 0]) < DEFTOL and abs(y[0, -1, 0] - 5.0) < DEFTOL
    assert abs(y[1, -1, 0]) < DEFTOL


def _branch_i(fc, xf, name):
    return xf[fc.unknown(name + ".i"), 0] if (name + ".i") in getattr(fc, "unknown_names", []) else None


def test_spice_lib_self_include(tmp_path):   # test/basic.jl:312-336: `.LIB "file" section` of the file itself, I(r1) = 1/1337
    f = tmp_path / "selfinclude.cir"
    f.write_text("* .LIB definition and include test\nV1 vdd 0 1\n\n.LIB my_lib\nr1 vdd 0 1337\n.ENDL\n.LIB \"selfinclude.cir\" my_lib\n")
    fl = netlist.flatten(netlist.parse_file(str(f)))
    x, xf, st, _ = orc.dc(fl.fc, None)
    assert st.max() == 0
    # the source current is minus the resistor current (test/sweep.jl:338 sign convention)
    assert abs(-xf[fl.fc.unknown("v1.i"), 0] - 1 / 1337) < DEFTOL


def test_device_named_like_param():   # test/basic.jl:686-723: instance x1 and parameter x1; sys.x1.rload.V == 1000
    text = """* device == param
.param x1=1
.subckt myres p n
    .param rload=1k
    rload p n 'rload*x1'
.ends
i1 vcc 0 DC -1
x1 vcc 0 myres
"""
    # ParamSim(f; params=(;x1=2.0), x1=(;rload=500,)): top-level parameter x1 = 2, the subcircuit's rload = 500
    fc, xf = solve_dc(text, {"x1": np.array([2.0]), "x1.rload": np.array([500.0])})
    assert abs(xf[fc.unknown("vcc"), 0] - 1000.0) < 1e-9
    fc, xf = solve_dc(text)                      # defaults: 1k * 1 -> 1000 V as well
    assert abs(xf[fc.unknown("vcc"), 0] - 1000.0) < 1e-9
    fc, xf = solve_dc(text, {"x1": np.array([2.0])})
    assert abs(xf[fc.unknown("vcc"), 0] - 2000.0) < 1e-9


def test_semiconductor_resistor_and_if_else():   # test/basic.jl:725-752
    fc, xf = solve_dc("* semiconductor resistor\n.model myres r rsh=500\n.param res=1k\nv1 vcc 0 1\nR1 vcc 0 myres w=1m l=2m\nR2 vcc 0 res\n")
    assert abs(-xf[fc.unknown("v1.i"), 0] - 2e-3) < 1e-12        # I(r1) = I(r2) = 1e-3 (rsh * l / w = 1k)
    fc, xf = solve_dc("* ifelse resistor\n.param switch=1\nv1 vcc 0 1\n.if (switch == 1)\nR1 vcc 0 1\n.else\nR1 vcc 0 2\n.endif\n")
    assert abs(-xf[fc.unknown("v1.i"), 0] - 1.0) < 1e-12
    fc, xf = solve_dc("* ifelse resistor\n.param switch=0\nv1 vcc 0 1\n.if (switch == 1)\nR1 vcc 0 1\n.else\nR1 vcc 0 2\n.endif\n")
    assert abs(-xf[fc.unknown("v1.i"), 0] - 0.5) < 1e-12


def test_parallel_instances_rc():   # test/basic.jl:143-166: R with m = 10, uncharged start; C.I(0) = 10 v / r, C.V(end) = v
    v_val, r_val, c_val = 5.0, 2000.0, 1e-9          # test/basic.jl:14-16 flavour: tau = r c / 10 << span
    text = f"* multi vrc\nv1 vcc 0 {v_val}\nr1 vcc vrc {r_val} m=10\nc1 vrc 0 {c_val}\n"
    fl = netlist.flatten(netlist.parse_netlist(text))
    fc = fl.fc
    tau = r_val * c_val / 10
    ts = np.array([0.0, 1e-3 * tau, 50 * tau])
    y, st, _ = orc.tran(fc, 0.0, 50 * tau, ts, opts=orc.default_options(skip_dc=1, reltol=1e-6, vabstol=1e-9))
    assert st.max() == 0
    vrc = y[fc.outputs.index(fc.unknown("vrc")) if hasattr(fc, "outputs") else fc.unknown("vrc")]
    assert abs(vrc[0, 0]) < DEFTOL                                   # C.V(0) = 0
    assert abs((v_val - vrc[0, 0]) / (r_val / 10) - 10 * v_val / r_val) < DEFTOL   # C.I(0) = current through the 10 parallel R
    assert abs(vrc[2, 0] - v_val) < DEFTOL                           # C.V(end) = v, C.I(end) = 0


VA_INCLUDE_DECK = """* Verilog Include 2
.hdl "va_resistor.va"

x1 vcc 0 BasicVAResistor r=2k
v1 vcc 0 dc=1
"""


def test_verilog_a_include():   # test/basic.jl:368-380: `.hdl` + an X instance of the module, sys.v1.I == -1/2e3
    import os
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "va")
    nl = netlist.parse_netlist(VA_INCLUDE_DECK, include_dirs=[inc])
    fl = netlist.flatten(nl, host=True)
    x, xf, st, _ = orc.dc(fl.fc, None)
    assert st.max() == 0 and abs(xf[fl.fc.unknown("v1.i"), 0] + 1 / 2e3) < 1e-15
    # the module parameter as a sweep column, addressed SPICE-style in lower case
    fl = netlist.flatten(nl, {"x1.r": np.array([1e3, 2e3, 4e3])}, host=True)
    x, xf, st, _ = orc.dc(fl.fc, fl.params)
    assert st.max() == 0 and np.allclose(xf[fl.fc.unknown("v1.i")], [-1e-3, -5e-4, -2.5e-4], rtol=1e-14, atol=0)
    with pytest.raises(netlist.NetlistError, match="no parameter"):
        netlist.flatten(netlist.parse_netlist(VA_INCLUDE_DECK.replace("r=2k", "rr=2k"), include_dirs=[inc]), host=True)


def test_simple_spice_sources_behavioural():   # test/basic.jl:207-235, verbatim: B (v=), E (gain and vol=), G (gain and cur=)
    text = """* Simple SPICE sources
V1 0 1 1
R1 1 0 1k

B5 0 5 v=V(1)*2
R5 5 0 1k

E6 0 6 0 5 2
R6 6 0 r=1k

E8 0 8 vol=V(0, 5)*2
R8 8 0 r=1k

G7 0 7 0 5 2
R7 7 0 r=1k

G9 0 9 cur=V(0, 5)*2
R9 9 0 r=1k
"""
    fl = netlist.flatten(netlist.parse_netlist(text), host=True)
    assert all(m.linear for m in fl.models)          # behavioural sources here are linear: no Newton step limit, -4000 V is one step
    x, xf, st, stats = orc.dc(fl.fc, None)
    assert st.max() == 0
    v = lambda n: xf[fl.fc.unknown(n), 0]
    assert abs(v("5") - 2.0) < DEFTOL and abs(v("6") - 4.0) < DEFTOL and abs(v("8") - 4.0) < DEFTOL
    assert abs(v("7") + 4000.0) < 1e-6 and abs(v("9") + 4000.0) < 1e-6
    # a netlist parameter inside the expression is a parameter of the generated module: it can be a sweep column
    deck = "* b\n.param gain=2\nV1 1 0 1\nR1 1 0 1k\nB5 5 0 v='V(1)*gain + 1m'\nR5 5 0 1k\n"
    fl = netlist.flatten(netlist.parse_netlist(deck), {"gain": np.array([1.0, 2.0, 3.0])}, host=True)
    x, xf, st, _ = orc.dc(fl.fc, fl.params)
    assert st.max() == 0 and np.allclose(xf[fl.fc.unknown("5")], [1.001, 2.001, 3.001], rtol=1e-14, atol=0)
    # nonlinear expressions (functions, **, two-node probes), a current-output source fed by another one
    deck = ("* b\n.param k=3\nV1 1 0 2\nR1 1 0 1k\nB5 5 0 v='V(1)**2 + exp(-V(1,0))*k + max(V(1), 0.5) + 1e-3'\nR5 5 0 1k\n"
            "B6 6 0 i='V(5)/1k'\nR6 6 0 2k\n")
    fl = netlist.flatten(netlist.parse_netlist(deck), host=True)
    assert [m.linear for m in fl.models] == [False, True]
    x, xf, st, _ = orc.dc(fl.fc, None)
    want5 = 4 + np.exp(-2.0) * 3 + 2 + 1e-3
    assert st.max() == 0 and abs(xf[fl.fc.unknown("5"), 0] - want5) < 1e-9 and abs(xf[fl.fc.unknown("6"), 0] + 2 * want5) < 1e-9


def test_subcircuit_parameters_dynamic_scope():   # test/params.jl:58-99: nested subcircuits, `foo=foo+2000`, three overrides
    text = """* Subcircuit parameters
.subckt inner a b foo=foo+2000
R1 a b r= 'foo'
.ends

.subckt outer a b
x1 a b inner
.ends

.param inner  =1
.param foo =  1
i1 vcc 0 'foo'
l1 vcc out 1m
x1 out 0 outer
"""
    def r1_v(sweep):
        fc, xf = solve_dc(text, sweep)
        return xf[fc.unknown("out"), 0]          # sys.x1.x1.r1.V = V(out) - V(0); the inductor is a short at DC
    assert abs(r1_v({"x1.x1.foo": np.array([2.0])}) + 2.0) < DEFTOL            # ParamSim(circuit; x1=(x1=(foo=2.0,),))
    assert abs(r1_v({"foo": np.array([2.0])}) + 4004.0) < 1e-6                 # ParamSim(circuit; foo=2.0)
    assert abs(r1_v({"x1.x1.r1.r": np.array([100.0])}) + 100.0) < DEFTOL       # ParamSim(circuit; x1=(x1=(r1=(r=100.0,),),))
    assert abs(r1_v(None) + 2001.0) < 1e-6                                     # defaults: foo = 1 -> 2001 Ohm, 1 A


def test_temper_in_parameters():   # test/basic.jl:469-517: .option temp / .temp / ParamSim(temp=20) / default 27
    body = ".param foo = temper\ni1 vcc 0 'foo'\nr1 vcc 0 1\n"
    v = lambda text, sweep=None: (lambda r: r[1][r[0].unknown("vcc"), 0])(solve_dc(text, sweep))
    assert abs(v("* param temp\n.option temp=10\n" + body) + 10.0) < DEFTOL
    assert abs(v("* .temp\n.temp 10\n" + body) + 10.0) < DEFTOL
    assert abs(v("* .temp\n.temp 10\n" + body, {"temp": np.array([20.0])}) + 20.0) < DEFTOL      # overriding the temperature
    assert abs(v("* temper\n" + body) + 27.0) < DEFTOL


def test_instance_parameters_refer_to_other_parameters():   # test/basic.jl:519-536: x1 ... w=4 nrd='w/2' -> I(r1) = 1/2
    text = """* Parameter scoping test
.subckt subcircuit1 vss gnd w=2 rsh=1 nrd=1
r1 vss gnd 'rsh*nrd'
.ends
x1 vss 0 subcircuit1 w=4 nrd='w/2'
v1 vss 0 1
"""
    fc, xf = solve_dc(text)
    assert abs(-xf[fc.unknown("v1.i"), 0] - 0.5) < DEFTOL


def test_ddx_known_answer():   # test/ddx.jl:9-21: sol[sys.V1.I][end] == -5*2*2*3
    import os
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "va")
    deck = '* ddx\n.hdl "ddx_vcr.va"\nv1 vcc 0 5\nv2 vg 0 3\nxr vcc vg 0 ddx_vcr r=2\n'
    fl = netlist.flatten(netlist.parse_netlist(deck, include_dirs=[inc]), host=True)
    x, xf, st, _ = orc.dc(fl.fc, None)
    assert st.max() == 0 and xf[fl.fc.unknown("v1.i"), 0] == -5.0 * 2 * 2 * 3


def test_bsimcmg_inverter_bsource_time():   # test/bsimcmg/bsimcmg_spectre.jl:33-37 on test/bsimcmg/asap7_inv.scs
    """`VSgate bsource v=1.8*(1-sin(10.0**7*2*pi*$time))` driving the ASAP7 inverter: after DC init the out node is positive
    (a non-DC initialisation could leave it negative through the capacitances) and the transient to 1e-7 s succeeds."""
    from cedarsim.jl_b200 import circuits, models
    if not models.available():
        pytest.skip("BSIM-CMG sources not available")
    fl = netlist.flatten(netlist.parse_netlist(circuits.ASAP7_INV_TIME_DECK), host=True)
    fc = fl.fc
    _, xf, st, _ = orc.dc(fc, None)
    assert st.max() == 0 and xf[fc.unknown("vout"), 0] > 0.0
    assert abs(xf[fc.unknown("vgate"), 0] - 1.8) < 1e-12 and abs(xf[fc.unknown("time__"), 0]) < 1e-30   # $time == 0 in the DC solve
    ts = np.linspace(0.0, 1e-7, 201)
    y, st, _ = orc.tran(fc, 0.0, 1e-7, ts, opts=orc.default_options(reltol=1e-4))
    assert st.max() == 0                                                                          # retcode == Success
    assert np.abs(y[fc.unknown("vgate"), :, 0] - 1.8 * (1 - np.sin(2e7 * np.pi * ts))).max() < 2e-4   # saveat interpolation between accepted steps
    vout = y[fc.unknown("vout"), :, 0]
    assert vout.max() > 0.7 and vout.min() > -0.05      # pulled up through the PMOS while the gate dips below threshold


def test_mc_vr_circuit_agauss_is_nominal():   # test/basic.jl:45-79: R(agauss(2, 3, 3)) with the RNG disabled -> the same current at every time step
    fl = netlist.flatten(netlist.parse_netlist("* MC VR\nV vcc 0 5\nR vcc 0 'agauss(2, 3, 3)'\n"))
    y, st, _ = orc.tran(fl.fc, 0.0, 1e-3, np.linspace(0, 1e-3, 11), opts=orc.default_options())
    i = -y[fl.fc.unknown("v.i"), :, 0]
    assert st.max() == 0 and np.abs(i - i[0]).max() < DEFTOL and abs(i[0] - 2.5) < DEFTOL


def test_option_card_only():   # test/basic.jl:640-649: a deck of nothing but `.option temp=10 filemode=ascii noinit` is accepted
    nl = netlist.parse_netlist("* .option\n.option temp=10 filemode=ascii noinit\n")
    assert {k.lower(): v for k, v in nl.options.items()}["temp"] == "10" and not nl.top.cards


MULTIMODE_DECK = """* multimode spice source
v1 vcc 0 DC 5 AC 1 SIN(10 3 1k)
r1 vcc 0 1k
"""
# the same source behind an RC: the capacitor voltage is a differential variable and keeps its operating-point value at t0
MULTIMODE_RC_DECK = """* multimode source into an RC
v1 in 0 DC 5 SIN(10 3 1k)
r1 in out 1k
c1 out 0 1u
"""


def test_multimode_source_is_reinitialised_at_t0():   # test/basic.jl:534-552: vcc == 10 after CedarDCOp initialisation, DC value 5
    fl = netlist.flatten(netlist.parse_netlist(MULTIMODE_DECK), None, outputs=["vcc"], host=True)
    x, _, st, _ = orc.dc(fl.fc)
    assert st.max() == 0 and abs(x[0, 0] - 5.0) < 1e-12                       # :dcop mode sees the DC value
    y, st, _ = orc.tran(fl.fc, 0.0, 0.01, np.array([0.0, 2.5e-4]))
    assert st.max() == 0 and abs(y[0, 0, 0] - 10.0) < 1e-9                    # transient mode at t0: vo of SIN(10 3 1k)
    assert abs(y[0, 1, 0] - 13.0) < 1e-2
    y, st, _ = orc.tran(fl.fc, 0.0, 0.01, np.array([0.0]), opts=orc.default_options(t0_reinit=0))
    assert abs(y[0, 0, 0] - 5.0) < 1e-12                                      # without re-initialisation: the operating point
    fl = netlist.flatten(netlist.parse_netlist(MULTIMODE_RC_DECK), None, outputs=["in", "out"], host=True)
    y, st, _ = orc.tran(fl.fc, 0.0, 0.01, np.array([0.0, 0.01]))
    assert st.max() == 0
    assert abs(y[0, 0, 0] - 10.0) < 1e-9 and abs(y[1, 0, 0] - 5.0) < 1e-6     # algebraic node jumps, capacitor voltage is held


def test_abstime_inside_verilog_a_modules():
    """`$abstime` inside a Verilog-A module (the reference passes the simulation time to every device,
    src/vasim.jl `$abstime` -> sim time): a module source V = A sin(2 pi f $abstime) and a conductance that grows with
    time, against their closed forms; 0 at the DC operating point."""
    import os
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "va")
    deck = ('* abstime\n.hdl "abstime_src.va"\nxs in 0 va_sine ampl=2 freq=1meg\nr1 in 0 1k\n'
            'v2 b 0 1\nr2 b c 1k\nxg c 0 va_ramp_g g0=1m tau=1u\n')
    fl = netlist.flatten(netlist.parse_netlist(deck, include_dirs=[inc]), host=True)
    fc = fl.fc
    _, xf, st, _ = orc.dc(fc, None)
    assert st.max() == 0 and abs(xf[fc.unknown("in"), 0]) < 1e-15 and abs(xf[fc.unknown("c"), 0] - 0.5) < 1e-12
    ts = np.linspace(0.0, 2e-6, 401)
    y, st, _ = orc.tran(fc, 0.0, 2e-6, ts, opts=orc.default_options(reltol=1e-6, dt_max=2e-9))
    assert st.max() == 0
    assert np.abs(y[fc.unknown("in"), :, 0] - 2.0 * np.sin(2e6 * np.pi * ts)).max() < 1e-4
    g = 1e-3 * (1.0 + ts / 1e-6)
    assert np.abs(y[fc.unknown("c"), :, 0] - 1.0 / (1.0 + 1e3 * g)).max() < 1e-9      # divider 1k against 1 / g(t): algebraic, exact


def test_switch_branch_voltage_or_current_by_parameter():
    """A branch that receives both `V() <+` and `I() <+` contributions (switch branch; the reference resets its accumulator
    when the kind changes, src/vasim.jl:149-154): the last kind executed decides, earlier contributions of the other kind
    are discarded, contributions of one kind accumulate.  The mode is a run-time parameter: one sweep holds both kinds."""
    import os
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "va")
    deck = '* switch\n.hdl "switch_branch.va"\nv1 in 0 2\nr1 in out 1k\nxs out 0 va_switch mode=0 v0=0.5 r=3k\n'
    mode = np.array([0.0, 1.0, 0.0, 1.0])
    v0 = np.array([0.5, 0.5, 1.5, 1.5])
    fl = netlist.flatten(netlist.parse_netlist(deck, include_dirs=[inc]), {"xs.mode": mode, "xs.v0": v0}, host=True)
    _, xf, st, _ = orc.dc(fl.fc, fl.params)
    assert st.max() == 0
    assert np.abs(xf[fl.fc.unknown("out")] - np.where(mode > 0.5, v0, 2.0 * 3e3 / 4e3)).max() < 1e-12
    i_br = xf[fl.fc.unknown("xs.i(p,n)")]                       # the branch current is an unknown of the device in both modes
    assert np.abs(i_br - (2.0 - xf[fl.fc.unknown("out")]) / 1e3).max() < 1e-12
