"""The reference-facing sweep API on the GPU: these read like the reference's own sweep tests
(test/sweep.jl:326-371) with `dc_` / `tran_` in place of `dc!` / `tran!`."""
import numpy as np
import pytest

from cedarsim.jl_b200 import netlist
from cedarsim.jl_b200.sweeps import CircuitSweep, ProductSweep, SerialSweep, Sweep, TandemSweep, dc_, tran_
from oracle import orc

pytestmark = pytest.mark.gpu
DEFTOL = 1e-7

TWO_R = "* two resistor\n.param R1=100 R2=100\nV vcc 0 1\nRa vcc out 'R1'\nRb out 0 'R2'\n"


def test_dc_sweep_two_resistors():   # test/sweep.jl:326-340
    cs = CircuitSweep(TWO_R, ProductSweep(R1=np.arange(100.0, 2001, 100), R2=np.arange(100.0, 2001, 100)))
    sols = dc_(cs, abstol=DEFTOL, reltol=DEFTOL)
    assert sols.shape == (20, 20) and len(sols) == 400
    for sol, point in zip(sols, cs):
        p = dict(point)
        assert sol.retcode == "Success"
        assert abs(-1 / (p["R1"] + p["R2"]) - sol[cs.sys.v.I]) < DEFTOL
    # column-major result array like the Julia one: sols[i, j] <-> (R1[i], R2[j])
    assert abs(sols[3, 7][cs.sys.v.I] + 1 / (400.0 + 800.0)) < DEFTOL
    assert sols.array(cs.sys.node_out).shape == (20, 20)


def test_dc_sweep_on_spice_subckt_param():   # test/sweep.jl:342-371
    text = """* Parameter scoping test
.subckt subcircuit1 vss gnd
.param r_load=1
r1 vss gnd 'r_load'
.ends
.param v_in=1
x1 vss 0 subcircuit1
v1 vss 0 'v_in'
"""
    cs = CircuitSweep(text, ProductSweep(**{"v_in": np.arange(1.0, 11), "x1.r_load": np.arange(1.0, 11)}))
    sols = dc_(cs, abstol=DEFTOL, reltol=DEFTOL)
    for sol in sols:
        p = sol.params
        assert abs(p["v_in"] / p["x1.r_load"] + sol[cs.sys.v1.I]) < DEFTOL
        assert abs(p["v_in"] / p["x1.r_load"] - sol[cs.sys.x1.r1.I]) < DEFTOL     # the reference's own accessor, test/sweep.jl:369


def test_serial_and_tandem_sweeps_keep_defaults():
    cs = CircuitSweep(TWO_R, SerialSweep(R1=[100.0, 300.0], R2=[700.0]))
    sols = dc_(cs)
    want = [-1 / 200.0, -1 / 400.0, -1 / 800.0]    # None -> the netlist default (100)
    assert np.allclose([s[cs.sys.v.I] for s in sols], want, atol=1e-12)
    cs = CircuitSweep(TWO_R, TandemSweep(R1=[100.0, 200.0], R2=[300.0, 400.0]))
    assert np.allclose([s[cs.sys.v.I] for s in dc_(cs)], [-1 / 400.0, -1 / 600.0], atol=1e-12)


def test_tran_sweep_rc_and_interpolation():
    text = "* rc\n.param r=1k\nV1 in 0 PULSE(0 1 0 1n 1n 1 2)\nR1 in out 'r'\nC1 out 0 1n\n.tran 10n 5u\n"
    cs = CircuitSweep(text, Sweep(r=[500.0, 1000.0, 2000.0]), outputs=["out"])
    sols = tran_(cs, reltol=1e-5)
    assert sols.shape == (3,) and len(sols.t) == 501
    for sol, r in zip(sols, (500.0, 1000.0, 2000.0)):
        assert sol.retcode == "Success"
        tt = 2.0e-6
        assert abs(sol(tt, idxs=cs.sys.node_out) - (1 - np.exp(-(tt - 0.5e-9) / (r * 1e-9)))) < 2e-3


def test_bsimcmg_inverter_deck_tran(host_bsimcmg):   # test/bsimcmg/inverter.jl: retcode == Success
    text = """** Test circuit
.include "jlpkg://ASAP7PDK/7nm_TT.pm"
mneg Q D VSS VSS nmos_lvt
mpos Q D VDD VDD pmos_lvt
VVDD VDD 0 'vdd'
VVSS VSS 0 0.0
CQ D 0 1e-15
VD D 0 AC 1 SIN (0.35 0.3 1e7)
.param vdd=0.7
.TRAN 1e-9 4.0e-7
"""
    cs = CircuitSweep(text, ProductSweep(**{"vdd": [0.6, 0.7, 0.8], "mneg.nfin": [1.0, 2.0, 3.0]}), outputs=["q", "d"], host=True)
    sols = tran_(cs, reltol=1e-3)
    assert sols.shape == (3, 3) and all(s.retcode == "Success" for s in sols)
    y, st, _ = orc.tran(cs.flat.fc, 0.0, 4e-7, sols.t, params=cs.flat.params, opts=orc.default_options(reltol=1e-3))
    assert st.max() == 0 and np.abs(sols.y - y).max() < 5e-3
    q = sols.array(cs.sys.node_q)
    assert q.shape == (3, 3, len(sols.t)) and q.min() < 0.1 and q.max() > 0.5   # the inverter switches rail to rail


@pytest.mark.gpu
def test_verilog_a_include_dc_sweep():   # test/basic.jl:368-380 through the sweep API: sys.v1.I == -1/r for every point
    import os
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "va")
    deck = '* Verilog Include 2\n.hdl "va_resistor.va"\n\nx1 vcc 0 BasicVAResistor r=2k\nv1 vcc 0 dc=1\n'
    r = np.linspace(500.0, 4000.0, 64)
    cs = CircuitSweep(deck, Sweep("x1.r", r), include_dirs=[inc])
    sols = dc_(cs)
    assert sols.status.max() == 0
    assert np.allclose(sols.array(cs.sys.v1.I), -1.0 / r, rtol=1e-12, atol=0)


def test_va_branch_current_observable_sweep():   # test/varegress.jl through the sweep API: sys.xr.var"I(p, n)" >= 0
    import os
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "va")
    r = np.linspace(500.0, 2000.0, 32)
    ts = np.linspace(0, 1e-5, 101)
    for mod in ("VAR", "VAR_rev"):
        deck = f'* varegress\n.hdl "varegress.va"\nv1 vcc 0 1\nxr vcc out {mod} r=1000\nc1 out 0 1n\n'
        cs = CircuitSweep(deck, Sweep("xr.r", r), outputs=["out", "xr.I(p, n)"], include_dirs=[inc])
        sols = tran_(cs, (0.0, 1e-5), saveat=ts, reltol=1e-6, skip_dc=1)          # u0 = 0: uncharged start, as the reference's test
        assert sols.status.max() == 0
        cur = sols.array(getattr(cs.sys.xr, "I(p, n)"))          # (32, 101)
        out = sols.array(cs.sys.node_out)
        assert np.all(cur >= 0.0)
        assert np.abs(cur - (1.0 - out) / r[:, None])[:, 1:].max() < 1e-9          # (the t0 sample is the start vector)
        assert np.abs(out - (1.0 - np.exp(-ts[None, :] / (r[:, None] * 1e-9)))).max() < 1e-4


def test_behavioural_sources_dc_sweep():   # test/basic.jl:207-235 on the GPU, with the controlling gain as a sweep column
    deck = ("* Simple SPICE sources\n.param gain=2\nV1 0 1 1\nR1 1 0 1k\nB5 0 5 v='V(1)*gain'\nR5 5 0 1k\n"
            "E8 0 8 vol='V(0, 5)*gain'\nR8 8 0 r=1k\nG9 0 9 cur='V(0, 5)*gain'\nR9 9 0 r=1k\n")
    g = np.linspace(0.5, 4.0, 64)
    cs = CircuitSweep(deck, Sweep("gain", g))
    sols = dc_(cs)
    assert sols.status.max() == 0
    assert np.allclose(sols.array(cs.sys.node_5), g, rtol=1e-12)                    # 2.0 at gain = 2
    assert np.allclose(sols.array(cs.sys.node_8), g * g, rtol=1e-12)                # 4.0
    assert np.allclose(sols.array(cs.sys.node_9), -1000.0 * g * g, rtol=1e-12)      # -4000.0: no step limit for linear devices


def test_bsource_time_inverter_sweep(host_bsimcmg):   # test/bsimcmg/bsimcmg_spectre.jl:33-37 as a supply sweep on the GPU
    """`$time` in a behavioural source (the hidden ramp net `time__`): out node positive after DC init, transient succeeds,
    fixed-step parity with the oracle at the north star's tolerance."""
    from cedarsim.jl_b200 import circuits
    vcc = np.array([1.2, 1.5, 1.8])
    cs = CircuitSweep(circuits.ASAP7_INV_TIME_DECK, Sweep(vcc=vcc), outputs=["vout", "vgate", "time__"], host=True)
    dc = dc_(cs)
    assert all(s.retcode == "Success" for s in dc) and dc.array(cs.sys.node_vout).min() > 0.0
    ts = np.linspace(0.0, 1e-7, 101)
    kw = dict(fixed_step=1, dt=1e-10)
    sols = tran_(cs, (0.0, 1e-7), saveat=ts, **kw)
    assert all(s.retcode == "Success" for s in sols)
    assert np.abs(sols.array("time__") - ts[None, :]).max() < 1e-20
    assert np.abs(sols.array(cs.sys.node_vgate) - vcc[:, None] * (1 - np.sin(2e7 * np.pi * ts))[None, :]).max() < 1e-9
    yo, so, _ = orc.tran(cs.flat.fc, 0.0, 1e-7, ts, params=cs.flat.params, opts=orc.default_options(**kw))
    assert so.max() == 0
    err = np.abs(sols.y - yo)
    assert np.all(err <= 1e-6 * np.abs(yo) + 1e-9), err.max()


def test_multimode_source_reinitialised_at_t0_on_gpu():
    """test/basic.jl:534-552 (`v1 vcc 0 DC 5 SIN(10 3 1k)`: 10 V at t0 after the operating point at 5 V) as a sweep of the
    load resistor; and the RC variant: the algebraic node jumps, the capacitor voltage is held.  GPU == oracle."""
    from test_oracle_golden import MULTIMODE_DECK, MULTIMODE_RC_DECK
    cs = CircuitSweep(MULTIMODE_DECK.replace("r1 vcc 0 1k", ".param rl=1k\nr1 vcc 0 'rl'"), Sweep(rl=np.linspace(500.0, 2000.0, 40)),
                      outputs=["vcc", "v1.i"])
    dc = dc_(cs)
    assert np.abs(dc.array(cs.sys.node_vcc) - 5.0).max() < 1e-12
    ts = np.linspace(0.0, 1e-3, 11)
    sols = tran_(cs, (0.0, 1e-3), saveat=ts)
    assert all(s.retcode == "Success" for s in sols)
    assert np.abs(sols.array(cs.sys.node_vcc)[:, 0] - 10.0).max() < 1e-9
    yo, so, _ = orc.tran(cs.flat.fc, 0.0, 1e-3, ts, params=cs.flat.params)
    assert np.abs(sols.y - yo).max() < 1e-9
    cs = CircuitSweep(MULTIMODE_RC_DECK.replace("c1 out 0 1u", ".param cl=1u\nc1 out 0 'cl'"), Sweep(cl=np.linspace(0.5e-6, 2e-6, 33)),
                      outputs=["in", "out"])
    kw = dict(fixed_step=1, dt=1e-5)
    sols = tran_(cs, (0.0, 1e-3), saveat=ts, **kw)
    assert np.abs(sols.array("in")[:, 0] - 10.0).max() < 1e-9 and np.abs(sols.array("out")[:, 0] - 5.0).max() < 1e-6
    yo, so, _ = orc.tran(cs.flat.fc, 0.0, 1e-3, ts, params=cs.flat.params, opts=orc.default_options(**kw))
    assert np.all(np.abs(sols.y - yo) <= 1e-6 * np.abs(yo) + 1e-9)
    sols = tran_(cs, (0.0, 1e-3), saveat=ts, t0_reinit=0, **kw)
    assert np.abs(sols.array("in")[:, 0] - 5.0).max() < 1e-12


def test_abstime_inside_verilog_a_module_sweep():
    """`$abstime` read inside Verilog-A modules (hidden time port, va/compiler.py _lower_abstime) through the sweep API:
    amplitude sweep of a module sine source and a conductance growing with time, closed forms at every point."""
    import os
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "va")
    deck = ('* abstime\n.hdl "abstime_src.va"\nxs in 0 va_sine ampl=2 freq=1meg\nr1 in 0 1k\n'
            'v2 b 0 1\nr2 b c 1k\nxg c 0 va_ramp_g g0=1m tau=1u\n')
    amp = np.linspace(0.5, 3.0, 48)
    ts = np.linspace(0.0, 2e-6, 401)
    cs = CircuitSweep(deck, Sweep("xs.ampl", amp), outputs=["in", "c"], include_dirs=[inc])
    sols = tran_(cs, (0.0, 2e-6), saveat=ts, reltol=1e-6, dt_max=2e-9)
    assert sols.status.max() == 0
    vin = sols.array(cs.sys.node_in)
    assert np.abs(vin - amp[:, None] * np.sin(2e6 * np.pi * ts)[None, :]).max() < 2e-4
    g = 1e-3 * (1.0 + ts / 1e-6)
    assert np.abs(sols.array(cs.sys.node_c) - (1.0 / (1.0 + 1e3 * g))[None, :]).max() < 1e-9


def test_native_front_end_matches_the_python_one():
    """The same decks through the library's own netlist reader (front_end="native": cb_netlist_flatten -> cb_netlist_circuit,
    no Python parser / flattener / struct packing) and through the Python front end: identical results on the GPU, and the
    reference's known answers (test/sweep.jl:326-371)."""
    r1 = np.arange(100.0, 2001, 100)
    sw = ProductSweep(R1=r1, R2=r1)
    a = dc_(CircuitSweep(TWO_R, sw, front_end="native"))
    b = dc_(CircuitSweep(TWO_R, sw))
    cs = CircuitSweep(TWO_R, sw, front_end="native")
    assert a.status.max() == 0 and np.array_equal(a.array(cs.sys.v.I), b.array(cs.sys.v.I))
    R1, R2 = np.meshgrid(r1, r1, indexing="ij")
    assert np.abs(a.array(cs.sys.v.I) + 1.0 / (R1 + R2)).max() < 1e-12
    deck = ("* rc ladder in a subcircuit, pulse drive\n.param vdd=1 rr=1k\n.subckt stage a b r=1k c=1p\nr1 a b 'r'\nc1 b 0 'c'\n.ends\n"
            "v1 in 0 PULSE(0 'vdd' 1n 0.1n 0.1n 4n 10n)\nx1 in n1 stage r='rr'\nx2 n1 n2 stage r='2*rr' c=2p\nx3 n2 out stage\n")
    ts = np.linspace(0.0, 20e-9, 201)
    sw = ProductSweep(vdd=np.linspace(0.8, 1.2, 8), rr=np.linspace(500.0, 2000.0, 8))
    ya = tran_(CircuitSweep(deck, sw, outputs=["out", "x2.b", "v1.i"], front_end="native"), (0.0, 20e-9), saveat=ts, reltol=1e-5)
    yb = tran_(CircuitSweep(deck, sw, outputs=["out", "x2.b", "v1.i"]), (0.0, 20e-9), saveat=ts, reltol=1e-5)
    assert ya.status.max() == 0 and np.array_equal(ya.y, yb.y)
    # and a Spectre-language deck (test/basic.jl:265-278) through cb_netlist_flatten_spectre
    sub = "\nsubckt myres vcc gnd\n    parameters r=1k\n    r1 (vcc gnd) resistor r=r\nends myres\n\nx1 (vcc 0) myres r=2k\nv1  (vcc 0) vsource dc=1\n"
    rr = np.array([1e3, 2e3, 4e3])
    cs = CircuitSweep(sub, Sweep("x1.r", rr), lang="spectre", front_end="native")
    sp = dc_(cs)
    assert sp.status.max() == 0 and np.allclose(-sp.array(cs.sys.v1.I), 1.0 / rr, rtol=1e-14, atol=0)     # sys.x1.r1.I == 0.5e-3 at r = 2k


def test_failed_points_are_retried_by_the_host_ladder(tmp_path):
    """Host retry policy (sweeps.RETRY_LADDER; the reference restarts a failed CedarDCOp initialisation, src/dcop.jl:53-94):
    with three Newton iterations and no gmin / source stepping most operating points of a diode behind a resistor fail;
    the failed points alone are solved again with the patient options and end on the implicit solution."""
    (tmp_path / "dio.va").write_text('`include "disciplines.vams"\nmodule dio(p, n);\n    inout p, n;\n    electrical p, n;\n'
                                     '    parameter real is = 1e-14;\n    analog begin\n        I(p, n) <+ is * (limexp(V(p, n) / 0.025852) - 1.0);\n'
                                     '    end\nendmodule\n')
    text = f"* diode\n.hdl \"{tmp_path / 'dio.va'}\"\n.param vin=1\nV1 in 0 'vin'\nR1 in d 1k\nX1 d 0 dio\n"
    vin = np.linspace(0.2, 5.0, 64)
    hard = dict(max_newton_dc=3, gmin_steps=0, source_steps=0)
    raw = dc_(CircuitSweep(text, Sweep("vin", vin), outputs=["d"]), retry=False, **hard)
    assert (raw.status != 0).sum() > 16                       # the first pass alone leaves failures
    cs = CircuitSweep(text, Sweep("vin", vin), outputs=["d"])
    sols = dc_(cs, **hard)
    assert sols.status.max() == 0 and sols.stats["recovered_points"] == (raw.status != 0).sum()
    vd = sols.array(cs.sys.node_d)
    assert np.abs((vin - vd) / 1e3 - 1e-14 * (np.exp(vd / 0.025852) - 1.0)).max() < 1e-9      # KCL at the diode node
    ok = raw.status == 0
    assert np.array_equal(vd[ok], raw.array(cs.sys.node_d)[ok])                                 # converged points are untouched


def test_switch_branch_sweep_holds_both_kinds():
    """Switch branch (va/compiler.py _lower_switch_branches) on the GPU: the same device is a voltage source at some sweep
    points and a conductance at others; DC and a transient with a capacitor across it."""
    import os
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "va")
    deck = '* switch\n.hdl "switch_branch.va"\nv1 in 0 2\nr1 in out 1k\nc1 out 0 1n\nxs out 0 va_switch mode=0 v0=0.5 r=3k\n'
    cs = CircuitSweep(deck, ProductSweep(**{"xs.mode": [0.0, 1.0], "xs.v0": np.linspace(0.2, 1.8, 16)}), outputs=["out"], include_dirs=[inc])
    sols = dc_(cs)
    assert sols.status.max() == 0
    out = sols.array(cs.sys.node_out)                           # (2, 16)
    assert np.abs(out[0] - 1.5).max() < 1e-12 and np.abs(out[1] - np.linspace(0.2, 1.8, 16)).max() < 1e-12
    ts = np.linspace(0.0, 5e-6, 51)
    tr = tran_(cs, (0.0, 5e-6), saveat=ts, reltol=1e-6)
    assert tr.status.max() == 0 and np.abs(tr.array(cs.sys.node_out) - out[:, :, None]).max() < 1e-6     # stays at the operating point (LTE tolerance: vabstol = 1e-6 V)
