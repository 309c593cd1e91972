"""Forward parameter sensitivities, restating the reference's known answers of
test/sensitivity.jl:14-41 (DefaultSim) and :43-68 (ParamLens): for V=1 over R1 -- out -- R2 with R1 = R2 = 1 the
sensitivity of `out` is dR1 = -out/2 = -1/4 and dR2 = -dR1 at every time point.

CPU half: the stencil layout / combination (host logic) with the perturbed points solved by the oracle.
GPU half: `sensitivities_` itself -- the stencil points are extra sweep points of one batched engine call."""
import numpy as np
import pytest

from cedarsim.jl_b200 import netlist
from cedarsim.jl_b200.sweeps import (CircuitSweep, ProductSweep, Sweep, combine_stencil, sensitivities_,
                                     sensitivity_columns)
from oracle import orc

TWO_R = "* two resistor\n.param R1=2000 R2=1\nV vcc 0 1\nRa vcc out 'R1'\nRb out 0 'R2'\n"
RC = "* rc\n.param r=1k c=1n\nV1 in 0 PWL(0 0 1n 1 1 1)\nR1 in out 'r'\nC1 out 0 'c'\n"
DIODE_VA = """`include "disciplines.vams"
module dio(a, c);
  inout a, c; electrical a, c;
  parameter real is = 1e-14;
  parameter real n = 1.0;
  analog I(a, c) <+ is * (limexp(V(a, c) / (n * 0.025852)) - 1.0);
endmodule
"""


def oracle_dc_sens(text, cols, wrt, order, rel_step=1e-3, node="out"):
    big, steps = sensitivity_columns(cols, wrt, rel_step, order)
    fl = netlist.flatten(netlist.parse_netlist(text), big)
    _, xf, st, _ = orc.dc(fl.fc, fl.params)
    assert st.max() == 0
    B = len(next(iter(cols.values())))
    v = xf[fl.fc.unknown(node)]
    return v[:B], [combine_stencil(v, B, j, steps[w], order) for j, w in enumerate(wrt)]


def test_stencil_layout():
    cols = {"a": np.array([1.0, 2.0, 0.0]), "b": np.array([10.0, 20.0, 30.0])}
    big, steps = sensitivity_columns(cols, ["b", "a"], rel_step=0.1, order=2)
    assert len(big["a"]) == 3 * (1 + 2 * 2)
    assert np.allclose(steps["a"], [0.1, 0.2, 0.1]) and np.allclose(steps["b"], [1, 2, 3])   # p == 0 -> absolute step
    assert np.allclose(big["b"][3:6], [9, 18, 27]) and np.allclose(big["b"][6:9], [11, 22, 33])
    assert np.allclose(big["a"][3:9], np.tile(cols["a"], 2))           # only the differentiated column moves
    assert np.allclose(big["a"][9:12], [0.9, 1.8, -0.1]) and np.allclose(big["a"][12:15], [1.1, 2.2, 0.1])
    # a polynomial of degree <= 4 is differentiated exactly by the 4th-order stencil
    big, steps = sensitivity_columns(cols, ["a"], rel_step=0.05, order=4)
    f = lambda a: 3 * a ** 4 - a ** 3 + 2 * a
    assert np.allclose(combine_stencil(f(big["a"]), 3, 0, steps["a"], 4), 12 * cols["a"] ** 3 - 3 * cols["a"] ** 2 + 2, atol=1e-12)
    with pytest.raises(KeyError):
        sensitivity_columns(cols, ["zz"])
    with pytest.raises(ValueError):
        sensitivity_columns({"a": np.array([1.0, np.nan])}, ["a"])


@pytest.mark.parametrize("order", [2, 4])
def test_two_resistor_known_answer_oracle(order):   # test/sensitivity.jl:14-41: dR1 == -dR2, dR1 == -out/2
    out, (d1, d2) = oracle_dc_sens(TWO_R, {"R1": np.array([1.0]), "R2": np.array([1.0])}, ["R1", "R2"], order)
    tol = 1e-6 if order == 2 else 1e-10
    assert abs(out[0] - 0.5) < 1e-12
    assert abs(d1[0] + d2[0]) < tol and abs(d1[0] + out[0] / 2) < tol


def test_param_lens_known_answer_oracle():   # test/sensitivity.jl:43-68: only R1 is a parameter, R2 keeps its default 1
    out, (d1,) = oracle_dc_sens(TWO_R, {"R1": np.array([1.0])}, ["R1"], 4)
    assert abs(d1[0] + out[0] / 2) < 1e-10


def test_sweep_of_closed_forms_oracle():   # d out / d R1 = -R2/(R1+R2)^2, d out / d R2 = R1/(R1+R2)^2 over a grid
    r1, r2 = np.meshgrid(np.arange(100.0, 1001, 300), np.arange(100.0, 1001, 300), indexing="ij")
    cols = {"R1": r1.ravel(order="F"), "R2": r2.ravel(order="F")}
    _, (d1, d2) = oracle_dc_sens(TWO_R, cols, ["R1", "R2"], 4)
    assert np.allclose(d1, -cols["R2"] / (cols["R1"] + cols["R2"]) ** 2, rtol=1e-9, atol=0)
    assert np.allclose(d2, cols["R1"] / (cols["R1"] + cols["R2"]) ** 2, rtol=1e-9, atol=0)


def test_column_sweep_and_no_cpu_fallback():
    """The expanded batch is an ordinary sweep of explicit columns; without a GPU the public call fails loudly in the
    engine (no CPU fallback) instead of differencing oracle solutions."""
    from cedarsim.jl_b200 import engine
    from cedarsim.jl_b200.sweeps import _ColumnSweep
    sw = _ColumnSweep({"b": [1.0, 2.0, np.nan], "a": [5.0, 6.0, 7.0]})
    assert sw.shape == (3,) and len(sw) == 3 and sw.sweepvars() == {"a", "b"}
    assert list(sw) == [(("a", 5.0), ("b", 1.0)), (("a", 6.0), ("b", 2.0)), (("a", 7.0), ("b", None))]
    cs = CircuitSweep(TWO_R, ProductSweep(R1=[1.0, 2.0], R2=[1.0]))
    with pytest.raises(KeyError, match="not a swept variable"):
        sensitivities_(cs, wrt=["R3"])
    with pytest.raises(ValueError, match="analysis must be"):
        sensitivities_(cs, analysis="pz")
    with pytest.raises(ValueError, match="order must be"):
        sensitivity_columns(cs.columns, ["R1"], order=3)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(engine.EngineError) as e:
            sensitivities_(cs)
        assert "no CPU fallback" in str(e.value)


# ---------------------------------------------------------------- GPU: the public call
@pytest.mark.gpu
def test_sensitivities_two_resistor_gpu():
    cs = CircuitSweep(TWO_R, ProductSweep(R1=[1.0], R2=[1.0]))
    sens = sensitivities_(cs)
    out = sens.solution.array(cs.sys.node_out)
    d1, d2 = sens.array(cs.sys.node_out, "R1"), sens.array(cs.sys.node_out, "R2")
    assert out.shape == d1.shape == (1, 1) and sens.retcodes[0, 0] == "Success"
    assert abs(d1[0, 0] + d2[0, 0]) < 1e-10 and abs(d1[0, 0] + out[0, 0] / 2) < 1e-10
    # transient form, as the reference poses it: the same numbers at every time point
    sens = sensitivities_(cs, analysis="tran", tspan=(0.0, 1e-5), saveat=np.linspace(0, 1e-5, 5), fixed_step=1, dt=1e-6)
    d1, d2 = sens.array(cs.sys.node_out, "R1"), sens.array(cs.sys.node_out, "R2")
    assert d1.shape == (1, 1, 5)
    assert np.abs(d1 + d2).max() < 1e-10 and np.abs(d1 + sens.solution.array(cs.sys.node_out) / 2).max() < 1e-10


@pytest.mark.gpu
def test_sensitivities_grid_and_observable_gpu():
    cs = CircuitSweep(TWO_R, ProductSweep(R1=np.arange(100.0, 1001, 100), R2=np.arange(100.0, 1001, 100)))
    sens = sensitivities_(cs, wrt=["R2"])
    R1, R2 = np.meshgrid(np.arange(100.0, 1001, 100), np.arange(100.0, 1001, 100), indexing="ij")
    assert sens.array(cs.sys.node_out, "R2").shape == (10, 10)
    assert np.allclose(sens.array(cs.sys.node_out, "R2"), R1 / (R1 + R2) ** 2, rtol=1e-9, atol=0)
    # an observable that depends on the parameter itself: I(Rb) = 1/(R1+R2), d/dR2 = -1/(R1+R2)^2
    assert np.allclose(sens.array(cs.sys.rb.I, "R2"), -1.0 / (R1 + R2) ** 2, rtol=1e-9, atol=0)
    assert abs(sens.point((3, 7), cs.sys.v.I, "R2") - 1.0 / (400.0 + 800.0) ** 2) < 1e-15
    assert (sens.retcodes == "Success").all()


@pytest.mark.gpu
def test_sensitivities_rc_transient_gpu():   # out(t) of a ramp-driven RC; d out / d r against the same stencil on the oracle
    r = np.array([500.0, 1000.0, 2000.0])
    ts = np.linspace(0, 5e-6, 51)
    kw = dict(fixed_step=1, dt=1e-8)
    cs = CircuitSweep(RC, Sweep(r=r), outputs=["out"])
    sens = sensitivities_(cs, analysis="tran", tspan=(0.0, 5e-6), saveat=ts, **kw)
    d = sens.array(cs.sys.node_out, "r")
    assert d.shape == (3, 51) and (sens.retcodes == "Success").all()
    big, steps = sensitivity_columns({"r": r}, ["r"], 1e-3, 4)
    fl = netlist.flatten(netlist.parse_netlist(RC), big, outputs=["out"])
    yo, so, _ = orc.tran(fl.fc, 0.0, 5e-6, ts, params=fl.params, opts=orc.default_options(**kw))
    assert so.max() == 0
    do = combine_stencil(yo[0], 3, 0, steps["r"], 4)            # [S, B]
    assert np.abs(d - do.T).max() <= 1e-6 * np.abs(do).max()
    # and the closed form of the step response after the 1 ns ramp (discretisation error of dt = 10 ns allowed for)
    tt = ts[10:] - 0.5e-9
    for k, rk in enumerate(r):
        want = -(tt / (rk * rk * 1e-9)) * np.exp(-tt / (rk * 1e-9))
        assert np.abs(d[k, 10:] - want).max() < 2e-3 * np.abs(want).max()


@pytest.mark.gpu
def test_sensitivities_nonlinear_va_diode_gpu(tmp_path):   # Newton-solved points: V -- R -- diode, d v(d) / d r vs implicit differentiation
    (tmp_path / "dio.va").write_text(DIODE_VA)
    text = f"* diode\n.hdl \"{tmp_path / 'dio.va'}\"\n.param r=1k vin=1\nV1 in 0 'vin'\nR1 in d 'r'\nX1 d 0 dio\n"
    cs = CircuitSweep(text, ProductSweep(r=[500.0, 1e3, 2e3], vin=[0.8, 1.0, 2.0]), outputs=["d"])
    sens = sensitivities_(cs)
    v = sens.solution.array(cs.sys.node_d)
    R, VIN = np.meshgrid([500.0, 1e3, 2e3], [0.8, 1.0, 2.0], indexing="ij")
    vt = 0.025852
    gd = 1e-14 * np.exp(v / vt) / vt                     # diode conductance at the solved point
    # (vin - v)/r = Id(v):  dv/dr = -(vin - v)/r^2 / (1/r + gd),  dv/dvin = (1/r) / (1/r + gd)
    assert np.allclose(sens.array(cs.sys.node_d, "r"), -(VIN - v) / R ** 2 / (1 / R + gd), rtol=1e-6, atol=0)
    assert np.allclose(sens.array(cs.sys.node_d, "vin"), (1 / R) / (1 / R + gd), rtol=1e-6, atol=0)


RC_AC = "* rc low-pass\n.param r=1k c=1n\nV1 in 0 DC 0 AC 1\nR1 in out 'r'\nC1 out 0 'c'\n"


@pytest.mark.gpu
def test_sensitivities_ac_and_noise_gpu():   # H = 1/(1 + j w r c): dH/dr = -j w c H^2;  S_out = 4 k T r |H|^2
    r = np.array([500.0, 1000.0, 2000.0])
    f = np.logspace(3, 7, 9)
    cs = CircuitSweep(RC_AC, Sweep(r=r), outputs=["out"])
    sens = sensitivities_(cs, analysis="ac", freqs=f, gmin=0.0)
    w = 2 * np.pi * f[None, :]
    H = 1.0 / (1.0 + 1j * w * r[:, None] * 1e-9)
    assert np.abs(sens.solution.array(cs.sys.node_out) - H).max() < 1e-12
    d = sens.array(cs.sys.node_out, "r")
    assert d.shape == (3, 9) and np.abs(d - (-1j * w * 1e-9 * H * H)).max() <= 1e-9 * np.abs(w * 1e-9 * H * H).max()
    sens = sensitivities_(cs, analysis="noise", freqs=f, temp=27.0, gmin=0.0)
    S = sens.solution.array(cs.sys.node_out)
    kT4 = 4 * 1.380649e-23 * 300.15
    assert np.allclose(S, kT4 * r[:, None] * np.abs(H) ** 2, rtol=1e-6, atol=0)
    x = (w * r[:, None] * 1e-9) ** 2
    want = S / r[:, None] * (1 - x) / (1 + x)          # d/dr [ r / (1 + (w r c)^2) ]
    assert np.abs(sens.array(cs.sys.node_out, "r") - want).max() <= 1e-7 * np.abs(S / r[:, None]).max()
    assert sens.point(1, cs.sys.node_out, "r").shape == (9,)



@pytest.mark.gpu
def test_sensitivities_bsimcmg_inverter_gain_gpu():
    # Transistor level, two independent routes to the same number: d q / d vin from the difference stencil over Newton-solved
    # operating points, and the low-frequency AC gain from the generated Jacobians (cb_ac) at the same points.
    from cedarsim.jl_b200 import circuits
    from cedarsim.jl_b200.sweeps import ac_
    cs = CircuitSweep(circuits.BSIMCMG_INVERTER_VIN_DECK, ProductSweep(**{"vin": np.linspace(0.1, 0.9, 17), "mneg.nfin": [1.0, 2.0, 3.0]}),
                      outputs=["q"], host=True)
    sens = sensitivities_(cs, rel_step=1e-3)
    assert (sens.retcodes == "Success").all()
    gain = sens.array(cs.sys.node_q, "vin")
    H = ac_(cs, [1.0]).array(cs.sys.node_q)[..., 0]
    assert gain.shape == H.shape == (17, 3)
    assert gain.min() < -3.0 and gain.max() < 0.0        # an inverting stage with gain in the transition region
    assert np.abs(H.imag).max() < 1e-6 * np.abs(H.real).max()
    assert np.abs(gain - H.real).max() <= 1e-5 * np.abs(H.real).max()
    # more nFET fins pull the output down: d q / d nfin < 0 everywhere, largest in the transition region
    dn = sens.array(cs.sys.node_q, "mneg.nfin")
    assert dn.max() < 0.0 and np.abs(dn).max() > 0.05


# ---- direct method (cb_sens_dc): one solve + a pair of triangular solves per parameter with the stored Newton factors
@pytest.mark.gpu
def test_direct_sensitivities_two_resistor_known_answer_gpu():   # test/sensitivity.jl:31-41,58-67: dR1 == -dR2 == -out/2
    cs = CircuitSweep(TWO_R, ProductSweep(R1=[1.0], R2=[1.0]))
    sens = sensitivities_(cs, method="direct")
    out = sens.solution.array(cs.sys.node_out)
    d1, d2 = sens.array(cs.sys.node_out, "R1"), sens.array(cs.sys.node_out, "R2")
    assert sens.retcodes[0, 0] == "Success" and abs(out[0, 0] - 0.5) < 1e-12
    assert abs(d1[0, 0] + d2[0, 0]) < 1e-9 and abs(d1[0, 0] + out[0, 0] / 2) < 1e-8


@pytest.mark.gpu
def test_direct_sensitivities_match_closed_forms_and_the_stencil_gpu(tmp_path):
    R1v, R2v = np.arange(100.0, 1001, 100), np.arange(100.0, 1001, 100)
    cs = CircuitSweep(TWO_R, ProductSweep(R1=R1v, R2=R2v))
    sens = sensitivities_(cs, wrt=["R2", "R1"], method="direct")
    R1, R2 = np.meshgrid(R1v, R2v, indexing="ij")
    assert np.allclose(sens.array(cs.sys.node_out, "R2"), R1 / (R1 + R2) ** 2, rtol=1e-7, atol=0)
    assert np.allclose(sens.array(cs.sys.node_out, "R1"), -R2 / (R1 + R2) ** 2, rtol=1e-7, atol=0)
    assert np.allclose(sens.array(cs.sys.v.I, "R2"), 1.0 / (R1 + R2) ** 2, rtol=1e-7, atol=0)
    # a Newton-solved nonlinear point: V -- R -- diode, implicit differentiation
    (tmp_path / "dio.va").write_text(DIODE_VA)
    text = f"* diode\n.hdl \"{tmp_path / 'dio.va'}\"\n.param r=1k vin=1\nV1 in 0 'vin'\nR1 in d 'r'\nX1 d 0 dio\n"
    cs = CircuitSweep(text, ProductSweep(r=[500.0, 1e3, 2e3], vin=[0.8, 1.0, 2.0]), outputs=["d"])
    direct = sensitivities_(cs, method="direct")
    stencil = sensitivities_(cs)
    v = direct.solution.array(cs.sys.node_d)
    R, VIN = np.meshgrid([500.0, 1e3, 2e3], [0.8, 1.0, 2.0], indexing="ij")
    gd = 1e-14 * np.exp(v / 0.025852) / 0.025852
    assert np.allclose(direct.array(cs.sys.node_d, "r"), -(VIN - v) / R ** 2 / (1 / R + gd), rtol=1e-6, atol=0)
    assert np.allclose(direct.array(cs.sys.node_d, "vin"), (1 / R) / (1 / R + gd), rtol=1e-6, atol=0)
    for name in ("r", "vin"):
        assert np.allclose(direct.array(cs.sys.node_d, name), stencil.array(cs.sys.node_d, name), rtol=1e-6, atol=0)
    assert direct.stats["newton_iters"] < stencil.solution.stats["newton_iters"] / 5      # B points solved, not B (1 + 4 n)


@pytest.mark.gpu
def test_direct_sensitivities_bsimcmg_inverter_gpu(host_bsimcmg):
    """Transistor level: d V(q) / d (supply, nFET fin count) of the BSIM-CMG inverter near its switching threshold, direct
    method (value-only device evaluations at fixed x*, stored factors) against the stencil of re-solved points."""
    from cedarsim.jl_b200 import circuits
    cs = CircuitSweep(circuits.BSIMCMG_INVERTER_VIN_DECK, ProductSweep(**{"vin": np.linspace(0.35, 0.65, 7), "mneg.nfin": [2.0, 3.0]}),
                      outputs=["q", "vvdd.i"])
    direct = sensitivities_(cs, method="direct")
    stencil = sensitivities_(cs)
    assert (direct.retcodes == "Success").all()
    for ref in (cs.sys.node_q, cs.sys.vvdd.I):
        for name in ("vin", "mneg.nfin"):
            a, b = direct.array(ref, name), stencil.array(ref, name)
            assert np.all(np.abs(a - b) <= 1e-5 * np.abs(b) + 1e-9 * np.abs(b).max()), (name, np.abs(a - b).max(), np.abs(b).max())
    assert np.abs(direct.array(cs.sys.node_q, "vin")).max() > 3.0      # gain of the inverter


def test_direct_method_argument_errors_without_gpu():
    cs = CircuitSweep(TWO_R, ProductSweep(R1=[1.0], R2=[1.0]))
    with pytest.raises(ValueError):
        sensitivities_(cs, analysis="tran", method="direct")
    with pytest.raises(ValueError):
        sensitivities_(cs, method="adjoint")
