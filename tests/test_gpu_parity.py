"""GPU parity tests: CUDA engine (through the C ABI) vs the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): DC node voltages within 1e-9 V; transient within 1e-6
relative / 1e-9 V absolute at shared output times in fixed-step comparison mode.  Adaptive-step
runs choose their own step sequences from LTE estimates, so they are compared at the LTE tolerance.
"""
import numpy as np
import pytest

from cedarsim.jl_b200 import circuits, models
from cedarsim.jl_b200.flat import FlatCircuit, Wave, W_PULSE, W_PWL, W_SIN, params_matrix

from helpers import run_dc_both, run_tran_both, x0_from

pytestmark = pytest.mark.gpu

DC_VTOL = 1e-9


def assert_tran_close(yg, yo, rtol=1e-6, atol=1e-9):
    err = np.abs(yg - yo)
    tol = rtol * np.abs(yo) + atol
    assert np.all(err <= tol), f"max err {err.max():.3e}, worst excess {(err - tol).max():.3e}"


def test_two_resistor_dc_sweep():
    # reference test/sweep.jl:326-340: I(V) = -1/(R1+R2) to 1e-7 on a 20x20 ProductSweep
    fc = circuits.two_resistor()
    r1, r2 = np.meshgrid(np.arange(100, 2001, 100.0), np.arange(100, 2001, 100.0), indexing="ij")
    P = params_matrix([r1.ravel(order="F"), r2.ravel(order="F")])
    (xg, xfg, sg, _), (xo, xfo, so, _) = run_dc_both(fc, [], P)
    assert sg.max() == 0 and so.max() == 0
    assert np.abs(xg[0] - (-1.0 / (P[0] + P[1]))).max() < 1e-12
    assert np.abs(xfg - xfo).max() < DC_VTOL


def test_rc_pulse_fixed_step():
    fc = FlatCircuit()
    fc.vsource("V", "in", "0", Wave(W_PULSE, v=[0, 1, 1e-8, 1e-9, 1e-9, 2e-6, 4e-6]))
    fc.resistor("R", "in", "out", fc.param("r"))
    fc.capacitor("C", "out", "0", 1e-9)
    fc.set_outputs(["out", "v.i"])
    P = params_matrix([np.linspace(500.0, 2000.0, 37)])
    ts = np.linspace(0, 5e-6, 101)
    for method in (0, 1, 2):
        (yg, sg, _), (yo, so, _) = run_tran_both(fc, [], P, 0.0, 5e-6, ts, fixed_step=1, dt=1e-8, method=method)
        assert sg.max() == 0 and so.max() == 0
        assert_tran_close(yg, yo)


def test_rlc_sin_adaptive():
    fc = FlatCircuit()
    fc.vsource("V", "in", "0", Wave(W_SIN, v=[0.0, 1.0, 1e6]))
    fc.resistor("R", "in", "a", 50.0)
    fc.inductor("L", "a", "out", fc.param("l"))
    fc.capacitor("C", "out", "0", 1e-9)
    fc.resistor("RL", "out", "0", 1e3)
    fc.set_outputs(["out", "l.i"])
    P = params_matrix([np.linspace(1e-6, 2e-5, 16)])
    ts = np.linspace(0, 5e-6, 201)
    (yg, sg, stg), (yo, so, sto) = run_tran_both(fc, [], P, 0.0, 5e-6, ts, reltol=1e-4)
    assert sg.max() == 0 and so.max() == 0
    assert np.abs(yg - yo).max() < 5e-3  # both within the LTE tolerance of the same algorithm
    assert abs(stg["steps_accepted"] - sto["steps_accepted"]) <= 0.05 * sto["steps_accepted"] + 5


def test_pwl_current_source_fixed():
    # shape of reference test/transients.jl:17-63 (PWL current into R, analytic at every time point)
    fc = FlatCircuit()
    fc.isource("I", "0", "out", Wave(W_PWL, t=[0, 1e-3, 2e-3, 3e-3], y=[0.0, 1e-3, 1e-3, 0.0]))
    fc.resistor("R", "out", "0", fc.param("r"))
    fc.set_outputs(["out"])
    P = params_matrix([np.array([10.0, 100.0, 1e3, 1e4])])
    ts = np.linspace(0, 3e-3, 61)
    (yg, sg, _), (yo, so, _) = run_tran_both(fc, [], P, 0.0, 3e-3, ts, fixed_step=1, dt=5e-5)
    assert sg.max() == 0
    exact = np.interp(ts, [0, 1e-3, 2e-3, 3e-3], [0, 1e-3, 1e-3, 0])[None, :] * P[0][:, None]
    assert np.abs(yg[0].T - exact).max() < 1e-9
    assert_tran_close(yg, yo)


def test_bsimcmg_fet_iv_dc(host_bsimcmg):
    # BASELINE config 4 at reduced size: ASAP7 nmos_lvt I-V grid
    fc, ms = circuits.fet_iv(host=host_bsimcmg)
    vg, vd = np.meshgrid(np.linspace(0, 0.9, 24), np.linspace(0, 0.9, 24), indexing="ij")
    P = params_matrix([vg.ravel(order="F"), vd.ravel(order="F")])
    (xg, xfg, sg, _), (xo, xfo, so, _) = run_dc_both(fc, ms, P)
    assert sg.max() == 0 and so.max() == 0
    nv = fc.n_nodes
    assert np.abs(xfg[:nv] - xfo[:nv]).max() < DC_VTOL
    assert np.abs(xfg[nv:] - xfo[nv:]).max() < 1e-12 + 1e-9 * np.abs(xfo[nv:]).max()


def test_bsimcmg_inverter_tran_fixed(host_bsimcmg):
    # BASELINE config 2 at reduced size: vdd x nfin x l sweep, fixed-step trapezoidal
    fc, ms = circuits.inverter(host=host_bsimcmg, tscale=0.01)
    vdd, nfin, ln = np.meshgrid(np.linspace(0.56, 0.84, 3), np.linspace(2, 6, 3), np.linspace(21e-9, 40e-9, 3), indexing="ij")
    P = np.zeros((3, 27))
    P[fc.param_names.index("vvdd.dc")] = vdd.ravel(order="F")
    P[fc.param_names.index("xneg.nfin")] = nfin.ravel(order="F")
    P[fc.param_names.index("xneg.l")] = ln.ravel(order="F")
    ts = np.linspace(0, 4e-9, 81)
    (yg, sg, _), (yo, so, _) = run_tran_both(fc, ms, P, 0.0, 4e-9, ts, fixed_step=1, dt=2e-12)
    assert sg.max() == 0 and so.max() == 0
    assert_tran_close(yg, yo)


def test_bsimcmg_inverter_tran_value_rounds(host_bsimcmg):
    # throughput options: chord iterations in value-only rounds (+ IDA-style rate test) must reach the same
    # solutions as the oracle's plain full Newton.  Plain acceptance test: same tolerance as above; the rate test
    # stops an iteration earlier, with the charges updated to first order (q + C dx), hence the slightly looser bound.
    fc, ms = circuits.inverter(host=host_bsimcmg, tscale=0.01)
    vdd, nfin, ln = np.meshgrid(np.linspace(0.56, 0.84, 3), np.linspace(2, 6, 3), np.linspace(21e-9, 40e-9, 3), indexing="ij")
    P = np.zeros((3, 27))
    P[fc.param_names.index("vvdd.dc")] = vdd.ravel(order="F")
    P[fc.param_names.index("xneg.nfin")] = nfin.ravel(order="F")
    P[fc.param_names.index("xneg.l")] = ln.ravel(order="F")
    ts = np.linspace(0, 4e-9, 81)
    kw = dict(fixed_step=1, dt=2e-12)
    (yg, sg, stg), (yo, so, _) = run_tran_both(fc, ms, P, 0.0, 4e-9, ts, engine_only=dict(value_rounds=1), **kw)
    assert sg.max() == 0 and so.max() == 0
    assert stg["value_rounds"] > 0 and 0 < stg["full_iters"] < stg["newton_iters"]
    assert_tran_close(yg, yo)
    (yg, sg, stg), _ = run_tran_both(fc, ms, P, 0.0, 4e-9, ts, engine_only=dict(value_rounds=1, nr_rate_test=1), **kw)
    assert sg.max() == 0 and stg["value_rounds"] > 0
    assert_tran_close(yg, yo, rtol=2e-6, atol=1e-8)


def test_bsimcmg_dff_adaptive_throughput_options(host_bsimcmg):
    # the bench configuration (bench.py): adaptive, Newton tolerance tied to the LTE tolerance, rate test, value rounds
    fc, ms = circuits.dff(host=host_bsimcmg)
    P = circuits.dff_mc_params(fc, 32)
    x0 = x0_from(fc, DFF_NODESET)
    ts = np.linspace(0, 6e-7, 61)
    kw = dict(reltol=1e-4, vabstol=1e-6, iabstol=1e-12)
    fast = dict(nr_reltol=1e-5, nr_vabstol=1e-7, nr_iabstol=1e-13, nr_rate_test=1, value_rounds=1)
    (yg, sg, stg), (yo, so, sto) = run_tran_both(fc, ms, P, 0.0, 6e-7, ts, x0=x0, engine_only=fast, **kw)
    assert sg.max() == 0 and so.max() == 0
    # both runs are within their LTE tolerance of the true waveform; edges are steep, so compare away from them
    settled = np.abs(np.gradient(yo, axis=1)).max(axis=(0, 2)) < 1e-3
    assert np.abs(yg - yo)[:, settled, :].max() < 2e-3
    assert stg["newton_iters"] < sto["newton_iters"]


DFF_NODESET = dict(q=0.0, q_neg=0.7, net0=0.0, net7=0.0, vdd=0.7, clkn=0.7, ncki=0.0, cki=0.7)


def test_bsimcmg_dff_mc_fixed(host_bsimcmg):
    # BASELINE config 3 at reduced size and span: 30-FET DFF, Monte-Carlo L/NFIN, fixed step
    fc, ms = circuits.dff(host=host_bsimcmg)
    P = circuits.dff_mc_params(fc, 16)
    x0 = x0_from(fc, DFF_NODESET)
    ts = np.linspace(0, 5.2e-8, 53)
    (yg, sg, stg), (yo, so, sto) = run_tran_both(fc, ms, P, 0.0, 5.2e-8, ts, x0=x0, fixed_step=1, dt=25e-12)
    assert sg.max() == 0 and so.max() == 0
    assert_tran_close(yg, yo)
    # borderline convergence decisions may differ by an iteration between FMA and non-FMA arithmetic
    assert abs(stg["newton_iters"] - sto["newton_iters"]) <= 5e-3 * sto["newton_iters"]


def test_bsimcmg_dff_adaptive_known_answers(host_bsimcmg):
    # full span, adaptive: Q follows the reference's known pattern 0,0,VDD,VDD,VDD
    # (test/gf180_dff.jl:29-33 asserts 0,0,5,5,5 V on the GF180 deck; same topology at 0.7 V here)
    fc, ms = circuits.dff(host=host_bsimcmg)
    P = circuits.dff_mc_params(fc, 64)
    x0 = x0_from(fc, DFF_NODESET)
    ts = np.array([1.5e-7, 2.5e-7, 4.5e-7, 5.5e-7, 6.0e-7])
    (yg, sg, stg), (yo, so, sto) = run_tran_both(fc, ms, P, 0.0, 6e-7, ts, x0=x0, reltol=1e-3)
    assert sg.max() == 0 and so.max() == 0
    want = np.array([0, 0, 0.7, 0.7, 0.7])[:, None]
    assert np.abs(yg[0] - want).max() < 1e-3
    assert np.abs(yg - yo).max() < 1e-3


def test_verilog_a_voltage_branches():
    """V() <+ branches and I() probes (branch-current unknowns of Verilog-A devices) on the GPU: DC known answers and a
    fixed-step RL transient with a Verilog-A inductor against the oracle (decks of tests/test_va_compiler.py)."""
    import os
    from cedarsim.jl_b200 import netlist
    from test_va_compiler import VBRANCH_DC, VBRANCH_RL
    inc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "va")
    fl = netlist.flatten(netlist.parse_netlist(VBRANCH_DC, include_dirs=[inc]), host=True)
    fc = fl.fc
    (xg, xfg, sg, _), (xo, xfo, so, _) = run_dc_both(fc, fl.models, np.zeros((0, 4)))
    assert sg.max() == 0 and so.max() == 0
    assert np.abs(xfg - xfo).max() < DC_VTOL
    assert np.abs(xfg[fc.unknown("a")] - 1.5).max() < 1e-12 and np.abs(xfg[fc.unknown("x1.i(p,n)")] + 0.05).max() < 1e-13
    assert np.abs(xfg[fc.unknown("o")] - 0.5).max() < 1e-12
    L = np.linspace(1e-6, 4e-6, 64)
    fl = netlist.flatten(netlist.parse_netlist(VBRANCH_RL, include_dirs=[inc]), {"x1.l": L}, host=True)
    ts = np.linspace(0, 5e-8, 51)
    (yg, sg, _), (yo, so, _) = run_tran_both(fl.fc, fl.models, fl.params, 0.0, 5e-8, ts, fixed_step=1, dt=1e-10)
    assert sg.max() == 0 and so.max() == 0
    assert_tran_close(yg, yo, rtol=1e-6, atol=1e-9)
    k = fl.fc.unknown("x1.i(p,n)")
    assert np.all(np.diff(yg[k], axis=0) >= -1e-12) and np.all(yg[k, -1] < 0.01)      # the current rises towards V / R


@pytest.mark.parametrize("B", [1, 31, 33, 65, 129, 257])
def test_ragged_batch_sizes(host_bsimcmg, B):
    """Batch sizes that are not multiples of the warp, of the k_lu window (64) or of the eval CTA (128): tail lanes idle
    without touching memory, and the points that exist match the oracle at the fixed-step parity tolerance."""
    fc, ms = circuits.inverter(host=host_bsimcmg, tscale=0.01)
    P = np.zeros((3, B))
    P[fc.param_names.index("vvdd.dc")] = np.linspace(0.6, 0.8, B)
    P[fc.param_names.index("xneg.nfin")] = 3.0
    P[fc.param_names.index("xneg.l")] = np.linspace(21e-9, 30e-9, B)
    ts = np.linspace(0, 1e-9, 21)
    (yg, sg, _), (yo, so, _) = run_tran_both(fc, ms, P, 0.0, 1e-9, ts, fixed_step=1, dt=2e-12)
    assert yg.shape == yo.shape == (len(fc.outputs), 21, B)
    assert sg.max() == 0 and so.max() == 0
    assert_tran_close(yg, yo)


def test_source_stepping_rescues_the_operating_point(host_bsimcmg):
    """Newton limited to 6 iterations and no gmin ladder: the plain solve fails for 7 of the 8 points, source stepping
    (all independent sources ramped from 0 in 20 stages, each from the previous solution) finds their operating points.
    Same ladder in the engine's k_control and in the oracle."""
    fc, ms = circuits.inverter(host=host_bsimcmg, tscale=0.01)
    B = 8
    P = np.zeros((3, B))
    P[fc.param_names.index("vvdd.dc")] = np.linspace(0.6, 0.8, B)
    P[fc.param_names.index("xneg.nfin")] = 3.0
    P[fc.param_names.index("xneg.l")] = 21e-9
    kw = dict(max_newton_dc=6, gmin_steps=0)
    (xg, xfg, sg, stg), (xo, xfo, so, sto) = run_dc_both(fc, ms, P, source_steps=0, **kw)
    assert np.array_equal(sg, so) and (sg != 0).sum() == 7
    (xg, xfg, sg, stg), (xo, xfo, so, sto) = run_dc_both(fc, ms, P, source_steps=20, **kw)
    assert sg.max() == 0 and so.max() == 0
    assert stg["dc_source_stepped"] == sto["dc_source_stepped"] == 7
    assert np.abs(xfg - xfo).max() < DC_VTOL


def test_pivot_growth_monitor_flags_instead_of_losing_accuracy(host_bsimcmg):
    fc, ms = circuits.inverter(host=host_bsimcmg, tscale=0.01)
    B = 4
    P = np.zeros((3, B))
    P[fc.param_names.index("vvdd.dc")] = 0.7
    P[fc.param_names.index("xneg.nfin")] = 3.0
    P[fc.param_names.index("xneg.l")] = 21e-9
    from cedarsim.jl_b200 import engine
    plan = engine.Circuit(fc, ms).plan(B)
    plan.set_params(P)
    x, xf, st, stats = plan.dc(engine.default_options())
    assert st.max() == 0 and stats["pivot_fallbacks"] == 0
    # an absurdly small bound: every factorisation is flagged; without the repair pass no point may be reported as converged
    x, xf, st, stats = plan.dc(engine.default_options(pivot_growth_max=1e-6, gmin_steps=2, source_steps=2, max_newton_dc=10))
    assert st.min() != 0 and stats["pivot_fallbacks"] > 0
    plan.close()


def test_flagged_points_are_resolved_with_partial_pivoting(host_bsimcmg):
    """The slow path behind the pivot-growth monitor (cb_options.pivot_repair, k_lu's repair pass + lu_dense_pp): with an
    absurdly small growth bound EVERY full iteration of EVERY point is flagged and re-solved, dense, with partial
    pivoting.  Operating points and a fixed-step transient of the 30-FET DFF (85 unknowns) must come out as from the
    static-pivot factorisation (same Newton iterates up to rounding: 1e-9 V) and as from the CPU oracle."""
    from cedarsim.jl_b200 import engine
    from oracle import orc
    import bench
    fc, ms = circuits.dff(host=host_bsimcmg)
    B = 40                                  # two groups, the second one partly filled
    P = np.ascontiguousarray(circuits.dff_mc_params(fc, 64)[:, :B])
    x0 = x0_from(fc, bench.DFF_NODESET)
    ts = np.linspace(0.0, 4e-9, 41)
    kw = dict(fixed_step=1, dt=25e-12)
    plan = engine.Circuit(fc, ms).plan(B)
    plan.set_params(P)
    plan.set_x0(x0)
    xa, xfa, sa, _ = plan.dc(engine.default_options())
    xb, xfb, sb, stb = plan.dc(engine.default_options(pivot_growth_max=1e-6, pivot_repair=1))
    assert sa.max() == 0 and sb.max() == 0 and stb["pivot_fallbacks"] >= B
    assert np.abs(xfa[:fc.n_nodes] - xfb[:fc.n_nodes]).max() < 1e-9
    ya, sta, _ = plan.tran(0.0, 4e-9, ts, engine.default_options(**kw))
    yb, stb_, stats = plan.tran(0.0, 4e-9, ts, engine.default_options(pivot_growth_max=1e-6, pivot_repair=1, value_rounds=2, **kw))   # chord iterations asked for:
    plan.close()                                                                                               # rescued points must stay on full ones
    assert sta.max() == 0 and stb_.max() == 0
    assert np.all(np.abs(ya - yb) <= 1e-6 * np.abs(ya) + 1e-9), np.abs(ya - yb).max()
    assert stats["pivot_fallbacks"] > 0 and stats["full_iters"] == stats["newton_iters"]
    orc.set_x0(x0)
    yo, so, _ = orc.tran(fc, 0.0, 4e-9, ts, params=P, opts=orc.default_options(**kw), nthreads=8)
    orc.set_x0(None)
    assert so.max() == 0 and np.all(np.abs(yb - yo) <= 1e-6 * np.abs(yo) + 1e-9)
