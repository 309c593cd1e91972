"""BASELINE.json configs 2, 3, 4 and 5 at their FULL sizes on the GPU (through the C ABI), checked by
(i) a seeded subset against the CPU oracle at the parity tolerances and (ii) size-independent properties of the
domain (all points converge, I-V monotone, zero current at zero bias, logic levels at the reference's sample times).
Config 3 is also what bench.py times (its `parity_check` key); GF180 / BSIM4 are not in the reference tree, so
the same topologies run on BSIM-CMG 107 + ASAP7 cards."""
import numpy as np
import pytest

from cedarsim.jl_b200 import circuits, engine
from cedarsim.jl_b200.flat import params_matrix
from oracle import orc

pytestmark = pytest.mark.gpu


def test_config4_fet_iv_1m_bias_points(host_bsimcmg):
    # ProductSweep(vg.dc = linspace(0, 0.9, 1024), vd.dc = linspace(0, 0.9, 1024)), column-major
    fc, ms = circuits.fet_iv(host=host_bsimcmg)
    n = 1024
    vg, vd = np.meshgrid(np.linspace(0, 0.9, n), np.linspace(0, 0.9, n), indexing="ij")
    P = params_matrix([vg.ravel(order="F"), vd.ravel(order="F")])
    B = P.shape[1]
    assert B == 1048576
    plan = engine.Circuit(fc, ms).plan(B)
    plan.set_params(P)
    x, xf, st, stats = plan.dc()
    plan.close()
    assert st.max() == 0
    idrain = -x[0].reshape((n, n), order="F")      # current into the drain = -I(Vd); [i_vg, i_vd]
    assert np.abs(idrain[:, 0]).max() < 1e-10       # vd = 0: only gate-to-drain tunnelling current (pA)
    assert idrain.min() > -1e-10
    assert np.all(np.diff(idrain, axis=1) > -1e-12)  # monotone in vd at fixed vg
    assert np.all(np.diff(idrain, axis=0) > -1e-12)  # monotone in vg at fixed vd
    assert 1e-5 < idrain[-1, -1] < 1e-3              # on-current of a 3-fin ASAP7 nFET, order of 100 uA
    # seeded subset against the oracle, DC tolerance 1e-9 V (currents: 1e-9 relative to the largest)
    rng = np.random.default_rng(4)
    sel = np.sort(rng.choice(B, 512, replace=False))
    xo, xfo, so, _ = orc.dc(fc, np.ascontiguousarray(P[:, sel]))
    assert so.max() == 0
    nv = fc.n_nodes
    assert np.abs(xf[:nv, sel] - xfo[:nv]).max() < 1e-9
    assert np.abs(xf[nv:, sel] - xfo[nv:]).max() < 1e-12 + 1e-9 * np.abs(xfo[nv:]).max()


def test_config2_inverter_product_sweep_65536(host_bsimcmg):
    # vvdd.dc (16) x xneg.nfin (64, 1x..4x) x xneg.l (64, 1x..3x), fixed-step trapezoidal,
    # dt = 50 ps, 8 000 steps over 400 ns, Q and D saved every 1 ns (S = 401)
    fc, ms = circuits.inverter(host=host_bsimcmg)
    vdd, nfin, ln = np.meshgrid(np.linspace(0.8 * 0.7, 1.2 * 0.7, 16), np.linspace(3.0, 12.0, 64), np.linspace(21e-9, 63e-9, 64),
                                indexing="ij")
    B = vdd.size
    assert B == 65536
    P = np.zeros((3, B))
    P[fc.param_names.index("vvdd.dc")] = vdd.ravel(order="F")
    P[fc.param_names.index("xneg.nfin")] = nfin.ravel(order="F")
    P[fc.param_names.index("xneg.l")] = ln.ravel(order="F")
    ts = np.linspace(0, 4e-7, 401)
    kw = dict(fixed_step=1, dt=50e-12)
    plan = engine.Circuit(fc, ms).plan(B)
    plan.set_params(P)
    y, st, stats = plan.tran(0.0, 4e-7, ts, engine.default_options(**kw))
    plan.close()
    assert st.max() == 0
    assert stats["steps_accepted"] == 8000 * B
    # logic levels at the reference's sample times (test/inverter.jl:40-50: D/Q at 0.5, 1.5, 2.5, 3.5e-7 s)
    q, d = y[0], y[1]
    vsup = P[fc.param_names.index("vvdd.dc")]
    for t, d_high in ((0.4e-7, False), (1.0e-7, True), (2.0e-7, False), (3.5e-7, True)):
        k = int(round(t / 4e-7 * 400))
        assert np.abs(d[k] - (0.7 if d_high else 0.0)).max() < 1e-9
        want_q = 0.0 * vsup if d_high else vsup
        assert np.abs(q[k] - want_q).max() < 5e-3, (t, np.abs(q[k] - want_q).max())
    # seeded subset against the oracle at the fixed-step parity tolerance
    rng = np.random.default_rng(2)
    sel = np.sort(rng.choice(B, 24, replace=False))
    yo, so, _ = orc.tran(fc, 0.0, 4e-7, ts, params=np.ascontiguousarray(P[:, sel]), opts=orc.default_options(**kw), nthreads=8)
    assert so.max() == 0
    err = np.abs(y[:, :, sel] - yo)
    assert np.all(err <= 1e-6 * np.abs(yo) + 1e-9), err.max()


def test_config5_corner_temperature_mismatch_131072(host_bsimcmg):
    """SURVEY 8(d) config 5 stand-in at full size: 4 corners x 32 temperatures (-40 .. 125 C, SimSpec.temp as a swept
    column) x 1 024 mismatch draws = 131 072 transient instances of one cell through the sweep API.  sky130 / BSIM4 are
    not in the reference tree (parity unpinned in the reference itself); the cell is the BSIM-CMG inverter, corners act
    through the model's variability handles DELVTRAND / U0MULT (circuits.CONFIG5_DECK)."""
    from cedarsim.jl_b200.sweeps import CircuitSweep, tran_
    cs = CircuitSweep(circuits.CONFIG5_DECK, circuits.config5_sweep(), outputs=["q", "d", "vvdd.i"], host=True)
    assert cs.shape == (4, 32, 1024) and len(cs) == 131072
    ts = np.linspace(0, 2.5e-9, 251)
    sols = tran_(cs, (0.0, 2.5e-9), saveat=ts, reltol=1e-3)
    assert sols.status.max() == 0
    q = sols.array(cs.sys.node_q)                      # (4, 32, 1024, 251)
    assert np.abs(q[..., 40] - 0.7).max() < 1e-3 and np.abs(q[..., 110]).max() < 1e-3 and np.abs(q[..., 220] - 0.7).max() < 1e-3
    # static supply current with the input low (sub-threshold leakage of the nFET): grows with temperature at every
    # corner / draw, and fast corners leak more than slow ones
    leak = -sols.array(cs.sys.vvdd.I)[..., 0]          # (4, 32, 1024)
    assert leak.min() > 0
    assert np.all(np.diff(leak, axis=1) > 0)
    assert np.all(leak[1] > leak[0]) and np.all(leak[0] > leak[2])      # ff > tt > ss
    # seeded subset against the oracle in the fixed-step comparison mode (1e-6 rel / 1e-9 abs)
    rng = np.random.default_rng(5)
    sel = np.sort(rng.choice(len(cs), 32, replace=False))
    fc, P = cs.flat.fc, np.ascontiguousarray(cs.flat.params[:, sel])
    kw = dict(fixed_step=1, dt=1e-12, temp=cs.flat.options["temp"])
    plan = engine.Circuit(fc, cs.flat.models).plan(len(sel))
    plan.set_params(P)
    y, st, _ = plan.tran(0.0, 2.5e-9, ts, engine.default_options(**kw))
    plan.close()
    yo, so, _ = orc.tran(fc, 0.0, 2.5e-9, ts, params=P, opts=orc.default_options(**kw), nthreads=8)
    assert st.max() == 0 and so.max() == 0
    err = np.abs(y - yo)
    # node voltages at the north-star bar
    assert np.all(err[:2] <= 1e-6 * np.abs(yo[:2]) + 1e-9), err[:2].max()
    # supply current: i = ... + C dv/dt, so a node-voltage difference inside that bar (dv <= 1e-9 V) shows as up to
    # 2 C dv / dt ~ 1e-11 A at dt = 1 ps and ~10 fF of node capacitance; asserted at 1e-6 of the waveform's peak (~4e-5 A)
    ipk = np.abs(yo[2]).max()
    assert np.all(err[2] <= 1e-6 * np.abs(yo[2]) + 1e-6 * ipk), (err[2].max(), ipk)


def test_config3_dff_monte_carlo_16384(host_bsimcmg):
    """SURVEY 8(d) config 3 at full size with the bench options: 16 384 Monte-Carlo instances of the 30-FET DFF, adaptive
    trapezoidal, 0 .. 600 ns.  Every point converges, Q follows the reference's known pattern (test/gf180_dff.jl:29-33:
    0, 0, VDD, VDD, VDD at 1.5 / 2.5 / 4.5 / 5.5 / 6.0e-7 s), and a seeded subset agrees with the oracle run with the same
    options within the LTE tolerance away from the edges."""
    from helpers import x0_from
    nodeset = dict(q=0.0, q_neg=0.7, net0=0.0, net7=0.0, vdd=0.7, clkn=0.7, ncki=0.0, cki=0.7)
    fc, ms = circuits.dff(host=host_bsimcmg)
    B = 16384
    P = circuits.dff_mc_params(fc, B)
    x0 = x0_from(fc, nodeset)
    ts = np.linspace(0, 6e-7, 61)
    kw = dict(reltol=1e-4, vabstol=1e-6, iabstol=1e-12, nr_reltol=1e-5, nr_vabstol=1e-7, nr_iabstol=1e-13, nr_rate_test=1)
    plan = engine.Circuit(fc, ms).plan(B)
    plan.set_params(P)
    plan.set_x0(x0)
    y, st, stats = plan.tran(0.0, 6e-7, ts, engine.default_options(value_rounds=2, **kw))
    plan.close()
    assert st.max() == 0
    q = y[0]   # outputs of circuits.dff: node_q, node_d
    for t, want in ((1.5e-7, 0.0), (2.5e-7, 0.0), (4.5e-7, 0.7), (5.5e-7, 0.7), (6.0e-7, 0.7)):
        k = int(round(t / 6e-7 * 60))
        assert np.abs(q[k] - want).max() < 1e-4, (t, np.abs(q[k] - want).max())   # the reference's atol
    # work per point as DESIGN.md section 5 quotes it
    assert 1500 < stats["steps_accepted"] / B < 2500 and stats["newton_iters"] / B < 6000
    rng = np.random.default_rng(3)
    sel = np.sort(rng.choice(B, 16, replace=False))
    orc.set_x0(x0)
    try:
        yo, so, _ = orc.tran(fc, 0.0, 6e-7, ts, params=np.ascontiguousarray(P[:, sel]), opts=orc.default_options(**kw), nthreads=8)
    finally:
        orc.set_x0(None)
    assert so.max() == 0
    settled = np.abs(np.gradient(yo, axis=1)).max(axis=(0, 2)) < 1e-3
    assert np.abs(y[:, :, sel] - yo)[:, settled, :].max() < 2e-3
