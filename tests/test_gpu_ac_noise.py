"""GPU parity of the small-signal path (cb_ac / cb_noise through the C ABI) against the CPU oracle, which is itself
pinned to the reference's ngspice tables (tests/test_oracle_ac_noise.py), and directly against those tables.

Tolerances: the DC operating points agree to <= 1e-9 V (parity bar of the north star); transfer functions and noise
densities are compared at rtol 1e-6 (the reference's own tolerance for these tables, test/ac.jl:148,237), the measured
differences are ~1e-10.
"""
import os

import numpy as np
import pytest

from cedarsim.jl_b200 import circuits, engine, netlist
from cedarsim.jl_b200.sweeps import CircuitSweep, ProductSweep, Sweep, ac_, acdec, noise_
from oracle import orc
from test_oracle_ac_noise import BSIMCMG_INVERTER, BUTTERWORTH, golden, isapprox

pytestmark = pytest.mark.gpu


def test_butterworth_ac_and_noise_sweep():
    # sweep the split load resistor: AC response and noise change, analytic answers known for res = 1
    res = np.array([0.5, 1.0, 2.0, 4.0])
    cs = CircuitSweep(BUTTERWORTH, Sweep(res=res), outputs=["vout", "vin", "l1.i"])
    f = acdec(20, 0.01, 10)
    sol = ac_(cs, f, temp=23.0, gmin=0.0)
    assert sol.array(cs.sys.node_vout).shape == (4, len(f))
    s = 2j * np.pi * f
    H = 1.0 / ((s + 1.0) * (s * s + s + 1.0))
    assert np.abs(sol.point(1, cs.sys.node_vout) - H).max() < 1e-12
    assert np.abs(sol.point(1, cs.sys.node_vin) - 1.0).max() < 1e-15
    yo, so = orc.ac(cs.flat.fc, f, cs.flat.params, opts=orc.default_options(temp=23.0, gmin=0.0))
    assert so.max() == 0 and sol.status.max() == 0
    assert np.abs(sol.y - yo).max() <= 1e-12 * np.abs(yo).max()
    nz = noise_(cs, f, temp=23.0, gmin=0.0)
    po, _ = orc.noise(cs.flat.fc, f, cs.flat.params, opts=orc.default_options(temp=23.0, gmin=0.0))
    assert np.all(np.abs(nz.y - po) <= 1e-10 * np.abs(po) + 1e-300)
    ng = golden("ngspice_noise_butterworth.txt")
    assert isapprox(np.sqrt(nz.point(1, cs.sys.node_vout)), ng[:, 1], 1e-6)       # test/ac.jl:148


def test_bsimcmg_inverter_noise_vs_ngspice_table(host_bsimcmg):
    """test/ac.jl:161-237 on the GPU: 61 frequencies, 1 kHz .. 1 PHz, rtol 1e-6 against ngspice."""
    fl = netlist.flatten(netlist.parse_netlist(BSIMCMG_INVERTER), None, outputs=["q"], host=True)
    ng = golden("ngspice_noise_bsimcmg_inverter.txt")
    plan = engine.Circuit(fl.fc, fl.models).plan(1)
    plan.set_params(None)
    psd, st, stats = plan.noise(ng[:, 0])
    assert st.max() == 0
    mine = np.sqrt(psd[0, :, 0])
    assert isapprox(mine, ng[:, 1], 1e-6)
    assert np.abs(mine / ng[:, 1] - 1).max() < 1e-7


def test_bsimcmg_inverter_ac_noise_sweep_vs_oracle(host_bsimcmg):
    """ProductSweep over the input bias and the nFET fin count: 16 x 4 operating points x 31 frequencies."""
    cs = CircuitSweep(circuits.BSIMCMG_INVERTER_VIN_DECK, ProductSweep(**{"vin": np.linspace(0.1, 0.9, 16), "mneg.nfin": [1.0, 2.0, 3.0, 4.0]}),
                      outputs=["q", "vvdd.i"], host=True)
    f = acdec(2, 1e3, 1e18)
    fc, P = cs.flat.fc, cs.flat.params
    sol = ac_(cs, f)
    yo, so = orc.ac(fc, f, P)
    assert sol.status.max() == 0 and so.max() == 0
    assert np.all(np.abs(sol.y - yo) <= 1e-6 * np.abs(yo) + 1e-30)
    gain = np.abs(sol.array(cs.sys.node_q)[:, :, 0])
    assert gain.max() > 3.0 and gain[0].max() < 1.0       # high gain near the switching threshold, none at the rails
    nz = noise_(cs, f)
    po, _ = orc.noise(fc, f, P)
    assert np.all(np.isfinite(nz.y)) and np.all(nz.y > 0)
    assert np.all(np.abs(nz.y - po) <= 1e-6 * po)


def test_dff_noise_vs_oracle(host_bsimcmg):
    """85 unknowns, 30 FETs, 657 LU entries: the complex LU in 190 KB of shared memory per CTA."""
    fc, ms = circuits.dff(host=True, sweep=True)
    fc.set_outputs(["q", "d_neg", "vvdd.i"])
    next(w for d, w in ((d, fc.waves[d.wave]) for d in fc.devices if d.name == "vd")).ac = 1.0   # VD ... AC 1
    B = 24
    P = circuits.dff_mc_params(fc, B)
    from cedarsim.jl_b200.flat import nodeset_vector
    x0 = nodeset_vector(fc, dict(q=0.0, q_neg=0.7, net0=0.0, net7=0.0, vdd=0.7, clkn=0.7, ncki=0.0, cki=0.7))
    f = acdec(1, 1e3, 1e12)
    plan = engine.Circuit(fc, ms).plan(B)
    plan.set_params(P)
    plan.set_x0(x0)
    psd, st, _ = plan.noise(f)
    y, st2, _ = plan.ac(f)
    orc.set_x0(x0)
    po, so = orc.noise(fc, f, P)
    yo, _ = orc.ac(fc, f, P)
    orc.set_x0(None)
    assert st.max() == 0 and so.max() == 0 and st2.max() == 0
    assert np.all(np.abs(psd - po) <= 1e-6 * po + 1e-40)
    assert np.all(np.abs(y - yo) <= 1e-6 * np.abs(yo) + 1e-20)


def test_noise_twice_on_one_plan_with_different_temperature(host_bsimcmg):
    """The noise variant's per-instance cache depends on temp: a second cb_noise on the same plan at another temperature
    must refill it (round-1 advisor finding: it kept the old-temperature cache)."""
    fl = netlist.flatten(netlist.parse_netlist(BSIMCMG_INVERTER), None, outputs=["q"], host=True)
    f = acdec(2, 1e3, 1e9)
    plan = engine.Circuit(fl.fc, fl.models).plan(1)
    plan.set_params(None)
    res = {}
    for temp in (27.0, 85.0, 27.0):
        psd, st, _ = plan.noise(f, engine.default_options(temp=temp))
        assert st.max() == 0
        po, _ = orc.noise(fl.fc, f, None, opts=orc.default_options(temp=temp))
        assert np.all(np.abs(psd - po) <= 1e-6 * po), temp
        res.setdefault(temp, []).append(psd)
    assert np.array_equal(res[27.0][0], res[27.0][1])
    assert np.abs(res[85.0][0] / res[27.0][0] - 1).max() > 0.05     # the two temperatures do differ
