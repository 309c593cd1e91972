"""Pin the CPU oracle's small-signal path against the golden vectors of the reference's own tests
(SURVEY.md 8(c) "small-signal tables: tightest pins on transistor Jacobians, rtol 1e-6").

* test/ac.jl:17-66   third-order Butterworth filter: AC response == 1 / ((s+1)(s^2+s+1)), observed source == 1,
                     inductor-voltage observable == s L3 H
* test/ac.jl:67-149  resistor thermal noise of the same filter vs the ngspice table and vs the analytic PSD
* test/ac.jl:161-237 BSIM-CMG 107 / ASAP7 inverter (test/bsimcmg/inverter_cmg_cedar.cir) output noise vs the ngspice
                     table, 61 rows from 1 kHz to 1 PHz.  This one pins, at rtol 1e-6, everything the transistor-level
                     hot path computes: the DC operating point, dI/dV and dQ/dV of the generated BSIM-CMG code, and the
                     flicker / thermal / shot noise sources of the model.

The tables live in tests/golden/ (extracted by scripts/make_golden_noise.py).  `isapprox` on vectors in the
reference is a norm-wise test; both that and the (stricter) element-wise maximum are asserted here.
"""
import os

import numpy as np
import pytest

from cedarsim.jl_b200 import circuits, netlist
from oracle import orc

HERE = os.path.dirname(os.path.abspath(__file__))

BUTTERWORTH = """*Third order low pass filter, butterworth, with w_c = 1
.param res=1

V1 vin 0 AC 1 SIN (0, 1, 0.15915494309189535)
L1 vin n1 1.5
C2 n1 0 1.3333333333333333
L3 n1 vout 0.5
* conceptually one resistor, split in two make a less trivial noise dss
R4 vout 0 '2*res'
R5 vout 0 '2*res'
"""

BSIMCMG_INVERTER = circuits.BSIMCMG_INVERTER_DECK   # test/bsimcmg/inverter_cmg_cedar.cir


def acdec(nd, fstart, fstop):   # src/ac.jl:286-303
    a, b = np.log10(fstart), np.log10(fstop)
    return 10.0 ** np.linspace(a, b, int(np.ceil((b - a) * nd)) + 1)


def golden(name):
    return np.loadtxt(os.path.join(HERE, "golden", name))


def isapprox(a, b, rtol):   # Julia isapprox on vectors: norm(a - b) <= rtol * max(norm(a), norm(b))
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


def test_butterworth_ac_response():   # test/ac.jl:37-66
    fl = netlist.flatten(netlist.parse_netlist(BUTTERWORTH), None, outputs=["vout", "vin", "n1"])
    f = acdec(20, 0.01, 10)
    y, st = orc.ac(fl.fc, f, opts=orc.default_options(temp=23.0, gmin=0.0))
    assert st.max() == 0
    s = 2j * np.pi * f
    H = 1.0 / ((s + 1.0) * (s * s + s + 1.0))
    assert np.abs(y[0, :, 0] - H).max() < 1e-12
    assert np.abs(y[1, :, 0] - 1.0).max() == 0.0            # directly observed source
    assert np.abs((y[2, :, 0] - y[0, :, 0]) - s * 0.5 * H).max() < 1e-12   # sys.l3.V = V(n1) - V(vout) = s L3 H


def test_butterworth_noise_vs_ngspice_and_analytic():   # test/ac.jl:67-149
    fl = netlist.flatten(netlist.parse_netlist(BUTTERWORTH), None, outputs=["vout"])
    ng = golden("ngspice_noise_butterworth.txt")
    f = acdec(20, 0.01, 10)
    assert np.allclose(f, ng[:, 0], rtol=1e-6)
    psd, st = orc.noise(fl.fc, f, opts=orc.default_options(temp=23.0, gmin=0.0))
    assert st.max() == 0
    mine = np.sqrt(np.abs(psd[0, :, 0]))
    k, T = 1.380649e-23, 23 + 273.15
    s = 2j * np.pi * f
    par = lambda a, b: a * b / (a + b)
    Z = par(par(s * 1.5, 1.0 / (s * (4.0 / 3.0))) + s * 0.5, 1.0)
    apsd = np.sqrt(np.abs(4 * k * T / 1.0 * Z * Z))
    assert isapprox(apsd, ng[:, 1], 1e-6)      # the reference's own cross-check of the table
    assert isapprox(mine, ng[:, 1], 1e-6)
    assert isapprox(mine, apsd, 1e-12)
    assert np.abs(mine / apsd - 1).max() < 1e-10


def test_bsimcmg_inverter_noise_vs_ngspice(host_bsimcmg):   # test/ac.jl:161-237
    fl = netlist.flatten(netlist.parse_netlist(BSIMCMG_INVERTER), None, outputs=["q"], host=True)
    ng = golden("ngspice_noise_bsimcmg_inverter.txt")
    f = acdec(5, 1e3, 1e15)
    assert len(f) == 61 and np.allclose(f, ng[:, 0], rtol=1e-8)
    psd, st = orc.noise(fl.fc, ng[:, 0])
    assert st.max() == 0
    mine = np.sqrt(np.abs(psd[0, :, 0]))
    assert isapprox(mine, ng[:, 1], 1e-6)                    # the reference's assertion
    assert np.abs(mine / ng[:, 1] - 1).max() < 1e-7          # element-wise, over 17 decades of frequency


def test_bsimcmg_inverter_ac_gain_matches_finite_difference(host_bsimcmg):
    """AC gain at low frequency == slope of the DC transfer curve (ties cb_ac's linearisation to the DC solver)."""
    fl = netlist.flatten(netlist.parse_netlist(circuits.BSIMCMG_INVERTER_VIN_DECK),
                         {"vin": np.array([0.5 - 1e-5, 0.5, 0.5 + 1e-5])}, outputs=["q"], host=True)
    x, _, st, _ = orc.dc(fl.fc, fl.params)
    assert st.max() == 0
    slope = (x[0, 2] - x[0, 0]) / 2e-5
    y, st = orc.ac(fl.fc, [1.0], fl.params)
    assert abs(y[0, 0, 1].real / slope - 1) < 1e-6 and abs(y[0, 0, 1].imag) < 1e-6 * abs(slope)
