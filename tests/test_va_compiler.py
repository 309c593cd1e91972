"""Verilog-A front end + AD code generator: generated Jacobians against finite differences, staging
(setup vs eval), card specialisation against the generic build."""
import os

import numpy as np
import pytest

from cedarsim.jl_b200 import models
from cedarsim.jl_b200.va.build import build_host
from cedarsim.jl_b200.va.compiler import VACompileError, compile_va_file, compile_va_text
from cedarsim.jl_b200.va.parser import parse
from cedarsim.jl_b200.va.preproc import preprocess_text

HERE = os.path.dirname(os.path.abspath(__file__))


def fd_check(hm, cache, v, eps=1e-6):
    I, Q, G, Cm = hm.run_eval(cache, v)
    worst = 0.0
    for l in range(len(v)):
        vp, vm = v.copy(), v.copy()
        vp[l] += eps
        vm[l] -= eps
        Ip, Qp, _, _ = hm.run_eval(cache, vp)
        Im, Qm, _, _ = hm.run_eval(cache, vm)
        worst = max(worst, np.abs((Ip - Im) / (2 * eps) - G[:, l]).max() / (np.abs(G).max() + 1e-300),
                    np.abs((Qp - Qm) / (2 * eps) - Cm[:, l]).max() / (np.abs(Cm).max() + 1e-300))
    return worst


def test_preprocessor_macros_and_ifdef():
    text = preprocess_text("`define A 2\n`define F(x,y) ((x)*(y)+`A)\n`ifdef A\nreal q = `F(3, 4);\n`else\nbad\n`endif\n")
    assert "((3)*(4)+2)" in text and "bad" not in text


def test_parser_number_scale_factors_and_precedence():
    mod = parse("module m(a,b); electrical a,b; parameter real R = 1.5k from (0:inf); real x;"
                " analog begin x = -2**2 + 3*4; I(a,b) <+ V(a,b)/R; end endmodule")[0]
    assert mod.params[0].default == ("num", 1500.0, False)
    assert mod.ports == ["a", "b"]


def test_nlvcr_jacobian_ddx_vaconvert():
    cm = compile_va_file(os.path.join(HERE, "va", "nlvcr.va"))
    assert cm.terminals == ["p", "n", "cp", "cn"]
    hm = build_host(cm)
    cache = hm.run_setup({"K": 0.7})
    v = np.array([0.3, -0.2, 0.15, 0.05])
    assert fd_check(hm, cache, v) < 1e-8
    I, Q, G, Cm = hm.run_eval(cache, v)
    assert abs(I[0] + I[1]) < 1e-18 and abs(Q[0] + Q[1]) < 1e-30        # KCL: what leaves p enters n
    # MODE*1.5 = 1.5 -> integer 2 (ties away from zero, src/va_env.jl:107) selects the nonlinear branch
    cache1 = hm.run_setup({"K": 0.7, "MODE": 0})
    I1, _, G1, _ = hm.run_eval(cache1, v)
    assert abs(I1[0] - 0.5 / 1000.0) < 1e-15 and G1[0, 2] == 0.0
    # dependent default R1 = 2*R0
    c2 = hm.run_setup({"K": 0.0, "R0": 500.0})
    I2, _, _, _ = hm.run_eval(c2, v)
    assert abs(I2[0] - 0.5 * (1 / 500.0 + 1 / 1000.0)) < 1e-15


def test_contribution_sign_reversed_pair():
    # reference test/varegress.jl: I(n,p) <+ x is the negative of I(p,n) <+ x
    src = """`include "disciplines.vams"
module r2(p, n); inout p, n; electrical p, n; parameter real R = 2.0;
analog begin I(n, p) <+ V(n, p) / R; end endmodule"""
    hm = build_host(compile_va_text(src))
    I, Q, G, _ = hm.run_eval(hm.run_setup({}), np.array([1.0, 0.0]))
    assert I[0] == 0.5 and I[1] == -0.5 and G[0, 0] == 0.5 and G[0, 1] == -0.5


def test_loops_and_case():
    src = """`include "disciplines.vams"
module lp(p, n); inout p, n; electrical p, n; parameter integer N = 3; parameter integer SEL = 1;
real acc, g; integer k;
analog begin
  acc = 0; k = 0;
  while (k < N) begin acc = acc + V(p,n) * V(p,n); k = k + 1; end
  case (SEL) 0: g = 1.0; 1, 2: g = 2.0; default: g = 3.0; endcase
  I(p,n) <+ g * acc;
end endmodule"""
    hm = build_host(compile_va_text(src))
    I, _, G, _ = hm.run_eval(hm.run_setup({}), np.array([0.5, 0.0]))
    assert abs(I[0] - 2.0 * 3 * 0.25) < 1e-15 and abs(G[0, 0] - 2.0 * 3 * 2 * 0.5) < 1e-15
    I, _, _, _ = hm.run_eval(hm.run_setup({"SEL": 7, "N": 1}), np.array([0.5, 0.0]))
    assert abs(I[0] - 3.0 * 0.25) < 1e-15


def test_unsupported_constructs_fail_loudly():
    # a probe of a discipline access the engine has no unknown for
    src = """`include "disciplines.vams"
module vs(p, n); inout p, n; electrical p, n; analog begin I(p,n) <+ Temp(p,n); end endmodule"""
    with pytest.raises(VACompileError):
        compile_va_text(src)


def test_bsimcmg_jacobian_fd_and_specialisation(host_bsimcmg):
    cards = models.asap7_cards()
    rng = np.random.default_rng(7)
    for card in ("nmos_lvt", "pmos_lvt"):
        cm = models.bsimcmg107_card(card)
        assert cm.terminals == ["d", "g", "s", "e", "di", "si"] and cm.ncache < 400 and sum(cm.exec_ops) > 5000
        hm = build_host(cm)
        cache = hm.run_setup({"L": 21e-9, "NFIN": 3})
        assert np.isfinite(cache).all()
        for _ in range(5):
            v = rng.uniform(-0.1, 0.8, 6)
            v[4] = v[0] + rng.normal() * 1e-3
            v[5] = v[2] + rng.normal() * 1e-3
            assert fd_check(hm, cache, v) < 2e-6
        # zero bias: finite Jacobian (pow(0, 0) derivative guard)
        I, Q, G, Cm = hm.run_eval(cache, np.zeros(6))
        assert np.isfinite(G).all() and np.isfinite(Cm).all()
    if os.path.exists(models.BSIMCMG_VA):
        generic = build_host(models.bsimcmg107())
        spec = build_host(models.bsimcmg107_card("nmos_lvt"))
        c0 = generic.run_setup(dict(cards["nmos_lvt"].params, L=21e-9, NFIN=3))
        c1 = spec.run_setup({"L": 21e-9, "NFIN": 3})
        for _ in range(20):
            v = rng.uniform(-0.2, 0.9, 6)
            for a, b in zip(spec.run_eval(c1, v), generic.run_eval(c0, v)):
                assert np.abs(a - b).max() <= 1e-12 * (np.abs(b).max() + 1e-300)


def test_bsimcmg_physical_sanity(host_bsimcmg):
    hm = build_host(models.bsimcmg107_card("nmos_lvt"))
    cache = hm.run_setup({"L": 21e-9, "NFIN": 3})
    ids = []
    for vg in np.linspace(0, 0.7, 8):
        I, _, _, _ = hm.run_eval(cache, np.array([0.7, vg, 0.0, 0.0, 0.7, 0.0]))
        ids.append(I[4])   # current leaving di through the channel
    ids = np.array(ids)
    assert np.all(np.diff(ids) > 0) and ids[0] < 1e-8 and 5e-5 < ids[-1] < 5e-4   # monotone, off ~nA, on ~100 uA


VBRANCH_DC = """* voltage branches and current probes
.hdl "vbranch.va"
x1 a 0 va_vsrc vdc=2 rs=10
r1 a 0 30
x2 a2 0 va_vsrc_rev vdc=2 rs=10
r2 a2 0 30
v3 c 0 1
x3 c 0 o 0 va_ccvs rsense=100 k=50
r3 o 0 1k
"""
VBRANCH_RL = """* RL with a Verilog-A inductor
.hdl "vbranch.va"
v1 in 0 PWL(0 0 1n 1)
r1 in x 100
x1 x 0 va_ind l=1u
"""


def test_voltage_branches_and_current_probes():
    """`V(a,b) <+` branches and `I(a,b)` probes get a branch-current unknown each (src/vasim.jl:786-808: only used
    branches do), row  V(a,b) - sum = 0  resp.  I_br - sum = 0, +I_br leaving a (src/simulate_ir.jl:112-120);
    a reversed pair contributes with the opposite sign (src/vasim.jl:165,177)."""
    from cedarsim.jl_b200 import netlist
    from oracle import orc
    inc = os.path.join(HERE, "va")
    cm = compile_va_file(os.path.join(inc, "vbranch.va"), module="va_ccvs")
    assert cm.terminals == ["p", "n", "op", "on", "I(op,on)", "I(p,n)"] and cm.branch_terms == [4, 5] and cm.nports == 4
    fl = netlist.flatten(netlist.parse_netlist(VBRANCH_DC, include_dirs=[inc]), host=True)
    fc = fl.fc
    assert fc.branch_names == ["v3.i", "x1.i(p,n)", "x2.i(n,p)", "x3.i(op,on)", "x3.i(p,n)"]      # currents after the node voltages
    x, xf, st, _ = orc.dc(fc, None)
    assert st.max() == 0
    val = lambda n: xf[fc.unknown(n), 0]
    # Thevenin source 2 V / 10 Ohm into 30 Ohm, written on (p,n) and on the reversed pair
    assert abs(val("a") - 1.5) < 1e-12 and abs(val("x1.i(p,n)") + 0.05) < 1e-14
    assert abs(val("a2") - 1.5) < 1e-12 and abs(val("x2.i(n,p)") - 0.05) < 1e-14
    # CCVS: the probed current of an I() <+ branch drives a V() <+ branch
    assert abs(val("x3.i(p,n)") - 0.01) < 1e-15 and abs(val("o") - 0.5) < 1e-13 and abs(val("v3.i") + 0.01) < 1e-15
    # inductor V <+ L ddt(I): ramp-and-hold response of the RL circuit against its closed form
    fl = netlist.flatten(netlist.parse_netlist(VBRANCH_RL, include_dirs=[inc]), {"x1.l": np.array([1e-6, 2e-6])}, host=True)
    fc = fl.fc
    ts = np.linspace(0, 5e-8, 51)
    y, st, _ = orc.tran(fc, 0.0, 5e-8, ts, params=fl.params, opts=orc.default_options(reltol=1e-6, vabstol=1e-9, iabstol=1e-12))
    assert st.max() == 0
    cur = y[fc.unknown("x1.i(p,n)")]
    for k, L in enumerate((1e-6, 2e-6)):
        tau, T = L / 100.0, 1e-9
        ramp = lambda t: (1 / 100.0) * (t - tau * (1 - np.exp(-t / tau))) / T
        exact = np.where(ts <= T, ramp(ts), ramp(ts) - ramp(np.maximum(ts - T, 0.0)))
        assert np.abs(cur[:, k] - exact).max() < 5e-7
    # a branch that receives both kinds of contribution is the reference's switch branch (src/vasim.jl:149-154): lowered to a
    # run-time mode + the branch's own current unknown (_lower_switch_branches; solved in tests/test_oracle_golden.py) ...
    cm = compile_va_text("`include \"disciplines.vams\"\nmodule sw(p,n); inout p,n; electrical p,n;\nanalog begin\n"
                         "if (V(p,n) > 0) V(p,n) <+ 0; else I(p,n) <+ 0;\nend\nendmodule\n")
    assert cm.terminals == ["p", "n", "I(p,n)"] and cm.branch_terms == [2]
    # ... except with ddt() in it, which is refused, not mis-compiled
    with pytest.raises(VACompileError, match="switch branch"):
        compile_va_text("`include \"disciplines.vams\"\nmodule sw(p,n); inout p,n; electrical p,n;\nanalog begin\n"
                        "if (V(p,n) > 0) V(p,n) <+ 1e-6 * ddt(I(p,n)); else I(p,n) <+ 0;\nend\nendmodule\n")


def test_branch_current_observable_of_a_va_device():
    """test/varegress.jl:20-38, :49-67: sys.R.var"I(p, n)" >= 0 along the RC charge, for the resistor written on (p,n)
    and for the one written on the reversed pair.  Asking for `<inst>.I(p, n)` as an output makes that branch current
    an unknown of the device (src/vasim.jl:786-808)."""
    from cedarsim.jl_b200 import netlist
    from oracle import orc
    inc = os.path.join(HERE, "va")
    ts = np.linspace(0, 1e-5, 101)
    for mod in ("VAR", "VAR_rev"):
        deck = f'* varegress\n.hdl "varegress.va"\nv1 vcc 0 1\nxr vcc out {mod} r=1000\nc1 out 0 1n\n'
        fl = netlist.flatten(netlist.parse_netlist(deck, include_dirs=[inc]), host=True, outputs=["out", "xr.I(p, n)"])
        assert fl.fc.branch_names == ["v1.i", "xr.i(p,n)"]
        y, st, _ = orc.tran(fl.fc, 0.0, 1e-5, ts, opts=orc.default_options(skip_dc=1, reltol=1e-6))   # u0 = 0: uncharged start
        assert st.max() == 0
        out, cur = y[0, :, 0], y[1, :, 0]
        assert np.all(cur >= 0.0)
        assert np.abs(cur[1:] - (1.0 - out[1:]) / 1000.0).max() < 1e-12          # the branch equation itself
        assert np.abs(out - (1.0 - np.exp(-ts / 1e-6))).max() < 1e-4             # RC charge
