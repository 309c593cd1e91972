"""Sweep iterators: restatement of the reference's own unit tests (test/sweep.jl:71-242, 373-410)."""
import numpy as np
import pytest

from cedarsim.jl_b200.sweeps import (ProductSweep, SerialSweep, Sweep, TandemSweep, find_param_ranges, split_axes,
                                     sweepify, sweepvars)


def frange(a, step, b):
    return list(np.arange(a, b + step / 2, step))


def test_sweep_basic():   # test/sweep.jl:71-92
    s = Sweep("R1", frange(0.1, 0.1, 1.0))
    assert len(s) == 10 and s.size() == (10,) and s.size(1) == 10 and s.size(2) == 1
    assert s.first() == (("R1", 0.1),)
    assert Sweep(R1=frange(0.1, 0.1, 1.0)) == s
    assert Sweep(**{"a.b": [1.0, 2.0]}).first() == (("a.b", 1.0),)
    s = Sweep(a=10.0)
    assert len(s) == 1 and s.size() == (1,) and s.first() == (("a", 10.0),)
    with pytest.raises(ValueError):
        Sweep(a=[1], b=[2])


def test_product_sweep():   # test/sweep.jl:94-116
    s = ProductSweep(R1=[1.0, 2.0], R2=[1.0, 2.0], R3=[1.0, 2.0])
    assert s.size() == (2, 2, 2) and s.size(3) == 2 and s.size(4) == 1
    assert s.first() == (("R1", 1.0), ("R2", 1.0), ("R3", 1.0))
    assert sweepvars(s) == {"R1", "R2", "R3"}
    pts = list(s)
    assert len(pts) == 8
    # first axis varies fastest (Base.Iterators.product)
    assert pts[1] == (("R1", 2.0), ("R2", 1.0), ("R3", 1.0))
    assert pts[2] == (("R1", 1.0), ("R2", 2.0), ("R3", 1.0))
    e = ProductSweep()
    assert e.size() == () and sweepvars(e) == set() and e.first() == ()


def test_tandem_sweep():   # test/sweep.jl:118-131
    s = TandemSweep(R1=[1.0, 2.0], R2=[1.0, 2.0], R3=[1.0, 2.0])
    assert s.size() == (2,) and s.size(2) == 1
    assert list(s) == [(("R1", 1.0), ("R2", 1.0), ("R3", 1.0)), (("R1", 2.0), ("R2", 2.0), ("R3", 2.0))]
    with pytest.raises(ValueError):
        TandemSweep(a=[1, 2], b=[1, 2, 3])


def test_serial_sweep():   # test/sweep.jl:133-150
    s = SerialSweep(R1=[1.0, 2.0], R2=[1.0, 2.0])
    assert s.size() == (4,) and len(s) == 4
    assert list(s) == [(("R1", 1.0), ("R2", None)), (("R1", 2.0), ("R2", None)),
                       (("R1", None), ("R2", 1.0)), (("R1", None), ("R2", 2.0))]
    cols = s.columns()
    assert np.isnan(cols["R2"][:2]).all() and cols["R1"][1] == 2.0


def test_nested_sweeps():   # test/sweep.jl:152-202
    s = ProductSweep(TandemSweep(a=[1, 2], b=[3, 4]), c=[5, 6, 7])
    assert s.size() == (2, 3) and len(s) == 6
    assert list(s)[0] == (("a", 1), ("b", 3), ("c", 5))
    assert list(s)[1] == (("a", 2), ("b", 4), ("c", 5))
    s = SerialSweep(ProductSweep(a=[1, 2], b=[1, 2]), TandemSweep(c=[1, 2], d=[3, 4]))
    assert len(s) == 6 and sweepvars(s) == {"a", "b", "c", "d"}
    assert list(s)[-1] == (("a", None), ("b", None), ("c", 2), ("d", 4))


def test_split_axes():   # test/sweep.jl:210-242
    ps = ProductSweep(A=range(1, 11), B=range(1, 11), C=range(1, 6), D=range(1, 6))
    for v in (["A", "C"], ("A", "C")):
        outer, inner = split_axes(ps, v)
        assert sweepvars(outer) == {"B", "D"} and sweepvars(inner) == {"A", "C"}
        assert outer.size() == (10, 5) and inner.size() == (10, 5)
        ps2 = ProductSweep(outer, inner)
        assert sweepvars(ps2) == sweepvars(ps) and ps2.size() != ps.size()
    with pytest.raises(ValueError):
        split_axes(ps, ["E"])
    with pytest.raises(ValueError):
        split_axes(SerialSweep(A=range(10), B=range(10)), ["A"])


def test_find_param_ranges():   # test/sweep.jl:373-400
    params = ProductSweep(
        ProductSweep(SerialSweep(Sweep(a=range(1, 11)), Sweep(a=range(11, 21))), b=range(1, 6, 2)),
        ProductSweep(TandemSweep(c=range(1, 11), d=range(1, 11)), SerialSweep(c=range(11, 16), d=range(1, 6))))
    r = find_param_ranges(params)
    assert r["a"] == (1, 20, 20) and r["b"] == (1, 5, 3) and r["c"] == (1, 15, 15) and r["d"] == (1, 10, 15)


def test_sweepify():   # test/sweep.jl:402-410
    s1 = sweepify([dict(r1=range(1, 11), r2=range(1, 7)), dict(r3=range(1, 5), r4=range(1, 3))])
    s2 = SerialSweep(ProductSweep(r1=range(1, 11), r2=range(1, 7)), ProductSweep(r3=range(1, 5), r4=range(1, 3)))
    assert list(s1) == list(s2)
    s1 = sweepify([("r1", range(1, 11)), ("r2", range(1, 11))])
    s2 = SerialSweep(r1=range(1, 11), r2=range(1, 11))
    assert list(s1) == list(s2)


def test_columns_column_major():
    s = ProductSweep(R1=[100.0, 200.0, 300.0], R2=[1.0, 2.0])
    c = s.columns()
    assert c["R1"].tolist() == [100.0, 200.0, 300.0, 100.0, 200.0, 300.0]
    assert c["R2"].tolist() == [1.0, 1.0, 1.0, 2.0, 2.0, 2.0]


def test_branch_observables_and_csv(tmp_path):
    """`sol[cs.sys.x1.r1.I]` as the reference's own sweep test reads it (test/sweep.jl:363-369), reconstructed from the
    node voltages; results come from the CPU oracle here (the GPU run of the same accessors is in test_gpu_sweep_api)."""
    import numpy as np
    from cedarsim.jl_b200.sweeps import CircuitSweep, ProductSweep, Sweep, SweepSolution
    from oracle import orc
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
    x, _, st, stats = orc.dc(cs.flat.fc, cs.flat.params)
    sols = SweepSolution(cs, x, st, stats, None)
    for sol in sols:
        p = sol.params
        assert abs(p["v_in"] / p["x1.r_load"] - sol[cs.sys.x1.r1.I]) < 1e-7      # test/sweep.jl:369
        assert abs(sol[cs.sys.x1.r1.V] - p["v_in"]) < 1e-7 and abs(sol[cs.sys.v1.I] + sol[cs.sys.x1.r1.I]) < 1e-7
    assert sols.array(cs.sys.x1.r1.I).shape == (10, 10)
    # transient: capacitor current of an RC step == resistor current; CSV export of one point
    rc = "* rc\n.param r=1k\nV1 in 0 PULSE(0 1 0 1n 1n 1 2)\nR1 in out 'r'\nC1 out 0 1n\n.tran 10n 5u\n"
    cs = CircuitSweep(rc, Sweep(r=[500.0, 2000.0]))
    ts = np.linspace(0, 5e-6, 501)
    y, st, stats = orc.tran(cs.flat.fc, 0.0, 5e-6, ts, params=cs.flat.params, opts=orc.default_options(reltol=1e-6))
    sols = SweepSolution(cs, y, st, stats, ts)
    ir, ic = sols[1][cs.sys.r1.I], sols[1][cs.sys.c1.I]
    assert np.abs(ir[5:-5] - ic[5:-5]).max() < 2e-3 * np.abs(ir).max()
    f = sols.write_csv(str(tmp_path / "rc.csv"), index=1)
    rows = open(f).read().splitlines()
    assert rows[0] == "t,in,out" and len(rows) == 502
    assert abs(float(rows[-1].split(",")[2]) - y[cs.flat.fc.outputs.index(cs.flat.fc.unknown("out")), -1, 1]) < 1e-15
    # the Plotly extension's figure (ext/CedarSimPlotlyLightExt.jl:11-46): one line trace per top-level net, sorted keys
    spec = sols.plot_spec(index=1, title="rc")
    assert [tr["name"] for tr in spec["data"]] == ["in", "out"] and spec["layout"]["title"] == "rc"
    assert all(tr["type"] == "scatter" and tr["mode"] == "lines" and len(tr["x"]) == len(tr["y"]) == 501 for tr in spec["data"])
    page = open(sols.save_html(str(tmp_path / "rc.html"), index=1)).read()
    assert "Plotly.newPlot" in page and '"name": "out"' in page


def test_empty_sweep_is_refused():   # src/sweeps.jl:414-417: the circuit is compiled from the first sweep point
    from cedarsim.jl_b200.sweeps import CircuitSweep
    with pytest.raises(ValueError, match="empty sweep"):
        CircuitSweep("* r\nv1 a 0 1\nr1 a 0 1k\n", Sweep("r1.r", np.array([])))


def test_retry_ladder_host_logic_with_a_fake_engine():
    """sweeps._retry_failed (host retry policy) without a GPU: the unconverged points alone are handed
    to a new small plan with the ladder's options, points that converge there replace their entries, the others keep
    their status, converged points are never touched; the sub-batch gets the failed points' parameter columns / nodeset."""
    from cedarsim.jl_b200 import sweeps as S

    calls = []

    class FakePlan:
        def __init__(self, n):
            self.n = n
        def set_params(self, P):
            self.P = P
        def set_x0(self, x0):
            self.x0 = x0
        def close(self):
            calls[-1]["closed"] = True

    class FakeCompiled:
        def plan(self, n, devices=None):
            calls.append({"n": n, "devices": devices})
            p = FakePlan(n)
            calls[-1]["plan"] = p
            return p

    class FakeFlat:
        params = np.arange(2 * 8, dtype=float).reshape(2, 8)

    class FakeCS:
        flat, devices, x0, _compiled = FakeFlat(), [3], np.arange(5 * 8, dtype=float).reshape(5, 8), FakeCompiled()
        def _options(self, kw):
            return dict(kw)

    y = np.zeros((2, 8))
    status = np.array([0, 1, 0, 2, 0, 0, 4, 0], dtype=np.int32)
    stats = {}

    def solve(plan, opts):      # rung 1 rescues the first failed point only, rung 2 the last one
        rung = 1 if opts["source_steps"] == 40 else 2
        st = np.ones(plan.n, dtype=np.int32)
        st[0 if rung == 1 else -1] = 0
        return np.full((2, plan.n), 10.0 * rung) + plan.P[:1], st

    S._retry_failed(FakeCS(), dict(reltol=1e-3), True, solve, y, status, stats, point_axis=1)
    assert [c["n"] for c in calls] == [3, 2] and all(c["devices"] == [3] and c["closed"] for c in calls)
    assert np.array_equal(calls[0]["plan"].P, FakeFlat.params[:, [1, 3, 6]]) and np.array_equal(calls[0]["plan"].x0, FakeCS.x0[:, [1, 3, 6]])
    assert np.array_equal(calls[1]["plan"].P, FakeFlat.params[:, [3, 6]])
    assert status.tolist() == [0, 0, 0, 2, 0, 0, 0, 0]                       # point 3 stays InitialFailure
    assert y[0].tolist() == [0, 10.0 + 1, 0, 0, 0, 0, 20.0 + 6, 0]          # rescued entries only
    assert stats == {"retried_points": 3, "recovered_points": 2}
    # retry=False-like: an empty ladder does nothing
    calls.clear()
    S._retry_failed(FakeCS(), {}, (), solve, y, status, stats, point_axis=1)
    assert not calls
