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
