"""The C-ABI library loads and exports every symbol include/cedarb200.h declares; circuits compile
(symbolic analysis + NVRTC) without a GPU; solving without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cedarsim.jl_b200 import circuits, engine
from cedarsim.jl_b200.flat import FlatCircuit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "cedarb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = engine.load()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in cedarb200.h but not exported"
    assert set(engine.SYMBOLS) <= set(syms)
    assert lib.cb_version() == 2


def test_options_defaults_and_struct_layout():
    o = engine.default_options()
    assert o.temp.value == 27.0 and o.temp.col == -1 and o.gmin.value == 1e-12
    assert o.reltol == 1e-3 and o.dc_abstol == 1e-10 and o.max_newton_dc == 200 and o.method == 1
    o2 = engine.default_options(reltol=1e-5, temp=50.0)
    assert o2.reltol == 1e-5 and o2.temp.value == 50.0
    with pytest.raises(KeyError):
        engine.default_options(bogus=1)


def test_symbolic_compile_without_gpu_and_no_cpu_fallback():
    fc = circuits.two_resistor()
    c = engine.Circuit(fc)
    info = c.lu_info()
    assert info["nnz_a"] == 6 and info["nnz_lu"] >= 6
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(engine.EngineError) as e:
            c.plan(4)
        assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_structurally_singular_circuit_is_rejected():
    fc = FlatCircuit()
    fc.vsource("V1", "a", "0", 1.0)
    fc.vsource("V2", "a", "0", 2.0)   # loop of voltage sources
    fc.resistor("R", "a", "0", 1.0)
    fc.set_outputs(["a"])
    with pytest.raises(engine.EngineError) as e:
        engine.Circuit(fc)
    assert e.value.code == -5


def test_invalid_arguments():
    fc = FlatCircuit()
    fc.resistor("R", "a", "0", 1.0)
    fc.set_outputs(["a"])
    pk = fc.pack()
    pk.struct.n_unknowns = 0
    h = C.c_void_p()
    assert engine.load().cb_circuit_create(pk.ref(), C.byref(h)) == -1
    assert b"unknown" in engine.load().cb_last_error()


def test_bsimcmg_nvrtc_compile_without_gpu(host_bsimcmg):
    fc, ms = circuits.dff()
    c = engine.Circuit(fc, ms)
    info = c.lu_info()
    assert fc.n_unknowns == 85 and info["nnz_lu"] >= info["nnz_a"] > 400
