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
    assert lib.cb_version() == 3


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


# ---- the header compiles as plain C, and every binding's struct mirror has the C layout -------------------------------
def _build_abi_smoke():
    import subprocess
    out = os.path.join(ROOT, "tests", "_build", "abi_smoke")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-o", out,
                    os.path.join(ROOT, "tests", "abi_smoke.c"), "-ldl", "-lm"], check=True)
    return out


def _c_layout():
    import json
    import subprocess
    return json.loads(subprocess.run([_build_abi_smoke(), "layout"], check=True, capture_output=True, text=True).stdout)


def test_header_compiles_as_c99_and_ctypes_mirrors_match_offsetof():
    from cedarsim.jl_b200 import flat
    lay = _c_layout()
    for name in ("cb_pref", "cb_device", "cb_wave", "cb_va_model", "cb_va_inst", "cb_flat_circuit", "cb_options", "cb_stats"):
        st = getattr(flat, name)
        assert C.sizeof(st) == lay[name][1], name
        seen = 0
        for fname, ftype in st._fields_:
            key = f"{name}.{fname}"
            if key not in lay:     # padding fields are not listed by the harness
                assert fname.startswith("_") or fname.endswith("_"), key
                continue
            assert [getattr(st, fname).offset, C.sizeof(ftype)] == lay[key], key
            seen += 1
        assert seen == sum(1 for k in lay if k.startswith(name + ".")), f"{name}: the C harness lists fields the ctypes mirror lacks"
    lib = engine.load()
    assert lib.cb_options_size() == lay["cb_options"][1] and lib.cb_stats_size() == lay["cb_stats"][1]


_JL_TYPES = {"Cdouble": (8, 8), "Float64": (8, 8), "Int32": (4, 4), "UInt32": (4, 4), "Int64": (8, 8), "Cint": (4, 4)}


def _julia_struct_layout(text, name, known):
    body = re.search(r"(?:mutable )?struct " + name + r"\n(.*?)\nend", text, flags=re.S).group(1)
    off, align, fields = 0, 1, {}
    for line in body.splitlines():
        m = re.match(r"\s*(\w+)::(\w+)\s*$", line)
        if not m:
            continue
        size, al = known[m.group(2)] if m.group(2) in known else _JL_TYPES[m.group(2)]
        off = (off + al - 1) // al * al
        fields[m.group(1)] = [off, size]
        off += size
        align = max(align, al)
    return fields, (off + align - 1) // align * align, align


def test_julia_extension_struct_mirrors_match_the_c_layout():
    lay = _c_layout()
    text = open(os.path.join(ROOT, "ext", "CedarSimB200Ext.jl")).read()
    known = {}
    for name in ("cb_pref", "cb_options", "cb_stats"):
        fields, size, align = _julia_struct_layout(text, name, known)
        known[name] = (size, align)
        assert size == lay[name][1], name
        c_fields = {k.split(".", 1)[1]: v for k, v in lay.items() if k.startswith(name + ".")}
        for f, v in c_fields.items():
            assert fields[f] == v, f"{name}.{f}"
        extra = set(fields) - set(c_fields)
        assert all(f.startswith("_") or f.endswith("_") for f in extra), extra
    # every entry point the extension binds exists in the header
    bound = set(re.findall(r"\(:(cb_\w+), lib\)", text))
    assert bound and bound <= set(header_symbols())


def test_stale_or_uninitialised_options_are_refused():
    lib = engine.load()
    from cedarsim.jl_b200 import flat
    o = flat.cb_options()
    assert lib.cb_options_init(C.byref(o), C.c_size_t(C.sizeof(o) - 8)) == -1       # a mirror that lacks the last field
    assert b"layout mismatch" in lib.cb_last_error()
    assert o.struct_size == 0                                                         # nothing was written
    assert lib.cb_options_init(C.byref(o), C.c_size_t(C.sizeof(o))) == 0
    assert o.struct_size == C.sizeof(o) and o.abi_version == 3


def test_flatckt_file_round_trip(tmp_path, host_bsimcmg):
    """save_flatckt -> cb_circuit_load -> cb_circuit_compile gives the same symbolic analysis as the struct path."""
    from cedarsim.jl_b200.flat import save_flatckt
    lib = engine.load()
    for fc, ms in ((circuits.two_resistor(), ()), circuits.inverter(host=False)):
        ref = engine.Circuit(fc, ms).lu_info()
        path = str(tmp_path / "c.flatckt")
        save_flatckt(fc, ms, path)
        h = C.c_void_p()
        assert lib.cb_circuit_load(path.encode(), C.byref(h)) == 0, lib.cb_last_error()
        assert lib.cb_circuit_compile(h, engine.CUBIN_CACHE.encode(), None) == 0, lib.cb_last_error()
        a, l, f = C.c_int32(), C.c_int32(), C.c_int64()
        assert lib.cb_circuit_lu_info(h, C.byref(a), C.byref(l), C.byref(f)) == 0
        assert {"nnz_a": a.value, "nnz_lu": l.value, "lu_flops": f.value} == ref
        lib.cb_circuit_destroy(h)
    bad = tmp_path / "bad.flatckt"
    bad.write_bytes(open(path, "rb").read()[:100])
    assert lib.cb_circuit_load(str(bad).encode(), C.byref(h)) == -1


def test_flatten_cli_writes_what_the_julia_extension_reads(tmp_path):
    import json
    from cedarsim.jl_b200 import flatten
    deck = tmp_path / "div.cir"
    deck.write_text("* divider\n.param r1=1k r2=1k\nV1 vcc 0 1\nR1 vcc out 'r1'\nR2 out 0 'r2'\n")
    sweep = tmp_path / "pts.csv"
    sweep.write_text("r1,r2\n100,100\n200,100\n100,300\n")
    flatten.main([str(deck), "--sweep", str(sweep), "--outputs", "out,v1.i", "--out", str(tmp_path / "run")])
    meta = json.load(open(tmp_path / "run.json"))
    assert meta["B"] == 3 and meta["P"] == 2 and meta["outputs"] == ["out", "v1.i"]
    P = np.fromfile(tmp_path / "run.params.f64").reshape(meta["P"], meta["B"])
    assert sorted(P[meta["param_names"].index("r1.r")]) == [100.0, 100.0, 200.0]   # columns are per-device values derived from the swept names
    h = C.c_void_p()
    lib = engine.load()
    assert lib.cb_circuit_load(str(tmp_path / "run.flatckt").encode(), C.byref(h)) == 0
    assert lib.cb_circuit_compile(h, None, None) == 0
    lib.cb_circuit_destroy(h)


@pytest.mark.gpu
def test_c_harness_solves_the_divider_sweep_through_the_abi():
    import json
    import subprocess
    r = subprocess.run([_build_abi_smoke(), "run", engine.LIB_PATH], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["points"] == 400 and out["max_abs_err"] < 1e-12


def test_c_harness_flattens_a_deck_with_the_native_front_end():
    """Pure C, no GPU: deck text -> cb_netlist_flatten -> cb_netlist_circuit -> cb_circuit_compile (reference deck of
    test/sweep.jl:342-371: a subcircuit parameter swept through the instance)."""
    import json
    import subprocess
    r = subprocess.run([_build_abi_smoke(), "netlist", engine.LIB_PATH, "flatten"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out == {"unknowns": 2, "params": ["x1.r1.r", "v1.dc"], "first_row": [5.0, 5.0]}


@pytest.mark.gpu
def test_c_harness_solves_a_deck_sweep_from_text():
    import json
    import subprocess
    r = subprocess.run([_build_abi_smoke(), "netlist", engine.LIB_PATH], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["points"] == 256 and out["max_abs_err"] < 1e-12
