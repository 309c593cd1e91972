"""A plan executed as several concurrent lanes (cb_plan_create_lanes) must return exactly what one lane returns:
every sweep point is solved on its own, the lanes only change which stream it runs on."""
import ctypes as C

import numpy as np
import pytest

from cedarsim.jl_b200 import circuits, engine
from cedarsim.jl_b200.flat import nodeset_vector

pytestmark = pytest.mark.gpu


def test_lanes_equal_single_lane_inverter(host_bsimcmg):
    fc, ms = circuits.inverter(tscale=0.01)
    B = 300                                    # lanes of 128 + 128 + 44 points
    P = np.zeros((3, B))
    P[fc.param_names.index("vvdd.dc")] = np.linspace(0.6, 0.8, B)
    P[fc.param_names.index("xneg.nfin")] = 3.0
    P[fc.param_names.index("xneg.l")] = np.linspace(21e-9, 30e-9, B)
    x0 = np.zeros((fc.n_unknowns, B))
    x0[fc.unknown("vdd")] = P[0]
    ts = np.linspace(0, 2e-9, 41)
    c = engine.Circuit(fc, ms)
    res = {}
    for lanes in (1, 3):
        p = c.plan(B, lanes=lanes)
        assert p.lanes == lanes
        p.set_params(P)
        p.set_x0(x0)
        xd, xf, sd, _ = p.dc()
        y, st, stats = p.tran(0.0, 2e-9, ts, engine.default_options(reltol=1e-4))
        # device-resident variant: parameters written into the plan's device buffer, results read back from HBM
        import torch
        dp = p.device_params_ptr()
        class A:   # noqa: E306
            def __init__(self, ptr, shape, ts_="<f8"):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": ts_, "data": (int(ptr), False), "version": 2}
        torch.as_tensor(A(dp, P.shape), device="cuda").copy_(torch.from_numpy(P[:, ::-1].copy()))
        dy, ds, _ = p.tran_device(0.0, 2e-9, ts, engine.default_options(reltol=1e-4))
        yd = torch.as_tensor(A(dy, (len(fc.outputs), len(ts), B)), device="cuda").cpu().numpy()
        sdv = torch.as_tensor(A(ds, (B,), "<i4"), device="cuda").cpu().numpy()
        res[lanes] = (xd, xf, sd, y, st, yd, sdv, stats["newton_iters"])
        p.close()
    for a, b in zip(res[1], res[3]):
        assert np.array_equal(a, b)
    assert res[1][4].max() == 0 and np.abs(res[1][5][:, :, ::-1] - res[1][3]).max() < 1e-3   # reversed params, same physics


def test_lanes_small_signal(host_bsimcmg):
    from cedarsim.jl_b200 import netlist
    from cedarsim.jl_b200.sweeps import acdec
    B = 200
    fl = netlist.flatten(netlist.parse_netlist(circuits.BSIMCMG_INVERTER_VIN_DECK),
                         {"vin": np.linspace(0.1, 0.9, B), "mneg.nfin": np.full(B, 2.0)}, outputs=["q"], host=True)
    c = engine.Circuit(fl.fc, fl.models)
    f = acdec(1, 1e3, 1e12)
    out = {}
    for lanes in (1, 2):
        p = c.plan(B, lanes=lanes)
        p.set_params(fl.params)
        out[lanes] = (p.ac(f)[0], p.noise(f)[0])
        p.close()
    assert np.array_equal(out[1][0], out[2][0]) and np.array_equal(out[1][1], out[2][1])


def _multi_case(devices):
    fc, ms = circuits.inverter(tscale=0.01)
    B = 700                                    # blocks of 384 + 316 points over two device entries
    P = np.zeros((3, B))
    P[fc.param_names.index("vvdd.dc")] = np.linspace(0.6, 0.8, B)
    P[fc.param_names.index("xneg.nfin")] = 3.0
    P[fc.param_names.index("xneg.l")] = np.linspace(21e-9, 30e-9, B)
    ts = np.linspace(0, 2e-9, 41)
    c = engine.Circuit(fc, ms)
    out = []
    for kw in (dict(lanes=1), dict(devices=devices, lanes=2)):
        p = c.plan(B, **kw)
        p.set_params(P)
        xd, xf, sd, _ = p.dc()
        y, st, _ = p.tran(0.0, 2e-9, ts, engine.default_options(reltol=1e-4))
        out.append((xd, xf, sd, y, st, p.n_devices, p.lanes))
        if "devices" in kw and len(set(devices)) > 1:
            with pytest.raises(engine.EngineError):     # device-resident entry points need a single-device plan
                p.tran_device(0.0, 2e-9, ts)
        p.close()
    for a, b in zip(out[0][:5], out[1][:5]):
        assert np.array_equal(a, b)
    return out[1][5], out[1][6]


def test_multi_device_plan_blocks_on_one_gpu(host_bsimcmg):
    """cb_plan_create_multi with the same GPU named twice: the partition into per-device blocks and lanes is exercised on
    a single-GPU box; results equal the single-lane plan bit for bit."""
    ndev, lanes = _multi_case([0, 0])
    assert ndev == 1 and lanes == 4


def test_multi_device_plan_two_gpus(host_bsimcmg):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    ndev, lanes = _multi_case([0, 1])
    assert ndev == 2 and lanes == 4
