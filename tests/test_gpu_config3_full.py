"""BASELINE config 3 in its fixed-step comparison mode (dt = 25 ps) over the FULL 0 .. 600 ns span
(24 000 steps per point, every clock and data edge of the deck), GPU engine through the C ABI against the CPU oracle at
the north-star tolerance: 1e-6 relative / 1e-9 V absolute at every shared output time.

Two option sets:
  * plain Newton with the default (tight) Newton tolerances -- the parity mode;
  * the options bench.py times: Newton tolerance tied to the LTE tolerance (1e-5 / 1e-7 V), Sundials-IDA-style rate
    test, chord (value-only) iterations `value_rounds = 2`, mixed rounds.  The oracle restates the chord schedule
    (oracle.cpp, Solver::newton), so both sides take the same iterations and the timed configuration gets the same tight
    bound as the parity mode instead of a tolerance-level one.
"""
import os

import numpy as np
import pytest

import bench
from cedarsim.jl_b200 import circuits

from helpers import run_tran_both, x0_from

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1
# CB_FILL_ORACLE_CACHE=1 python -m pytest tests/test_gpu_config3_full.py -m gpu: compute and store the oracle's answers of
# these tests (minutes of CPU) without a GPU; the assertions on the engine are skipped
ORACLE_ONLY = os.environ.get("CB_FILL_ORACLE_CACHE") == "1"


def _close(yg, yo, rtol=1e-6, atol=1e-9):
    err = np.abs(yg - yo)
    tol = rtol * np.abs(yo) + atol
    assert np.all(err <= tol), f"max err {err.max():.3e}, worst excess {(err - tol).max():.3e}"


def _run(B, **opts):
    fc, ms = circuits.dff(host=True)
    P = np.ascontiguousarray(circuits.dff_mc_params(fc, bench.TOTAL_POINTS)[:, :B])   # the first B of the bench's own draws
    ts = np.linspace(bench.T0, bench.T1, 601)
    return run_tran_both(fc, ms, P, bench.T0, bench.T1, ts, x0=x0_from(fc, bench.DFF_NODESET), nthreads=THREADS,
                         cache_oracle=True, oracle_only=ORACLE_ONLY, fixed_step=1, dt=bench.FIXED_DT, **opts)


def test_dff_mc_fixed_step_full_span_plain_newton(host_bsimcmg):
    g, (yo, so, sto) = _run(64)
    if ORACLE_ONLY:
        return
    yg, sg, stg = g
    assert sg.max() == 0 and so.max() == 0
    _close(yg, yo)
    assert stg["steps_accepted"] == sto["steps_accepted"] == 64 * 24000
    assert np.abs(yg[0, [150, 250, 450, 550, 600]] - np.array([0, 0, 0.7, 0.7, 0.7])[:, None]).max() < 1e-3   # test/gf180_dff.jl:29-33 pattern


@pytest.mark.parametrize("mixed", [1, 0])
def test_dff_mc_fixed_step_full_span_bench_options(host_bsimcmg, mixed):
    B = 256 if mixed else 64
    # (mixed_rounds is an engine-side schedule: the oracle's answer is the same file for both values)
    g, (yo, so, sto) = _run(B, engine_only=dict(mixed_rounds=mixed), **bench.OPTS, **bench.ENGINE_OPTS_BASE)
    if ORACLE_ONLY:
        return
    yg, sg, stg = g
    assert sg.max() == 0 and so.max() == 0
    _close(yg, yo)
    assert stg["steps_accepted"] == sto["steps_accepted"] == B * 24000
    # same iteration counts up to borderline convergence decisions (FMA vs non-FMA arithmetic)
    assert abs(stg["newton_iters"] - sto["newton_iters"]) <= 2e-3 * sto["newton_iters"]
    assert 0 < stg["full_iters"] < stg["newton_iters"]


def test_dff_mc_adaptive_full_span_bench_options(host_bsimcmg):
    """The configuration bench.py's headline `value` times (adaptive LTE control + the options above), 128 points, full
    span.  Step-size control is restated statement for statement by the oracle, so both sides take the same step
    sequence; asserted at the Newton tolerance of the run (1e-5 relative / 1e-7 V), the measured difference is printed."""
    fc, ms = circuits.dff(host=True)
    B = 128
    P = np.ascontiguousarray(circuits.dff_mc_params(fc, bench.TOTAL_POINTS)[:, :B])
    ts = np.linspace(bench.T0, bench.T1, 601)
    g, (yo, so, sto) = run_tran_both(fc, ms, P, bench.T0, bench.T1, ts, x0=x0_from(fc, bench.DFF_NODESET),
                                     nthreads=THREADS, cache_oracle=True, oracle_only=ORACLE_ONLY,
                                     engine_only=dict(mixed_rounds=1), **bench.OPTS, **bench.ENGINE_OPTS_BASE)
    if ORACLE_ONLY:
        return
    yg, sg, stg = g
    assert sg.max() == 0 and so.max() == 0
    err = np.abs(yg - yo)
    print(f"adaptive bench options: max |dV| = {err.max():.3e} V; steps {stg['steps_accepted']} vs {sto['steps_accepted']}, "
          f"rejected {stg['steps_rejected']} vs {sto['steps_rejected']}, iterations {stg['newton_iters']} vs {sto['newton_iters']}")
    _close(yg, yo, rtol=1e-5, atol=1e-7)
    assert abs(stg["steps_accepted"] - sto["steps_accepted"]) <= 2e-3 * sto["steps_accepted"]
