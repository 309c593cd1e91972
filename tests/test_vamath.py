"""Accuracy of the branch-free exp / log / pow / reciprocal / sqrt of the CUDA prelude (csrc/va_prelude.h).

The functions are IEEE fma / mul / add sequences plus a low-precision hardware seed (MUFU.RCP64H / RSQ64H), so the
same source compiled for the host with fma() reproduces the device arithmetic; the seed is emulated by a
single-precision reciprocal (the Newton steps make the result independent of the seed's low bits).  Checked against
libm in ulps.  No GPU needed."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRELUDE = os.path.join(ROOT, "cedarsim.jl_b200", "csrc", "va_prelude.h")

SHIM = r"""
#include <math.h>
#include <stdint.h>
#include <string.h>
#define VA_FN static inline
#define VA_MATH_FN static inline
#define VA_MATH_FN2 static inline
#define __constant__ static const
#define __restrict__
typedef int bool_t;
#define bool int
static inline int __double2hiint(double d) { uint64_t b; memcpy(&b, &d, 8); return (int)(b >> 32); }
static inline int __double2loint(double d) { uint64_t b; memcpy(&b, &d, 8); return (int)(b & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo) { uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &b, 8); return d; }
// hardware seeds: ~20 good bits over the whole normal range, computed from the upper 32 bits of the operand (rcp);
// flush-to-zero on subnormal inputs and results
static inline double seed_rcp(double x) {
    if (isnan(x)) return x;
    if (fabs(x) < 2.2250738585072014e-308) return copysign(INFINITY, x);
    uint64_t b; memcpy(&b, &x, 8); b &= 0xffffffff00000000ull; double xh; memcpy(&xh, &b, 8);
    double r = (1.0 / xh) * (1.0 + 4e-7);
    return fabs(r) < 2.2250738585072014e-308 ? copysign(0.0, x) : r;
}
static inline double seed_rsqrt(double x) {
    if (isnan(x) || x < 0) return NAN;
    if (x < 2.2250738585072014e-308) return INFINITY;
    if (isinf(x)) return 0.0;
    return (1.0 / sqrt(x)) * (1.0 - 4e-7);
}
"""


def _build(tmp_path):
    text = open(PRELUDE).read()
    a = text.index("VA_FN double va_rcp_normal")
    b = text.index("#ifdef VA_EXACT_DIV")
    body = text[a:b]
    body = re.sub(r'asm\("rcp\.approx\.ftz\.f64 %0, %1;" : "=d"\((\w+)\) : "d"\((\w+)\)\);', r"\1 = seed_rcp(\2);", body)
    body = re.sub(r'asm\("rsqrt\.approx\.ftz\.f64 %0, %1;" : "=d"\((\w+)\) : "d"\((\w+)\)\);', r"\1 = seed_rsqrt(\2);", body)
    assert "asm(" not in body
    src = SHIM + body + """
void t_exp(const double* x, double* y, int n) { for (int i = 0; i < n; i++) y[i] = va_exp(x[i]); }
void t_log(const double* x, double* y, int n) { for (int i = 0; i < n; i++) y[i] = va_log(x[i]); }
void t_rcp(const double* x, double* y, int n) { for (int i = 0; i < n; i++) y[i] = va_rcp(x[i]); }
void t_sqrt(const double* x, double* y, int n) { for (int i = 0; i < n; i++) y[i] = va_sqrt(x[i]); }
void t_pow(const double* a, const double* b, double* y, int n) { for (int i = 0; i < n; i++) y[i] = va_pow(a[i], b[i]); }
"""
    cfile = os.path.join(tmp_path, "vamath.c")
    so = os.path.join(tmp_path, "vamath.so")
    open(cfile, "w").write(src)
    subprocess.run(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", so, cfile, "-lm"], check=True)
    return C.CDLL(so)


def _call(fn, *arrs):
    n = len(arrs[0])
    y = np.zeros(n)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    fn(*[dp(np.ascontiguousarray(a, dtype=np.float64)) for a in arrs], dp(y), C.c_int(n))
    return y


def _ulps(got, want):
    return np.abs(got - want) / np.maximum(np.abs(want) * 2.0 ** -52, 5e-324)


def test_va_math_accuracy(tmp_path):
    lib = _build(str(tmp_path))
    rng = np.random.default_rng(7)
    n = 400000
    x = np.concatenate([rng.uniform(-708, 709, n), rng.uniform(-80, 80, n), rng.uniform(-1, 1, n), rng.normal(0, 1e-5, n)])
    assert _ulps(_call(lib.t_exp, x), np.exp(x)).max() <= 1.5
    y = np.concatenate([np.exp(rng.uniform(-700, 700, n)), np.exp(rng.uniform(-40, 40, n)), 1 + rng.uniform(-0.5, 0.5, n),
                        1 + rng.normal(0, 1e-4, n), rng.uniform(1e-320, 1e-308, 1000)])
    assert _ulps(_call(lib.t_log, y), np.log(y)).max() <= 3.0
    z = np.concatenate([rng.uniform(-1e3, 1e3, n), np.exp(rng.uniform(-600, 600, n)) * rng.choice([-1, 1], n)])
    assert _ulps(_call(lib.t_rcp, z), 1.0 / z).max() <= 2.0
    assert _ulps(_call(lib.t_sqrt, np.abs(z)), np.sqrt(np.abs(z))).max() <= 2.0
    a, b = np.exp(rng.uniform(-10, 10, n)), rng.uniform(-10, 10, n)
    want = np.power(a, b)
    assert (np.abs(_call(lib.t_pow, a, b) - want) / want).max() <= 1e-13


def test_va_math_special_values(tmp_path):
    lib = _build(str(tmp_path))
    inf, nan = np.inf, np.nan
    e = _call(lib.t_exp, np.array([0.0, -800.0, -inf, 800.0, inf, nan, -708.0]))
    assert e[0] == 1.0 and e[1] == 0.0 and e[2] == 0.0 and e[3] == inf and e[4] == inf and np.isnan(e[5]) and e[6] > 0
    l = _call(lib.t_log, np.array([1.0, 0.0, -1.0, inf, nan, 5e-324]))
    assert l[0] == 0.0 and l[1] == -inf and np.isnan(l[2]) and l[3] == inf and np.isnan(l[4])
    assert abs(l[5] - np.log(5e-324)) < 1e-12
    p = _call(lib.t_pow, np.array([2.0, 0.0, 0.0, -2.0, -2.0, -2.0, 5.0, 0.0]), np.array([0.5, 0.0, 2.0, 3.0, 2.0, 0.5, 0.0, -1.0]))
    assert abs(p[0] - np.sqrt(2)) < 1e-15 and p[1] == 1.0 and p[2] == 0.0 and abs(p[3] + 8.0) < 1e-14 and abs(p[4] - 4.0) < 1e-14
    assert np.isnan(p[5]) and p[6] == 1.0 and p[7] == inf
    r = _call(lib.t_rcp, np.array([0.0, inf, 4.0]))
    assert r[0] == inf and r[1] == 0.0 and r[2] == 0.25
    s = _call(lib.t_sqrt, np.array([0.0, 4.0, -1.0, inf]))
    assert s[0] == 0.0 and s[1] == 2.0 and np.isnan(s[2]) and s[3] == inf
