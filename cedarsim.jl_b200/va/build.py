"""Host-side build of generated model code: C prelude + gcc -> shared object (used by the CPU
oracle and by tests that check the generator against finite differences).  The CUDA prelude
for the same generated text lives in csrc/va_prelude.h and is compiled by NVRTC in the engine.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from typing import Optional

from .compiler import CompiledModel

_HERE = os.path.dirname(os.path.abspath(__file__))
GEN_DIR = os.path.join(os.path.dirname(_HERE), os.environ.get("CB_GEN_DIR", "_gen"))
HOST_CC = "/usr/bin/gcc"


def c_prelude(nterm: int, count_ops: bool = False) -> str:
    ops = ("double va_ops_[4];\n#define VA_OPS(a, m, d, s) (va_ops_[0] += (a), va_ops_[1] += (m), va_ops_[2] += (d), va_ops_[3] += (s))\n"
           if count_ops else "#define VA_OPS(a, m, d, s)\n")
    return ops + f"""
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#define NT {nterm}
#define VA_FN static inline
static inline double va_limexp(double x) {{ return x < 80.0 ? exp(x) : exp(80.0) * (1.0 + x - 80.0); }}
static inline double va_dlimexp(double x) {{ return x < 80.0 ? exp(x) : exp(80.0); }}
#define VA_SETUP_BEGIN(NAME) void NAME##_setup(const double* par_, const uint8_t* given_, double temp_c_, double gmin_, double* cache_) {{
#define VA_SETUP_END(NAME) }}
#define PAR(i) par_[i]
#define GIVEN(i) (given_[i] != 0)
#define TEMP_K (temp_c_ + 273.15)
#define GMIN_V gmin_
#define CACHE_ST(s, v) cache_[s] = (double)(v)
/* uniform slots (the same for every device of the model): on the host they sit behind the stream slots of the row */
#define CACHE_STU(k, v) cache_[VA_NSTREAM + (k)] = (double)(v)
#define CACHE_LDU(k) cache_[VA_NSTREAM + (k)]
#define VA_EVAL_BEGIN(NAME) void NAME##_eval(const double* cache_, const double* v_, double* I_, double* Q_, double* G_, double* C_) {{
#define VA_EVAL_END(NAME) }}
#define CACHE_LD(s) cache_[s]
#define CACHE_LDG(s) cache_[s]
#define VA_RCP(x) (1.0 / (x))
#define VA_SQRT(x) sqrt(x)
#define VA_DPOW(a, b, p) ((b) == 0.0 ? 0.0 : ((a) == 0.0 ? (b) * pow((a), (b) - 1.0) : (b) * (p) / (a)))
#define VA_CHUNK(k)
#define VT(k) v_[k]
#define OUT_I(k, v) I_[k] = (v)
#define OUT_Q(k, v) Q_[k] = (v)
#define OUT_J(idx, k, l, g, c) do {{ G_[(k) * NT + (l)] = (g); C_[(k) * NT + (l)] = (c); }} while (0)
/* value-only variant: currents and charges, no derivatives (the oracle's chord iterations, like the engine's k_evalv_*) */
#define VA_SETUPV_BEGIN(NAME) void NAME##_setupv(const double* par_, const uint8_t* given_, double temp_c_, double gmin_, double* cache_) {{
#define VA_SETUPV_END(NAME) }}
#define VA_EVALV_BEGIN(NAME) void NAME##_evalv(const double* cache_, const double* v_, double* I_, double* Q_) {{
#define VA_EVALV_END(NAME) }}
/* noise variant: power of every noise source at the bias point (and the flicker exponent) */
#define VA_SETUPN_BEGIN(NAME) void NAME##_setupn(const double* par_, const uint8_t* given_, double temp_c_, double gmin_, double* cache_) {{
#define VA_SETUPN_END(NAME) }}
#define VA_EVALN_BEGIN(NAME) void NAME##_noise(const double* cache_, const double* v_, double* N_, double* NE_) {{
#define VA_EVALN_END(NAME) }}
#define OUT_N(k, v) N_[k] = (v)
#define OUT_NE(k, v) NE_[k] = (v)
"""


class HostModel:
    """A generated model compiled for the host; exposes setup/eval through ctypes."""

    def __init__(self, cm: CompiledModel, so_path: str):
        self.cm = cm
        self.lib = C.CDLL(so_path)
        self.setup = getattr(self.lib, cm.name + "_setup")
        self.eval = getattr(self.lib, cm.name + "_eval")
        self.setup_addr = C.cast(self.setup, C.c_void_p).value
        self.eval_addr = C.cast(self.eval, C.c_void_p).value
        self.setupn = self.noise = None
        self.setupn_addr = self.noise_addr = 0
        self.setupv_addr = self.evalv_addr = 0
        if cm.source_v:
            self.setupv_addr = C.cast(getattr(self.lib, cm.name + "_setupv"), C.c_void_p).value
            self.evalv_addr = C.cast(getattr(self.lib, cm.name + "_evalv"), C.c_void_p).value
        if cm.source_n:
            self.setupn = getattr(self.lib, cm.name + "_setupn")
            self.noise = getattr(self.lib, cm.name + "_noise")
            self.setupn_addr = C.cast(self.setupn, C.c_void_p).value
            self.noise_addr = C.cast(self.noise, C.c_void_p).value

    def shape(self):
        from ..flat import VAModelShape
        cm = self.cm
        # host cache rows = stream slots + uniform slots (CACHE_LDU reads behind the stream)
        return VAModelShape(cm.name, list(cm.terminals), list(cm.params), cm.ncache + cm.nuni, list(cm.jrow), list(cm.jcol),
                            self.setup_addr, self.eval_addr, ncache_n=cm.ncache_n + cm.nuni_n,
                            noise_pos=[int(s[0]) for s in cm.noise_sources], noise_neg=[int(s[1]) for s in cm.noise_sources],
                            host_setupn=self.setupn_addr, host_noise=self.noise_addr, branch_terms=list(cm.branch_terms), linear=bool(cm.linear),
                            ncache_v=cm.ncache_v + cm.nuni_v, host_setupv=self.setupv_addr, host_evalv=self.evalv_addr)

    # convenience for tests
    def run_noise(self, params: dict, v, temp_c=27.0, gmin=1e-12):
        """(pwr[K], exp[K]) of the model's noise sources at terminal voltages v."""
        import numpy as np
        cache = self.run_setup(params, temp_c, gmin, noise=True)
        K = len(self.cm.noise_sources)
        v = np.ascontiguousarray(v, dtype=np.float64)
        pw = np.zeros(max(1, K)); ex = np.zeros(max(1, K))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self.noise(dp(cache), dp(v), dp(pw), dp(ex))
        return pw[:K], ex[:K]

    def run_setup(self, params: dict, temp_c=27.0, gmin=1e-12, noise=False):
        import numpy as np
        cm = self.cm
        lut = {p.lower(): i for i, p in enumerate(cm.params)}
        par = np.zeros(max(1, len(cm.params)))
        given = np.zeros(max(1, len(cm.params)), dtype=np.uint8)
        for k, v in params.items():
            par[lut[k.lower()]] = v
            given[lut[k.lower()]] = 1
        cache = np.zeros(max(1, (cm.ncache_n + cm.nuni_n) if noise else (cm.ncache + cm.nuni)))
        (self.setupn if noise else self.setup)(par.ctypes.data_as(C.POINTER(C.c_double)), given.ctypes.data_as(C.POINTER(C.c_uint8)),
                   C.c_double(temp_c), C.c_double(gmin), cache.ctypes.data_as(C.POINTER(C.c_double)))
        return cache

    def executed_ops(self, cache, v):
        """(add, mul, div, special) source-level FP64 operation counts on the path executed for bias v
        (needs build_host(..., count_ops=True)); special = exp/log/sqrt/pow/trig calls."""
        import numpy as np
        arr = (C.c_double * 4).in_dll(self.lib, "va_ops_")
        for k in range(4):
            arr[k] = 0.0
        self.run_eval(cache, v)
        return tuple(int(arr[k]) for k in range(4))

    def run_eval(self, cache, v):
        import numpy as np
        nt = len(self.cm.terminals)
        v = np.ascontiguousarray(v, dtype=np.float64)
        I = np.zeros(nt); Q = np.zeros(nt); G = np.zeros((nt, nt)); Cm = np.zeros((nt, nt))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self.eval(dp(cache), dp(v), dp(I), dp(Q), dp(G), dp(Cm))
        return I, Q, G, Cm


def build_host(cm: CompiledModel, out_dir: Optional[str] = None, opt: str = "-O2", count_ops: bool = False,
               fast_tag: Optional[str] = None) -> HostModel:
    """fast_tag: build for bench.py's `cpu_fast` baseline arm instead of the checker: -O3 -march=native with the
    compiler's default FP contraction; the tag (a hash of the host CPU's flags) keeps such builds per host."""
    out_dir = out_dir or GEN_DIR
    os.makedirs(out_dir, exist_ok=True)
    # (a value-only variant may still carry OUT_J lines of bias-independent Jacobian entries: it has no G / C arguments)
    text = c_prelude(len(cm.terminals), count_ops) + cm.source + \
        ("#undef OUT_J\n#define OUT_J(idx, k, l, g, c) do { } while (0)\n" + cm.source_v if cm.source_v else "") + (cm.source_n or "")
    flags = [opt, "-ffp-contract=off"] if fast_tag is None else ["-O3", "-march=native"]
    key = hashlib.sha1((text + opt + (fast_tag or "")).encode()).hexdigest()[:16]
    base = os.path.join(out_dir, f"{cm.name}_{key}")
    so = base + ".so"
    if not os.path.exists(so):
        # per-process scratch names + atomic rename: concurrent builders (pytest-xdist workers, ranks) never
        # see or clobber each other's half-written files
        src, tmp = f"{base}.{os.getpid()}.c", f"{so}.{os.getpid()}.tmp"
        with open(src, "w") as f:
            f.write(text)
        cmd = [HOST_CC, *flags, "-fPIC", "-shared", "-fno-math-errno", "-w", "-o", tmp, src, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"host compile of generated model failed:\n{r.stderr[:4000]}")
        os.replace(src, base + ".c")
        os.replace(tmp, so)
    return HostModel(cm, so)
