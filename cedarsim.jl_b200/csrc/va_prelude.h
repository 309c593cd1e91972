// va_prelude.h -- CUDA-side prelude prepended to the generated Verilog-A model code before
// NVRTC compiles it for sm_100a.  It gives the macro vocabulary of va/compiler.py a device
// meaning: one thread per (sweep point, device instance); consecutive threads are
// consecutive sweep points, so every global access below is a coalesced, batch-interleaved
// [slot][B] access.
#pragma once

static const char* const CB_VA_PRELUDE = R"CUDA(
typedef unsigned char uint8_t;
#define INFINITY __longlong_as_double(0x7ff0000000000000LL)
#define NAN __longlong_as_double(0x7ff8000000000000LL)
#define VA_FN static __device__ __forceinline__
#define VA_OPS(a, m, d, s)

struct VaArgs {
    long long B;                 // sweep points on this device
    const double* x;             // [N][B] current Newton iterate
    const double* alpha;         // [B]    d/dt discretisation coefficient of each point
    const int* list;             // [*count] sweep points that take part in this launch (device-wide compaction by k_control)
    const double* cache;         // [ndev][NCACHE][B] bias-independent values
    double* out;                 // [ndev][NOUT][B]   I | Q | J = dI/dV + alpha dQ/dV
    const int* term;             // [ndev][NT] unknown index per terminal, -1 = ground
    const double* params;        // [P][B] swept parameters
    const double* par_val;       // [ndev][NPARAM]
    const int* par_col;          // [ndev][NPARAM] column of params or -1
    const uint8_t* given;        // [ndev][NPARAM]
    double temp_val; double gmin_val;
    int temp_col; int gmin_col;
    const int* count;            // number of entries of `list` (a device counter: the grid is sized for all B points and
                                 // CTAs beyond the count exit at once)
    double* uni;                 // [NUNI] cached values that are the same for every device of the model (they depend on the
                                 // model card, temperature and gmin only), or [B][NUNI] when temperature / gmin are swept
    int uni_per_inst;
    int want;                    // blocked mode (list == nullptr): thread k is sweep point k and takes part iff active[k] == want
    const int* active;
};

// Branch-free reciprocal, square root, exp, log and pow for the eval stream.  Two reasons: (1) the compiler's own
// x / y, sqrt(), exp(), log(), pow() carry rarely-taken slow paths behind BSSY / BRA / BSYNC, and every one of them
// redirects instruction fetch in these ~20k-instruction straight-line kernels; (2) the library versions materialise
// every polynomial coefficient with two move instructions (exp: 55 instructions, log: 80, pow: 255), which made
// the transcendental expansions about half of the whole device evaluation.  The versions below read their
// coefficients from constant memory (uniform loads, two doubles per instruction): exp ~30, log ~40, pow ~75
// instructions.  Accuracy (checked against libm on the CPU with the same operation sequence, tests/test_vamath.py):
// exp <= 1 ulp, log <= 2 ulp, pow relative error <= 1e-14 for |b ln a| <= 100, 1/x and sqrt <= 2 ulp.
// Deviations from IEEE library behaviour, all outside what compact models evaluate: exp underflows to 0 below
// -708.39 (no denormal results); pow(a, b) = exp(b log |a|) with the sign rule for integer b.
//
// exp / log / pow / reciprocal / sqrt are real calls (__noinline__), not inlined: the eval kernels are bound by
// instruction supply (a ~290 KB straight-line stream that every warp walks once, DESIGN.md section 5), so a smaller
// cold stream beats a shorter dynamic one -- the bodies below stay in the instruction cache.  Measured on the DFF
// bench workload: 18.1k -> 13.0k static SASS instructions per kernel, k_eval -6 %, k_evalv -16 %, +9.5 % points/s
// (profiles/variants_r1v.log).  -DVA_MATH_NOINLINE=0 inlines everything again, =1 keeps 1/x and sqrt inline.
#ifndef VA_MATH_NOINLINE
#define VA_MATH_NOINLINE 2
#endif
#if VA_MATH_NOINLINE >= 1
#define VA_MATH_FN static __device__ __noinline__
#else
#define VA_MATH_FN VA_FN
#endif
#if VA_MATH_NOINLINE >= 2
#define VA_MATH_FN2 static __device__ __noinline__
#else
#define VA_MATH_FN2 VA_FN
#endif
VA_FN double va_rcp_normal(const double x) {   // |x| normal and 1/x normal
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    double e = fma(-x, r0, 1.0);
    e = fma(e, e, e);
    double r = fma(r0, e, r0);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
VA_MATH_FN2 double va_rcp(const double x) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    double e = fma(-x, r0, 1.0);
    e = fma(e, e, e);
    double r = fma(r0, e, r0);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return fabs(r) <= 1.7976931348623157e308 ? r : r0;   // inputs outside the normal range: the seed is the IEEE answer
}
VA_MATH_FN2 double va_sqrt(const double x) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    double g = x * y0, h = 0.5 * y0;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    const double d = fma(-g, g, x);
    g = fma(d, h, g);
    const bool normal = x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308;
    return normal ? g : (x >= 0.0 ? (x < 1.0 ? 0.0 : x) : NAN);
}
// exp(x) = 2^n (1 + r + r^2 q(r)),  n = rint(x log2 e),  r = x - n ln2 (two-part ln2), |r| <= ln2 / 2;
// q = degree-9 near-minimax fit of (e^r - 1 - r) / r^2 (interpolation at Chebyshev nodes, max error 1.3e-17)
__constant__ double va_kexp[14] = {
    1.4426950408889634, 6755399441055744.0, 6.93147180369123816490e-01, 1.90821492927058770002e-10,
    2.5100395159429244017e-8, 2.7620101012098000228e-7, 2.7557268439678002354e-6, 0.000024801521269532121095,
    0.00019841269863066695027, 0.0013888888917230723233, 0.0083333333333300592137, 0.04166666666662409382,
    0.16666666666666667454, 0.50000000000000010231};
VA_MATH_FN double va_exp(const double x) {
    const double* __restrict__ K = va_kexp;
    double t = fma(x, K[0], K[1]);
    const int n = __double2loint(t);
    t -= K[1];
    double r = fma(-t, K[2], x);
    r = fma(-t, K[3], r);
    double p = K[4];
    p = fma(p, r, K[5]); p = fma(p, r, K[6]); p = fma(p, r, K[7]); p = fma(p, r, K[8]); p = fma(p, r, K[9]);
    p = fma(p, r, K[10]); p = fma(p, r, K[11]); p = fma(p, r, K[12]); p = fma(p, r, K[13]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    p *= __hiloint2double((n + 1023) << 20, 0);
    p = x < -708.39 ? 0.0 : p;                 // 2^n would be subnormal (also covers -inf)
    return x > 709.78 ? INFINITY : p;          // NaN falls through as NaN
}
// log(x) = e ln2 + 2 atanh(f),  x = 2^e m,  m in [sqrt(1/2), sqrt(2)),  f = (m - 1) / (m + 1),  s = f^2,
// 2 atanh(f) = 2 f + f s P(s);  P = degree-7 fit of 2 (atanh(sqrt s) / sqrt s - 1) / s on [0, 0.0295] (error 6e-20)
__constant__ double va_klog[10] = {
    6.93147180369123816490e-01, 1.90821492927058770002e-10,
    0.1308803485092755383, 0.13268638369266823879, 0.15386244737621666606, 0.18181795547521136235,
    0.22222222392581045097, 0.28571428570799785031, 0.40000000000000883236, 0.66666666666666666463};
VA_MATH_FN double va_log(const double x0) {
    const double* __restrict__ K = va_klog;
    const bool tiny = x0 < 2.2250738585072014e-308;
    const double x = tiny ? x0 * 18014398509481984.0 : x0;    // subnormal inputs: scale by 2^54
    int hi = __double2hiint(x);
    int e = (hi >> 20) - (tiny ? 1023 + 54 : 1023);
    hi = (hi & 0x000fffff) | 0x3ff00000;
    const bool up = hi >= 0x3ff6a09f;                        // m >= sqrt(2): halve m
    hi = up ? hi - 0x00100000 : hi;
    e = up ? e + 1 : e;
    const double m = __hiloint2double(hi, __double2loint(x));
    const double f = (m - 1.0) * va_rcp_normal(m + 1.0);
    const double s = f * f;
    double p = K[2];
    p = fma(p, s, K[3]); p = fma(p, s, K[4]); p = fma(p, s, K[5]); p = fma(p, s, K[6]); p = fma(p, s, K[7]);
    p = fma(p, s, K[8]); p = fma(p, s, K[9]);
    const double de = (double)e;
    double r = fma(s * f, p, de * K[1]);
    r = fma(2.0, f, r);
    r = fma(de, K[0], r);
    const bool ok = x0 > 0.0 && x0 <= 1.7976931348623157e308;
    return ok ? r : (x0 == 0.0 ? -INFINITY : (x0 > 0.0 ? x0 : NAN));
}
VA_MATH_FN double va_pow(const double a, const double b) {
    const double t = b * va_log(fabs(a));
    const double r = va_exp(b == 0.0 ? 0.0 : t);         // pow(a, 0) = 1 for every a, including 0 and inf
    // negative base: defined for integer exponents only, sign by parity
    const double hb = 0.5 * b;
    const bool b_int = rint(b) == b, b_odd = b_int && rint(hb) != hb;
    return a < 0.0 ? (b_int ? (b_odd ? -r : r) : NAN) : r;
}
#ifdef VA_EXACT_DIV
#define VA_RCP(x) (1.0 / (x))
#define VA_SQRT(x) sqrt(x)
#else
#define VA_RCP(x) va_rcp(x)
#define VA_SQRT(x) va_sqrt(x)
#endif
// The remaining libm functions a compact model calls now and then (BSIM-CMG: tan / cos / sin in the initial guess of
// its surface-potential iteration, ~200 inlined SASS instructions per site with their slow paths) are the CUDA library
// versions behind one out-of-line body each: same values, ~1 000 fewer instructions in every BSIM-CMG eval kernel.
#if VA_MATH_NOINLINE >= 1
VA_MATH_FN double va_sin(const double x) { return sin(x); }
VA_MATH_FN double va_cos(const double x) { return cos(x); }
VA_MATH_FN double va_tan(const double x) { return tan(x); }
VA_MATH_FN double va_tanh(const double x) { return tanh(x); }
VA_MATH_FN double va_atan(const double x) { return atan(x); }
#define sin(x) va_sin(x)
#define cos(x) va_cos(x)
#define tan(x) va_tan(x)
#define tanh(x) va_tanh(x)
#define atan(x) va_atan(x)
#endif
#ifndef VA_LIBM
#define exp(x) va_exp(x)
#define log(x) va_log(x)
#define pow(a, b) va_pow(a, b)
#endif

// d/da a^b = b a^(b-1) given p = a^b (finite at a == 0 for b >= 1; b == 0 -> 0 avoids 0 * inf): an IEEE division and a
// rarely-taken second pow, ~50 instructions when inlined at every site
VA_MATH_FN double va_dpow(const double a, const double b, const double p) {
    return b == 0.0 ? 0.0 : (a == 0.0 ? b * pow(a, b - 1.0) : b * p / a);
}
#define VA_DPOW(a, b, p) va_dpow(a, b, p)

VA_FN double va_limexp(double x) { return x < 80.0 ? exp(x) : exp(80.0) * (1.0 + x - 80.0); }
VA_FN double va_dlimexp(double x) { return x < 80.0 ? exp(x) : exp(80.0); }

#define PAR(i) (par_col_[i] >= 0 ? a.params[(size_t)par_col_[i] * a.B + inst] : par_val_[i])
#define GIVEN(i) (given_[i] != 0)
#define TEMP_K (temp_c_ + 273.15)
#define GMIN_V gmin_
#define CACHE_ST(s, v) cache_[VA_SLOT_OFF(s)] = (double)(v)
// uniform slots: written by device 0 of the model (by its first point only, unless the table is per point)
#define CACHE_STU(k, v) { if (uni_on_) uni_w_[(k)] = (double)(v); }
#define CACHE_LDU(k) __ldg(uni_ + (k))
#define VT(k) vt_[k]
#define OUT_I(k, v) out_[(size_t)(k) * a.B] = (v)
#define OUT_Q(k, v) out_[(size_t)(NT + (k)) * a.B] = (v)
// J = dI/dV + alpha dQ/dV for the Newton matrix, and dQ/dV on its own (rows 2 NT + NJ ...) for the first-order
// charge update q(x + dx) ~ q(x) + C dx of k_lu
#define OUT_J(idx, k, l, g, c) { const double c_ = (c); out_[(size_t)(2 * NT + (idx)) * a.B] = (g) + alpha_ * c_; \
                                 out_[(size_t)(2 * NT + NJ + (idx)) * a.B] = c_; }

// Per-instance cache: NCACHE_P = NCACHE rounded up to whole 32-byte sectors of doubles per (device, point), in the order
// the eval function consumes it.  Layouts (VA_CACHE_LAYOUT, the engine only sizes the allocation):
//   0  [device][block of 128 points][slot][128]          batch-interleaved: a warp of consecutive points reads whole rows,
//                                                         but a launch over a point LIST with gaps fetches every sector
//                                                         for the fraction of its bytes that belong to listed points
//   1  [device][point][slot]                              one contiguous row per point, 8-byte copies
//   2  [device][point][slot]                              ... 16-byte copies that bypass L1 (cp.async.cg)
//   3  [device][block of 128][chunk of 4 slots][128][4]   SECTOR-interleaved: the 4-slot chunk of a point is one 32-byte
//                                                         DRAM sector, the sectors of 128 consecutive points are contiguous;
//                                                         two 16-byte cp.async.cg per chunk.  HBM traffic = the rows of the
//                                                         listed points only, and a warp request still spans few lines
#ifndef VA_CACHE_LAYOUT
#define VA_CACHE_LAYOUT 0
#endif
#ifndef VA_CACHE_LAYOUT_V
#define VA_CACHE_LAYOUT_V VA_CACHE_LAYOUT   // layout of the value-only variant's cache (its launches have gaps in lock-step rounds)
#endif
// VA_LAYOUT: the layout of the variant being compiled; the engine redefines it between the full / value-only / noise
// blocks of a model (engine.py cuda_source), every macro below is evaluated where the generated code expands it.
#define VA_LAYOUT VA_CACHE_LAYOUT
#define VA_CACHE_BLK 128   // rows are allocated for B rounded up to a multiple of this
#define NCACHE_P ((NCACHE + 3) / 4 * 4)
template <int L>
VA_FN constexpr int va_slot_off(const int s) {
    return L == 0 ? s * VA_CACHE_BLK : L == 3 ? (s / 4) * (4 * VA_CACHE_BLK) + (s % 4) : s;
}
#define VA_SLOT_OFF(s) va_slot_off<VA_LAYOUT>(s)
template <int L>
VA_FN size_t va_cache_index(const long long B, const int dev, const int ncp, const long long inst) {
    const size_t nblk = (size_t)((B + VA_CACHE_BLK - 1) / VA_CACHE_BLK);
    if (L == 0) return (((size_t)dev * nblk + (size_t)(inst / VA_CACHE_BLK)) * ncp) * VA_CACHE_BLK + (size_t)(inst % VA_CACHE_BLK);
    if (L == 3) return (((size_t)dev * nblk + (size_t)(inst / VA_CACHE_BLK)) * ncp) * VA_CACHE_BLK + (size_t)(inst % VA_CACHE_BLK) * 4;
    return ((size_t)dev * (nblk * VA_CACHE_BLK) + (size_t)inst) * (size_t)ncp;
}
#define VA_SETUP_BEGIN(NAME) VA_SETUP_BEGIN_(k_setup_##NAME)
#define VA_SETUPV_BEGIN(NAME) VA_SETUP_BEGIN_(k_setupv_##NAME)
#define VA_SETUPV_END(NAME) }
#define VA_SETUP_BEGIN_(KERNEL)                                                                  \
    extern "C" __global__ void __launch_bounds__(128) KERNEL(VaArgs a) {                         \
        const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;                 \
        if (inst >= a.B) return;                                                                 \
        const int dev = blockIdx.y;                                                              \
        const double* par_val_ = a.par_val + (size_t)dev * NPARAM;                               \
        const int* par_col_ = a.par_col + (size_t)dev * NPARAM;                                  \
        const uint8_t* given_ = a.given + (size_t)dev * NPARAM;                                  \
        const double temp_c_ = a.temp_col >= 0 ? a.params[(size_t)a.temp_col * a.B + inst] : a.temp_val; \
        const double gmin_ = a.gmin_col >= 0 ? a.params[(size_t)a.gmin_col * a.B + inst] : a.gmin_val;   \
        double* cache_ = (double*)a.cache + va_cache_index<VA_LAYOUT>(a.B, dev, NCACHE_P, inst);              \
        double* uni_w_ = a.uni + (a.uni_per_inst ? (size_t)inst * NUNI : 0);                      \
        const bool uni_on_ = dev == 0 && (a.uni_per_inst || inst == 0);                          \
        (void)uni_w_; (void)uni_on_;                                                             \
        (void)gmin_; (void)temp_c_; (void)par_val_; (void)par_col_; (void)given_;
#define VA_SETUP_END(NAME) }

// ---- cache streaming -------------------------------------------------------------------------
// The generator lays the per-instance cache out as a stream in consumption order and marks, between
// top-level statements, where chunk k (VA_CHUNK_ROWS stream positions) is first needed.  Every
// thread copies ITS OWN column of a chunk into a shared-memory ring with 8-byte cp.async copies,
// VA_AHEAD chunks ahead of the consumer: the HBM latency of the ~250 scattered cache reads is taken
// off the dependency chain without holding the values in registers.  A thread only ever reads what
// it copied itself, so there is no block-level barrier and divergent / retired threads are harmless.
// Ring safety (chunk k + VA_AHEAD overwrites chunk k + VA_AHEAD - VA_STAGES while positions down to
// need - VA_WINDOW + 1 may still be read): (VA_STAGES - VA_AHEAD - 1) * VA_CHUNK_ROWS >= VA_WINDOW - 1.
#ifndef VA_EVAL_THREADS
#define VA_EVAL_THREADS 128
#endif
#ifndef VA_EVAL_MINBLOCKS
#define VA_EVAL_MINBLOCKS 5
#endif
// Ring geometry: 4-row chunks, 1 chunk ahead, 4 stages = 16 rows = 16 KB per 128-thread CTA.  Small on purpose: the
// eval kernels spill (~900 B of local memory per thread at 96 registers) and what the ring does not take of the SM's
// 256 KB is L1 for those spills -- the engine sets the shared-memory carve-out of these kernels to what their resident
// CTAs need (META[3] = CTAs per SM).  A 40 KB ring (8-row chunks, 2 ahead, 5 stages) measured 3.5 % slower end to end
// and 7 % slower for k_eval alone (profiles/variants_r1aa.log); 4 rows ahead are ~6 000 cycles of lead, several HBM
// latencies.
#ifndef VA_AHEAD
#define VA_AHEAD 1
#endif
#ifndef VA_STAGES
#define VA_STAGES 4
#endif
// ring geometry of the value-only variant (no spills to keep L1 for, little arithmetic per cache row: a deeper ring
// hides more of the row latency; ncu: 90 % of its shared-memory-read stalls wait for the row copies)
#ifndef VA_AHEAD_V
#define VA_AHEAD_V VA_AHEAD
#endif
#ifndef VA_STAGES_V
#define VA_STAGES_V VA_STAGES
#endif

// VA_CACHE_HINT=1: the cache rows are read once per launch and never again before ~1 GB of other rows has passed: mark
// them evict-first in L2 so that they do not push out the lines that ARE re-used (register spills of the eval kernels,
// the device outputs k_lu reads next).
#ifndef VA_CACHE_HINT
#define VA_CACHE_HINT 1   // measured +3 % points/s on the bench workload (profiles/probe_r2i.log: 4 733 -> 4 874)
#endif
VA_FN void va_cp8(unsigned dst, const double* src) {
#if VA_CACHE_HINT
    unsigned long long pol_;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_));   // not volatile: one per kernel after CSE
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "l"(pol_) : "memory");
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
#endif
}
VA_FN void va_cp16(unsigned dst, const double* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// Ring geometry in shared memory.  8-byte copies: [stage][row][thread] (a thread's slots are NTHR doubles apart,
// conflict-free).  16-byte copies: [stage][row pair][thread][2].
#define VA_RING_IDX(s) (VA_LAYOUT >= 2 ? (((((s) / VA_CHUNK_ROWS) % va_stages_) * (VA_CHUNK_ROWS / 2) + ((s) % VA_CHUNK_ROWS) / 2) * (2 * VA_EVAL_THREADS) + ((s) & 1)) \
                                       : (((((s) / VA_CHUNK_ROWS) % va_stages_) * VA_CHUNK_ROWS + (s) % VA_CHUNK_ROWS) * VA_EVAL_THREADS))
#define VA_RING_TID(t) (VA_LAYOUT >= 2 ? 2 * (t) : (t))
template <int L, int ROWS, int STAGES, int NTHR>
VA_FN void va_issue(const int chunk, const int ncache, const unsigned sbase, const double* cache) {
    if (L >= 2) {
        static_assert(ROWS % 2 == 0, "16-byte cache copies need an even number of rows per chunk");
#pragma unroll
        for (int r = 0; r < ROWS; r += 2) {
            const int p = chunk * ROWS + r;
            if (p < ncache) va_cp16(sbase + (unsigned)((((chunk % STAGES) * (ROWS / 2) + r / 2) * (2 * NTHR)) * 8), cache + va_slot_off<L>(p));
        }
    } else {
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
            const int p = chunk * ROWS + r;
            if (p < ncache) va_cp8(sbase + (unsigned)((((chunk % STAGES) * ROWS + r) * NTHR) * 8), cache + va_slot_off<L>(p));
        }
    }
}
// (Experiments that did not pay and were removed: CTA-wide barriers at the chunk markers or every ~50 generated
// lines, and a leader warp running one chunk ahead, to make the warps of a CTA share instruction fetches -- at
// 128..640 threads per CTA none changed the instruction-cache request count; see DESIGN.md section 5.)
#define VA_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define VA_CHUNK(k)                                                                              \
    {                                                                                            \
        if ((k) + va_ahead_ < VA_NCHUNK)                                                         \
            va_issue<VA_LAYOUT, VA_CHUNK_ROWS, va_stages_, VA_EVAL_THREADS>((k) + va_ahead_, NCACHE_P, sbase_, cache_); \
        VA_COMMIT();                                                                             \
        asm volatile("cp.async.wait_group %0;" ::"n"(va_ahead_) : "memory");                      \
    }
#define CACHE_LD(s) ring_[VA_RING_IDX(s)]
#define CACHE_LDG(s) __ldg(cache_ + VA_SLOT_OFF(s))

// value-only variant (k_evalv_*: currents and charges, no Jacobian; its own cache, see CompiledModel.source_v)
#ifndef VA_EVALV_MINBLOCKS
#define VA_EVALV_MINBLOCKS 5
#endif
// noise variant (k_setupn_* / k_evaln_*: power and flicker exponent of every noise source at the operating point,
// rows [source] and [NNOISE + source] of its own output block; its own cache; used by cb_noise only)
#define VA_SETUPN_BEGIN(NAME) VA_SETUP_BEGIN_(k_setupn_##NAME)
#define VA_SETUPN_END(NAME) }
#define VA_EVALN_BEGIN(NAME) VA_EVAL_BEGIN_(k_evaln_##NAME, va_metan_##NAME, 1, VA_AHEAD, VA_STAGES)
#define VA_EVALN_END(NAME) VA_EVAL_END(NAME)
#define OUT_N(k, v) out_[(size_t)(k) * a.B] = (v)
#define OUT_NE(k, v) out_[(size_t)(NNOISE + (k)) * a.B] = (v)
#define VA_EVAL_BEGIN(NAME) VA_EVAL_BEGIN_(k_eval_##NAME, va_meta_##NAME, VA_EVAL_MINBLOCKS, VA_AHEAD, VA_STAGES)
#define VA_EVALV_BEGIN(NAME) VA_EVAL_BEGIN_(k_evalv_##NAME, va_metav_##NAME, VA_EVALV_MINBLOCKS, VA_AHEAD_V, VA_STAGES_V)
#define VA_EVALV_END(NAME) VA_EVAL_END(NAME)
// Thread mapping: thread k of the launch takes the k-th entry of the point list of this kind of iteration (full /
// value-only), which k_control compacted device-wide; threads beyond the count exit before any work.  The lists ascend
// inside runs of up to 32 points, so a warp reads a few contiguous row segments; with every point live the launch is as
// coalesced as the identity mapping.
#define VA_EVAL_BEGIN_(KERNEL, META, MINBLOCKS, AHEAD, STAGES)                                   \
    extern "C" __device__ int META[6] = {VA_EVAL_THREADS, (STAGES) * VA_CHUNK_ROWS * VA_EVAL_THREADS * 8, NCACHE, MINBLOCKS, NUNI, 0}; \
    extern "C" __global__ void __launch_bounds__(VA_EVAL_THREADS, MINBLOCKS) KERNEL(VaArgs a) {  \
        constexpr int va_ahead_ = (AHEAD), va_stages_ = (STAGES);   /* ring geometry of this kernel (VA_CHUNK, CACHE_LD) */ \
        static_assert((va_stages_ - va_ahead_ - 1) * VA_CHUNK_ROWS >= VA_WINDOW - 1, "cache ring too shallow"); \
        extern __shared__ __align__(16) double va_ring_[];                                       \
        if (blockDim.x != VA_EVAL_THREADS) __trap();                                             \
        long long inst;                                                                          \
        {                                                                                        \
            const long long k_ = (long long)blockIdx.x * VA_EVAL_THREADS + threadIdx.x;          \
            if (a.list) {                                                                        \
                if (k_ >= (long long)*a.count) return;                                           \
                inst = a.list[k_];                                                               \
            } else {   /* blocked mode: no lists, the point's role decides */                    \
                if (k_ >= a.B || a.active[k_] != a.want) return;                                 \
                inst = k_;                                                                       \
            }                                                                                    \
        }                                                                                        \
        const int dev = blockIdx.y;                                                              \
        const double* __restrict__ cache_ = a.cache + va_cache_index<VA_LAYOUT>(a.B, dev, NCACHE_P, inst);    \
        const double* __restrict__ uni_ = a.uni + (a.uni_per_inst ? (size_t)inst * NUNI : 0);    \
        (void)uni_;                                                                              \
        const double* ring_ = va_ring_ + VA_RING_TID(threadIdx.x);                               \
        const unsigned sbase_ = (unsigned)__cvta_generic_to_shared(va_ring_ + VA_RING_TID(threadIdx.x)); \
        _Pragma("unroll") for (int c_ = 0; c_ < va_ahead_; c_++) {                               \
            if (c_ < VA_NCHUNK) va_issue<VA_LAYOUT, VA_CHUNK_ROWS, va_stages_, VA_EVAL_THREADS>(c_, NCACHE_P, sbase_, cache_); \
            VA_COMMIT();                                                                         \
        }                                                                                        \
        const double alpha_ = a.alpha[inst];                                                     \
        double* __restrict__ out_ = a.out + ((size_t)dev * NOUT) * a.B + inst;                   \
        double vt_[NT];                                                                          \
        _Pragma("unroll") for (int k_ = 0; k_ < NT; k_++) {                                      \
            const int n_ = a.term[dev * NT + k_];                                                \
            vt_[k_] = n_ < 0 ? 0.0 : a.x[(size_t)n_ * a.B + inst];                               \
        }
#define VA_EVAL_END(NAME) }
)CUDA";
