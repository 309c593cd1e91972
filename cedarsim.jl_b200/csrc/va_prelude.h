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
    const int* active;           // [B]    0 = point finished, skip
    const double* cache;         // [ndev][NCACHE][B] bias-independent values
    double* out;                 // [ndev][NOUT][B]   I | Q | J = dI/dV + alpha dQ/dV
    const int* term;             // [ndev][NT] unknown index per terminal, -1 = ground
    const double* params;        // [P][B] swept parameters
    const double* par_val;       // [ndev][NPARAM]
    const int* par_col;          // [ndev][NPARAM] column of params or -1
    const uint8_t* given;        // [ndev][NPARAM]
    double temp_val; double gmin_val;
    int temp_col; int gmin_col;
};

// Branch-free reciprocal and square root for the eval stream.  The compiler's own x / y and sqrt() carry a
// rarely-taken slow path behind BSSY / BRA / BSYNC: every one of them redirects instruction fetch, and in
// these ~25k-instruction straight-line kernels fetch is the bottleneck (BSYNC alone drew 19 % of the stall
// samples).  MUFU seed (>= 20 bits) + the same Newton steps the compiler emits; inputs outside the normal
// range fall back to the raw seed, which is already the IEEE answer there (inf, 0, nan).  Result within
// 1 ulp of the correctly rounded one.
VA_FN double va_rcp(const double x) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    double e = fma(-x, r0, 1.0);
    e = fma(e, e, e);
    double r = fma(r0, e, r0);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return fabs(r) <= 1.7976931348623157e308 ? r : r0;
}
VA_FN double va_sqrt(const double x) {
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
#ifdef VA_EXACT_DIV
#define VA_RCP(x) (1.0 / (x))
#define VA_SQRT(x) sqrt(x)
#else
#define VA_RCP(x) va_rcp(x)
#define VA_SQRT(x) va_sqrt(x)
#endif

VA_FN double va_limexp(double x) { return x < 80.0 ? exp(x) : exp(80.0) * (1.0 + x - 80.0); }
VA_FN double va_dlimexp(double x) { return x < 80.0 ? exp(x) : exp(80.0); }

#define PAR(i) (par_col_[i] >= 0 ? a.params[(size_t)par_col_[i] * a.B + inst] : par_val_[i])
#define GIVEN(i) (given_[i] != 0)
#define TEMP_K (temp_c_ + 273.15)
#define GMIN_V gmin_
#define CACHE_ST(s, v) cache_[(size_t)(s) * a.B] = (double)(v)
#define VT(k) vt_[k]
#define OUT_I(k, v) out_[(size_t)(k) * a.B] = (v)
#define OUT_Q(k, v) out_[(size_t)(NT + (k)) * a.B] = (v)
#define OUT_J(idx, k, l, g, c) out_[(size_t)(2 * NT + (idx)) * a.B] = (g) + alpha_ * (c)

#define VA_SETUP_BEGIN(NAME)                                                                     \
    extern "C" __global__ void __launch_bounds__(128) k_setup_##NAME(VaArgs a) {                 \
        const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;                 \
        if (inst >= a.B) return;                                                                 \
        const int dev = blockIdx.y;                                                              \
        const double* par_val_ = a.par_val + (size_t)dev * NPARAM;                               \
        const int* par_col_ = a.par_col + (size_t)dev * NPARAM;                                  \
        const uint8_t* given_ = a.given + (size_t)dev * NPARAM;                                  \
        const double temp_c_ = a.temp_col >= 0 ? a.params[(size_t)a.temp_col * a.B + inst] : a.temp_val; \
        const double gmin_ = a.gmin_col >= 0 ? a.params[(size_t)a.gmin_col * a.B + inst] : a.gmin_val;   \
        double* cache_ = (double*)a.cache + ((size_t)dev * NCACHE) * a.B + inst;                 \
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
#ifndef VA_AHEAD
#define VA_AHEAD 2
#endif
#ifndef VA_STAGES
#define VA_STAGES 5
#endif

VA_FN void va_cp8(unsigned dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
template <int ROWS, int STAGES, int NTHR>
VA_FN void va_issue(const int chunk, const int ncache, const unsigned sbase, const double* cache, const size_t B) {
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const int p = chunk * ROWS + r;
        if (p < ncache) va_cp8(sbase + (unsigned)((((chunk % STAGES) * ROWS + r) * NTHR) * 8), cache + (size_t)p * B);
    }
}
// The eval body is ~25k straight-line instructions (hundreds of KB of SASS, far beyond the 32 KB L1.5
// instruction cache): a warp that runs alone streams all of it from L2.  A block-wide barrier at every
// chunk marker keeps the warps of a CTA within one chunk of each other, so one instruction fetch
// serves all of them.  (Threads of finished sweep points have exited; barriers count live threads.)
#ifndef VA_CONVOY
#define VA_CONVOY 0
#endif
#if VA_CONVOY
#define VA_CONVOY_SYNC() __syncthreads()
#else
#define VA_CONVOY_SYNC()
#endif
#define VA_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define VA_CHUNK(k)                                                                              \
    {                                                                                            \
        if ((k) + VA_AHEAD < VA_NCHUNK)                                                          \
            va_issue<VA_CHUNK_ROWS, VA_STAGES, VA_EVAL_THREADS>((k) + VA_AHEAD, NCACHE, sbase_, cache_, (size_t)a.B); \
        VA_COMMIT();                                                                             \
        asm volatile("cp.async.wait_group %0;" ::"n"(VA_AHEAD) : "memory");                      \
        VA_CONVOY_SYNC();                                                                        \
    }
#define CACHE_LD(s) ring_[((((s) / VA_CHUNK_ROWS) % VA_STAGES) * VA_CHUNK_ROWS + (s) % VA_CHUNK_ROWS) * VA_EVAL_THREADS]
#define CACHE_LDG(s) __ldg(cache_ + (size_t)(s) * a.B)

#define VA_EVAL_BEGIN(NAME)                                                                      \
    extern "C" __device__ int va_meta_##NAME[2] = {VA_EVAL_THREADS, VA_STAGES * VA_CHUNK_ROWS * VA_EVAL_THREADS * 8}; \
    extern "C" __global__ void __launch_bounds__(VA_EVAL_THREADS, VA_EVAL_MINBLOCKS) k_eval_##NAME(VaArgs a) {      \
        static_assert((VA_STAGES - VA_AHEAD - 1) * VA_CHUNK_ROWS >= VA_WINDOW - 1, "cache ring too shallow"); \
        extern __shared__ double va_ring_[];                                                     \
        if (blockDim.x != VA_EVAL_THREADS) __trap();                                             \
        const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;                 \
        if (inst >= a.B) return;                                                                 \
        if (!a.active[inst]) return;                                                             \
        const int dev = blockIdx.y;                                                              \
        const double* __restrict__ cache_ = a.cache + ((size_t)dev * NCACHE) * a.B + inst;       \
        const double* ring_ = va_ring_ + threadIdx.x;                                            \
        const unsigned sbase_ = (unsigned)__cvta_generic_to_shared(va_ring_ + threadIdx.x);      \
        _Pragma("unroll") for (int c_ = 0; c_ < VA_AHEAD; c_++) {                                \
            if (c_ < VA_NCHUNK) va_issue<VA_CHUNK_ROWS, VA_STAGES, VA_EVAL_THREADS>(c_, NCACHE, sbase_, cache_, (size_t)a.B); \
            VA_COMMIT();                                                                         \
        }                                                                                        \
        const double alpha_ = a.alpha[inst];                                                     \
        if (VA_CONVOY == 2 && (threadIdx.x >> 5) != 0) __syncthreads();                          \
        double* __restrict__ out_ = a.out + ((size_t)dev * NOUT) * a.B + inst;                   \
        double vt_[NT];                                                                          \
        _Pragma("unroll") for (int k_ = 0; k_ < NT; k_++) {                                      \
            const int n_ = a.term[dev * NT + k_];                                                \
            vt_[k_] = n_ < 0 ? 0.0 : a.x[(size_t)n_ * a.B + inst];                               \
        }
#define VA_EVAL_END(NAME) if (VA_CONVOY == 2 && (threadIdx.x >> 5) == 0) __syncthreads(); }
)CUDA";
