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

VA_FN double va_limexp(double x) { return x < 80.0 ? exp(x) : exp(80.0) * (1.0 + x - 80.0); }
VA_FN double va_dlimexp(double x) { return x < 80.0 ? exp(x) : exp(80.0); }

#define PAR(i) (par_col_[i] >= 0 ? a.params[(size_t)par_col_[i] * a.B + inst] : par_val_[i])
#define GIVEN(i) (given_[i] != 0)
#define TEMP_K (temp_c_ + 273.15)
#define GMIN_V gmin_
#define CACHE_ST(s, v) cache_[(size_t)(s) * a.B] = (double)(v)
#define CACHE_LD(s) __ldg(cache_ + (size_t)(s) * a.B)
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

#ifndef VA_EVAL_THREADS
#define VA_EVAL_THREADS 128
#endif
#ifndef VA_EVAL_MINBLOCKS
#define VA_EVAL_MINBLOCKS 8
#endif
#ifdef VA_PREFETCH_L2
#define VA_PREFETCH_ALL()                                                                        \
    for (int s_ = 0; s_ < NCACHE; s_++)                                                          \
        asm volatile("prefetch.global.L2 [%0];" ::"l"(cache_ + (size_t)s_ * a.B));
#else
#define VA_PREFETCH_ALL()
#endif
#define VA_EVAL_BEGIN(NAME)                                                                      \
    extern "C" __global__ void __launch_bounds__(VA_EVAL_THREADS, VA_EVAL_MINBLOCKS) k_eval_##NAME(VaArgs a) {      \
        const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;                 \
        if (inst >= a.B) return;                                                                 \
        if (!a.active[inst]) return;                                                             \
        const int dev = blockIdx.y;                                                              \
        const double alpha_ = a.alpha[inst];                                                     \
        const double* __restrict__ cache_ = a.cache + ((size_t)dev * NCACHE) * a.B + inst;       \
        double* __restrict__ out_ = a.out + ((size_t)dev * NOUT) * a.B + inst;                   \
        double vt_[NT];                                                                          \
        _Pragma("unroll") for (int k_ = 0; k_ < NT; k_++) {                                      \
            const int n_ = a.term[dev * NT + k_];                                                \
            vt_[k_] = n_ < 0 ? 0.0 : a.x[(size_t)n_ * a.B + inst];                               \
        }                                                                                        \
        VA_PREFETCH_ALL()
#define VA_EVAL_END(NAME) }
)CUDA";
