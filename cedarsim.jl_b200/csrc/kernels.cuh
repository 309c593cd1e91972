// kernels.cuh -- hand-written sm_100a kernels of the batched Newton / transient engine.
//
// One lock-step "round" advances every unfinished sweep point by one Newton iteration:
//   k_eval_<model> / k_evalv_<model>  (generated, va/compiler.py)  device currents, charges (+ Jacobian stamps)
//   k_lu<false> / k_lu<true>          assembly + batched sparse LU + solves / solves only with the stored factors
//   k_control                         Newton update, convergence, DC / transient state machine of every point
// An ITERATION of a point is either FULL (devices evaluated with their Jacobians, matrix refactored) or VALUE-ONLY
// (devices evaluated without derivatives, chord iteration with the factors of the point's last full iteration).  Every
// step attempt starts with a full iteration.  k_control decides the kind of every point's next iteration and compacts
// the points of each kind DEVICE-WIDE into two dense lists (list_full / list_any + counters, double-buffered by round
// parity); the eval kernels and k_lu of the next round run over those lists, so their launches are dense whatever
// fraction of the points takes part.  See solve() in cedarb200.cu for the two schedules (mixed rounds: every unfinished
// point iterates in every round; lock-step rounds: full round, then value_rounds value-only rounds).
//
// Layout: every per-point array is [k][B] (batch-interleaved, B fastest): consecutive threads are
// consecutive sweep points, so a warp access is one contiguous 256-byte row segment.
//
// The step-control algorithm is the engine's own (the reference delegates it to Sundials
// IDA / OrdinaryDiffEq); the CPU oracle restates the same algorithm for
// parity checks.
#pragma once
#include <cstdint>

namespace cbk {

// PH_REINIT: consistent re-initialisation at t0 between the operating point and the first step (cb_options.t0_reinit):
// a Newton solve of the backward-Euler equations of a step of vanishing length h0 with the sources at their transient
// values -- charges are held, algebraic unknowns follow the sources (reference src/dcop.jl:146-153)
enum { PH_DC = 0, PH_TRAN_INIT = 1, PH_TRAN = 2, PH_DONE = 3, PH_REINIT = 4 };
// integer per-point state rows
enum { IS_PHASE = 0, IS_IT, IS_STAGE, IS_NH, IS_BPI, IS_KSTEP, IS_STATUS, IS_HITBP, IS_METHOD, IS_NP,
       IS_SIDX, IS_NNEWTON, IS_NACC, IS_NREJ, IS_RETRY, IS_NFULL, IS_SRCSTEP, IS_NGROWTH, IS_COUNT };
// double per-point state rows
enum { DS_T = 0, DS_TNEW, DS_H, DS_H1, DS_H2, DS_HPROP, DS_GSHUNT, DS_LIM, DS_NRM, DS_KAPPA, DS_COUNT };
// a.active[]: the point's role in the coming round: 0 = finished, 1 = value-only iteration (stored factors), 2 = full
// iteration (fresh Jacobian: DC iterations, first iteration of a step attempt, every vcycle-th iteration), 3 = idle:
// waits for the next full round (lock-step schedule only; in mixed rounds nobody idles)
enum { ACT_DONE = 0, ACT_ANY = 1, ACT_FULL = 2, ACT_IDLE = 3 };
// compacted point lists of one round: counters cnt[0] = full, cnt[1] = value-only, cnt[2] = idle (the idle list exists
// only when the control step is fused into k_lu, which visits the points of its lists instead of scanning all B)
struct Lists { const int* full; const int* any; const int* idle; const int* cnt; };

struct Pref { double value; int col; int pad; };

struct WaveDev {
    int kind, has_dc;
    Pref dc;
    int npts, pad;
    const double* t;
    const Pref* y;
    Pref v[7];
};

struct Opts {
    double reltol, vabstol, iabstol, nr_reltol, nr_vabstol, nr_iabstol, dc_abstol, dv_max;
    double dt, dt_min, dt_max, t0, t1, teps, span, kappa0, kappa_floor;
    int max_newton_dc, max_newton_tran, method, fixed_step, gmin_steps, skip_dc, dc_only, rate_test, source_steps, reinit;
    long long nfixed, nsave;
};

struct NArgs {
    long long B;
    int N, NV, nnz_lu, O, nwaves, nbp, sm_stride, pad0;
    // symbolic schedule
    const int *diag_pos, *l_ptr, *l_pos, *l_row, *u_ptr, *u_pos, *pair_ptr, *pair_dst;
    const int *uc_ptr, *uc_pos, *uc_row, *row_to_step, *col_to_step;
    // assembly of J (per LU entry) and of the residual (per original row)
    const int *a_ptr, *a_src, *a_lin;
    const double* a_mult;
    const unsigned char* a_diag;
    const int *ri_ptr, *ri_src, *rq_ptr, *rq_src, *rl_ptr, *rl_col, *rl_lin, *rs_ptr, *rs_wave;
    const double *ri_mult, *rq_mult, *rs_coef;
    const double *lin_g, *lin_c;
    long long lin_inst_stride;  // 0 when no linear device value is swept
    long long lin_ent_stride;   // B or 1
    const WaveDev* waves;
    const double* bp;
    const unsigned char* lte_mask;
    const int* outputs;
    const double* saveat;
    const double* params;
    // per-point state, all [k][B]
    double *X, *XN, *X1, *X2, *XP, *QN, *Q1, *QD, *BETA, *alpha, *dst;
    int *ist, *active;
    double* scratch;  // GLOB kernels: [nnz_lu + 3N + nwaves][B]
    const double* dev_out;
    double* y_out;  // tran: [O][S][B]; dc: [O][B]
    int* done_count;
    Opts o;
};

// IEEE division as ONE out-of-line body (~40 instructions with its slow path) instead of ~25 inlined instructions at each
// of the ~70 division sites of the control step, no unrolling of its loops, out-of-line waveform / interpolation helpers:
// k_control shrinks from 7.8 k to 4.6 k SASS instructions.  Same results bit for bit; measured flat in time at 2 048 and
// 16 384 points (profiles/probe_r2u.log) -- kept for the smaller footprint, -DCB_CTRL_SMALL_CODE=0 restores the inlined form.
#ifndef CB_CTRL_SMALL_CODE
#define CB_CTRL_SMALL_CODE 1
#endif
#if CB_CTRL_SMALL_CODE
__device__ __noinline__ double cb_ddiv(double a, double b) { return a / b; }
#define DIV(a, b) cb_ddiv((a), (b))
#define CB_CTRL_INLINE __noinline__
#define CB_CTRL_UNROLL _Pragma("unroll 1")
#else
#define DIV(a, b) ((a) / (b))
#define CB_CTRL_INLINE __forceinline__
#define CB_CTRL_UNROLL _Pragma("unroll 4")
#endif

__device__ __forceinline__ double pv(const Pref& p, const double* params, long long B, long long inst) {
    return p.col < 0 ? p.value : params[(size_t)p.col * B + inst];
}

// reference src/spectre_env.jl:15-21 and 43-69, 1-based index i as there
__device__ inline double pwl_eval(const double* ts, const Pref* ys, int n, double t, const double* params,
                                  long long B, long long inst) {
    int lo = 0, hi = n;  // first index with ts[idx] >= t
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ts[mid] < t) lo = mid + 1; else hi = mid;
    }
    int i = lo + 1;
    if (i <= n && ts[i - 1] == t) i += 1;
    if (i <= 1) return pv(ys[0], params, B, inst);
    if (i > n) return pv(ys[n - 1], params, B, inst);
    const double y0 = pv(ys[i - 2], params, B, inst), y1 = pv(ys[i - 1], params, B, inst);
    if (y0 == y1) return y1;
    if (ts[i - 1] == ts[i - 2]) return 0.5 * (y0 + y1);
    const double slope = (y1 - y0) / (ts[i - 1] - ts[i - 2]);
    return y0 + (t - ts[i - 2]) * slope;
}

__device__ inline double pwl4(const double* ts, const double* ys, double t) {
    int i = 1;
    while (i <= 4 && ts[i - 1] < t) i++;
    if (i <= 4 && ts[i - 1] == t) i += 1;
    if (i <= 1) return ys[0];
    if (i > 4) return ys[3];
    if (ys[i - 2] == ys[i - 1]) return ys[i - 1];
    if (ts[i - 1] == ts[i - 2]) return 0.5 * (ys[i - 2] + ys[i - 1]);
    const double slope = (ys[i - 1] - ys[i - 2]) / (ts[i - 1] - ts[i - 2]);
    return ys[i - 2] + (t - ts[i - 2]) * slope;
}

__device__ inline double wave_tran(const WaveDev& w, double t, const double* params, long long B, long long inst) {
    switch (w.kind) {
        case 0: return pv(w.dc, params, B, inst);
        case 1: return pwl_eval(w.t, w.y, w.npts, t, params, B, inst);
        case 2: {  // src/spectre_env.jl:153-166
            const double v1 = pv(w.v[0], params, B, inst), v2 = pv(w.v[1], params, B, inst);
            const double td = w.v[2].value, tr = w.v[3].value, tf = w.v[4].value, pw = w.v[5].value,
                         per = w.v[6].value;
            const double ts[4] = {td, td + tr, td + tr + pw, td + tr + pw + tf};
            const double ys[4] = {v1, v2, v2, v1};
            const double tt = isinf(per) ? t : fmod(t, per);
            return pwl4(ts, ys, tt);
        }
        case 3: {  // src/spectre_env.jl:169-176
            const double vo = pv(w.v[0], params, B, inst), va = pv(w.v[1], params, B, inst),
                         freq = pv(w.v[2], params, B, inst), td = pv(w.v[3], params, B, inst),
                         theta = pv(w.v[4], params, B, inst), phase = pv(w.v[5], params, B, inst),
                         ncyc = pv(w.v[6], params, B, inst);
            const double d2r = 3.14159265358979323846 / 180.0;
            if (td < t && t < ncyc / freq)
                return vo + va * exp(-(t - td) * theta) * sin((360.0 * freq * (t - td) + phase) * d2r);
            return vo + va * sin(phase * d2r);
        }
    }
    return 0.0;
}

__device__ CB_CTRL_INLINE double wave_value(const WaveDev& w, double t, bool dcop, const double* params, long long B,
                                    long long inst) {
    if (dcop) return w.has_dc ? pv(w.dc, params, B, inst) : wave_tran(w, 0.0, params, B, inst);
    return wave_tran(w, t, params, B, inst);
}

// value of the accepted interpolation polynomial at tt for unknown i (mirrors oracle predict())
__device__ CB_CTRL_INLINE double poly_at(int nh, double tt, double tn, double xn, double h1, double x1, double h2,
                                         double x2) {
    if (nh <= 0) return xn;
    const double a = tt - tn;
    if (nh == 1) return xn + DIV(a * (xn - x1), h1);
    const double d1 = DIV(xn - x1, h1), d2 = DIV(x1 - x2, h2), dd = DIV(d1 - d2, h1 + h2);
    return xn + a * d1 + a * (a + h1) * dd;
}

// ------------------------------------------------------------------------------------------------
// k_control: per-point Newton update + DC / transient state machine, one thread per point.
// Partner of the circuit-specialised generated kernel k_solve (gen_solve_source() in cedarb200.cu),
// which assembles J and r from the device outputs, runs the straight-line static-pivot LU and
// leaves dx = -J^-1 r, the charges q and max|r| in batch-interleaved arrays.  Every access here
// is [k][B] with consecutive threads on consecutive points (fully coalesced).  The control logic
// mirrors the CPU oracle statement for statement.
//
// Newton convergence (transient): weighted update norm  n_k = max_i |dx_i| / (nr_reltol max(|x_i|,|x_i + dx_i|) + atol_i).
// Converged when n_k <= 1.  With cb_options.nr_rate_test, iterations after the first also use the observed contraction
// rho = n_k / n_{k-1} (as Sundials IDA does, the solver behind the reference's tran!): the error left after this
// update is about n_k rho / (1 - rho), and the attempt converges when 3 x that estimate is <= 1.  Level 2 also
// accepts the first update of an attempt when 3 kappa n_1^2 <= 1: a full Newton step converges quadratically, and
// kappa = n_2 / n_1^2 is measured per point whenever a second iteration is taken (start value and floor from the
// strongest nonlinearity a semiconductor device has, exp(v / Vt): e_1 ~ e_0^2 / (2 Vt)).  Both rely on the charges
// handed over by k_lu being those of the updated iterate, q(x) + C dx.  The DC operating
// point keeps the plain test n_k <= 1 together with the residual tolerance of CedarDCOp.
struct CtlArgs {
    double* WV;  // [nwaves][B] source values for the next evaluation
    int mixed;   // 1 = mixed rounds: the kind of a point's next iteration follows its own iteration count (full, then
                 // vcycle - 1 value-only, full, ...); 0 = lock-step rounds: the host's schedule decides (next_vround)
    int next_vround;  // lock-step: the next round is value-only (points that need a fresh Jacobian idle through it)
    int* dc_count;  // number of points still in the DC phase (lock-step: the host schedules full rounds only while > 0)
    int vcycle, pad_;
    int *next_full, *next_any, *next_idle, *next_cnt;   // lists of the NEXT round, built here (counters zeroed by an earlier k_lu)
};
struct CArgs {
    NArgs n;
    const double *DX, *QK, *RMAX, *DVMAX;
    const int* BAD;
    CtlArgs k;
};

__device__ __forceinline__ void store_waves(const NArgs& a, double* WV, long long inst, bool dcop, double t) {
    for (int w = 0; w < a.nwaves; w++) WV[(size_t)w * a.B + inst] = wave_value(a.waves[w], t, dcop, a.params, a.B, inst);
}

#ifndef LU_PTS
#define LU_PTS 32     // points per group (lane = point); 16: two entry-workers per warp, half the shared memory per CTA
#endif
#ifndef LU_W
#define LU_W 16       // entry-workers per CTA
#endif
#ifndef LU_MINB
#define LU_MINB 1
#endif
#ifndef CB_LU_LIN_UNROLL
#define CB_LU_LIN_UNROLL 1
#endif
#ifndef LU_GU
#define LU_GU 8      // independent HBM loads in flight per worker in the gather phases
#endif
typedef unsigned short u16;

// Mapping: a CTA owns CTRL_PTS consecutive points; warp l of the CTA ("lane l" of each point) owns the
// unknowns i = l, l + CTRL_LANES, ...  Every warp access is one contiguous row segment of 32 points.  The
// scalar control logic is replicated across the lanes of a point (same inputs, same result); the two
// cross-lane reductions (Newton convergence AND, LTE max) and the three hand-offs of rows between lanes go
// through shared memory / __syncthreads.  One thread per point (the previous version) left the GPU with
// 16 384 threads -- under one warp per SM sub-partition -- and ran at 5 % of HBM bandwidth.
#ifndef CTRL_PTS
#define CTRL_PTS 32
#endif
// Lanes per point: 8 for large batches (256-thread CTAs, 4 per SM: a 16 384-point launch is one wave); 32 for small ones, where the
// launch is a fraction of a wave and only its latency counts (3 instead of 11 unknowns per thread in the two passes).
// The control step of CTRL_PTS points: thread (pt, lane) = lane `lane` of CTRL_LANES of point `pt`; every thread of the CTA
// must call it (CTA-wide barriers).  Two callers:
//   k_control (FUSED = false): CTA = CTRL_PTS consecutive points, the solver's results come from global memory
//     (DX, QK rows; RMAX, DVMAX, BAD), every point of the batch is visited, idle ones included;
//   k_lu<.., true> (FUSED = true): CTA = one group of list entries straight after their solve, the update vector and the
//     charges are still in the group's shared memory (sv = vals + pt: dx of unknown i at sv[(nnz + cts[i]) * LU_PTS], charge
//     of row i at sv[(nnz + N + i) * LU_PTS]); idle points are visited through the idle list, which this step also builds.
// act = role of the point in THIS round (ACT_*), inb = the thread has a point.  s_red: two rows of CTRL_PTS slots, zeroed
// by the caller behind a barrier (Newton norm and LTE estimate: non-negative doubles, reduced over the lanes by atomicMax
// on their bit patterns).
template <int CTRL_LANES, bool FUSED>
__device__ __forceinline__ void control_points(const NArgs& a, const CtlArgs& c, const long long inst, const bool inb, const int act,
                                               const int pt, const int lane, unsigned long long (*s_red)[CTRL_PTS],
                                               const double* DXg, const double* QKg, const double rmax_in, const double dvmax_in,
                                               const int bad_in, const double* sv, const u16* cts, const int nnz) {
    static_assert(CTRL_PTS == 32, "lane 0 of the points of a CTA must be exactly one warp (list compaction by ballot)");
    const long long B = a.B;
    const bool live = act == ACT_ANY || act == ACT_FULL;   // took part in this round
    const bool pv = act == ACT_ANY;                         // ... with a value-only iteration
    const int N = a.N, NV = a.NV;
    const Opts& o = a.o;
    const long long ii = inb ? inst : 0;   // dead threads keep in-bounds addresses, never dereferenced
    int phase = live ? a.ist[(size_t)IS_PHASE * B + ii] : PH_DONE;
    double* __restrict__ X = a.X + ii;   double* __restrict__ XN = a.XN + ii; double* __restrict__ X1 = a.X1 + ii;
    double* __restrict__ X2 = a.X2 + ii; double* __restrict__ XP = a.XP + ii; double* __restrict__ QN = a.QN + ii;
    double* __restrict__ Q1 = a.Q1 + ii; double* __restrict__ QD = a.QD + ii; double* __restrict__ BETA = a.BETA + ii;
    const double* __restrict__ DX = FUSED ? nullptr : DXg + ii; const double* __restrict__ QK = FUSED ? nullptr : QKg + ii;
    const unsigned char* __restrict__ mask = a.lte_mask;
#define IST(k) a.ist[(size_t)(k) * B + ii]
#define DST(k) a.dst[(size_t)(k) * B + ii]
#define V(arr, i) arr[(size_t)(i) * B]
#define DXV(i) (FUSED ? sv[(size_t)(nnz + cts[i]) * LU_PTS] : V(DX, i))
#define QKV(i) (FUSED ? sv[(size_t)(nnz + N + (i)) * LU_PTS] : V(QK, i))
    int it = 0, stage = 0, nh = 0, bpi = 0, kstep = 0, status = 0, hit_bp = 0, method = 0, np = 0, sidx = 0, nnewton = 0,
        nacc = 0, nrej = 0, retry = 0;
    double t = 0, tnew = 0, h = 0, h1 = 0, h2 = 0, hprop = 0, gshunt = 0, lim = 0, alpha = 0, rmax = 0, dvmax = 0, nrm_prev = 0, kappa = 0;
    int badpt = 0;
    if (live) {
        it = IST(IS_IT); stage = IST(IS_STAGE); nh = IST(IS_NH); bpi = IST(IS_BPI); kstep = IST(IS_KSTEP);
        status = IST(IS_STATUS); hit_bp = IST(IS_HITBP); method = IST(IS_METHOD); np = IST(IS_NP);
        sidx = IST(IS_SIDX); nnewton = IST(IS_NNEWTON); nacc = IST(IS_NACC); nrej = IST(IS_NREJ); retry = IST(IS_RETRY);
        t = DST(DS_T); tnew = DST(DS_TNEW); h = DST(DS_H); h1 = DST(DS_H1); h2 = DST(DS_H2);
        hprop = DST(DS_HPROP); gshunt = DST(DS_GSHUNT); lim = DST(DS_LIM); nrm_prev = DST(DS_NRM); kappa = DST(DS_KAPPA);
        alpha = a.alpha[ii]; rmax = rmax_in; dvmax = dvmax_in; badpt = bad_in;
    }
    const double alpha_old = alpha;
    bool finish = false, begin = false, newton_ok = false, newton_fail = false;
    bool do_accept = false;        // shift history, QD/QN from this iterate
    int copy_mode = 0;             // 1: X <- 0, XN <- 0   2: XN <- X   3: X <- XN   4: QN <- QK, QD <- 0
    int out_from = 0, out_to = 0, nh_out = 0;  // save points [out_from, out_to) are emitted from the NEW history
    bool out_init = false;         // save points at t0 are emitted from XN (operating point), or from X after PH_REINIT
    bool out_from_x = false, reinit_begin = false;

    // ---- pass 1: Newton update of this lane's unknowns, convergence and LTE partials ------------
    const bool solving = live && phase != PH_TRAN_INIT && (!badpt || (badpt & 4));   // (bit 2: re-solved by k_lu's repair pass)
    const bool want_lte = solving && phase == PH_TRAN && !o.fixed_step && np >= 1;
    double sc = 1.0, ratio = 0.0;
    if (solving) {
        lim = o.dv_max;
        sc = dvmax > lim ? DIV(lim, dvmax) : 1.0;
        if (want_lte) {
            if (method == 0) ratio = DIV(h, 2.0 * h + h1);
            else {
                const double pc = DIV(h * (h + h1) * (h + h1 + h2), 6.0);
                const double lc = method == 1 ? DIV(h * h * h, 12.0) : DIV(h * h * (h + h1) * (h + h1), 6.0 * (2.0 * h + h1));
                ratio = np >= 2 ? DIV(lc, lc + pc) : DIV(h, 2.0 * h + h1);
            }
        }
    }
    {
        double nrm = 0.0;
        double err = 0.0;
        if (solving) {
            CB_CTRL_UNROLL
            for (int i = lane; i < N; i += CTRL_LANES) {
                const double dx = sc * DXV(i);
                const double xo = V(X, i), xn = xo + dx;
                const double atol = i < NV ? o.nr_vabstol : o.nr_iabstol;
                nrm = fmax(nrm, DIV(fabs(dx), o.nr_reltol * fmax(fabs(xn), fabs(xo)) + atol));
                V(X, i) = xn;
                if (want_lte && mask[i]) {
                    const double tol = o.reltol * fmax(fabs(xn), fabs(V(XN, i))) + (i < NV ? o.vabstol : o.iabstol);
                    err = fmax(err, DIV(ratio * fabs(xn - V(XP, i)), tol));
                }
            }
        }
        if (nrm > 0.0) atomicMax(&s_red[0][pt], (unsigned long long)__double_as_longlong(nrm));
        if (err > 0.0) atomicMax(&s_red[1][pt], (unsigned long long)__double_as_longlong(err));
    }
    __syncthreads();   // reductions; also publishes the updated X rows to the other lanes of the point
    const double nrm = __longlong_as_double((long long)s_red[0][pt]), err = __longlong_as_double((long long)s_red[1][pt]);
    const int phase_in = phase;

    // ---- scalar control, replicated across the lanes of a point ------------------------------------
    if (live) {
        if (phase == PH_TRAN_INIT) {
            copy_mode = 4;
            if (status != 0) { out_init = true; finish = true; }
            else if (o.reinit) { reinit_begin = true; phase = PH_REINIT; it = 0; }   // t0 samples wait for the re-initialised state
            else { out_init = true; begin = true; phase = PH_TRAN; }
        } else {
            nnewton++;
            if (!pv && lane == 0) a.ist[(size_t)IS_NFULL * B + ii]++;
            if ((badpt & 2) && lane == 0) a.ist[(size_t)IS_NGROWTH * B + ii]++;   // pivot-growth monitor of k_lu
            if (badpt & 4) retry = 1;   // re-solved with partial pivoting by k_lu's repair pass: the iteration stands, but the
                                        // stored factors are the ones that failed -> full iterations for the rest of the attempt
            if (badpt && !(badpt & 4)) {
                newton_fail = true;
                status = 4;
            } else {
                const double restol = phase == PH_DC ? o.dc_abstol : 1e300;
                double est = nrm;   // estimate of the weighted error left after this update
                if (o.rate_test && phase == PH_TRAN) {
                    if (it == 0) {
                        // level 2: the first (full Newton) update of an attempt leaves ~ kappa n_1^2 (quadratic
                        // convergence; kappa learnt from the attempts that did take a second iteration)
                        if (o.rate_test >= 2 && !pv) est = fmin(nrm, 3.0 * kappa * nrm * nrm);
                    } else {
                        if (it == 1 && nrm_prev > 0.0) kappa = fmax(fmax(DIV(nrm, nrm_prev * nrm_prev), 0.7 * kappa), o.kappa_floor);
                        if (nrm < nrm_prev) {
                            // safety 3; a chord update (value-only round) contracts half as fast as the ratio observed
                            // across the preceding Newton update suggests (e_2 ~ 2 (e_1 / e_0) e_1)
                            const double rho = DIV(nrm, nrm_prev);
                            est = nrm * fmin(1.0, DIV((pv ? 6.0 : 3.0) * rho, 1.0 - rho));
                        }
                    }
                }
                const int conv = (est <= 1.0) && (sc == 1.0) && (rmax <= restol);
                nrm_prev = nrm;
                it++;
                if (conv) newton_ok = true;
                else if (it >= (phase == PH_TRAN ? o.max_newton_tran : o.max_newton_dc)) { newton_fail = true; status = 1; }

                if (phase == PH_TRAN && newton_ok) {
                    double fac = 2.0;
                    bool reject = false;
                    if (want_lte) {
                        const int p = (method == 0 || np < 2) ? 1 : 2;
                        fac = err > 0.0 ? 0.9 * pow(err, -1.0 / (p + 1)) : 2.0;
                        fac = fmin(2.0, fmax(0.2, fac));
                        if (err > 1.0) {
                            reject = true;
                            nrej++;
                            hprop = h * fac;
                            if (hprop < o.dt_min) { status = 3; finish = true; }
                            else begin = true;
                        }
                    }
                    if (!reject) {
                        nacc++;
                        do_accept = true;
                        h2 = h1; h1 = h;
                        nh = nh + 1 < 2 ? nh + 1 : 2;
                        nh_out = nh;
                        t = tnew;
                        kstep++;
                        out_from = sidx;
                        while (sidx < o.nsave && a.saveat[sidx] <= t + o.teps) sidx++;
                        out_to = sidx;
                        if (!o.fixed_step) {
                            hprop = h * fac;
                            if (hit_bp) {
                                nh = 0;
                                const double nb = (bpi + 1 < a.nbp) ? a.bp[bpi + 1] - t : o.t1 - t;
                                hprop = fmin(hprop, 0.1 * fmin(h, nb > 0.0 ? nb : h));
                                hprop = fmax(hprop, o.span * 1e-9);
                            }
                        }
                        begin = true;
                    }
                }
            }
            if (phase == PH_REINIT && (newton_ok || newton_fail)) {
                // converged: X is the state at t0 (charges kept, algebraic unknowns at the transient source values);
                // failed: the operating point stands.  Either way the transient starts now.
                copy_mode = newton_ok ? 5 : 3;
                out_init = true; out_from_x = newton_ok;
                status = 0; it = 0;
                begin = true; phase = PH_TRAN;
            } else if (phase == PH_DC && (newton_ok || newton_fail)) {
                bool dc_done = false;
                it = 0;
                if (stage < 0) {
                    if (newton_ok) { dc_done = true; status = 0; }
                    else {
                        copy_mode = 1;
                        stage = 0; gshunt = 1e-2; status = 0;
                        if (o.gmin_steps == 0) gshunt = 0.0;
                    }
                } else if (stage < o.gmin_steps) {
                    copy_mode = newton_ok ? 2 : 3;
                    stage++; gshunt *= 0.1; status = 0;
                    if (stage == o.gmin_steps) gshunt = 0.0;
                } else if (stage == o.gmin_steps) {          // the shunt-free solve that ends the gmin ladder
                    if (newton_ok) { dc_done = true; status = 0; }
                    else if (o.source_steps > 0) {           // source stepping: sources ramped from 0, first stage from x = 0
                        copy_mode = 1;
                        stage++; gshunt = 0.0; status = 0;
                    } else { dc_done = true; status = 2; }
                } else {                                     // source-stepping stage s = stage - gmin_steps of source_steps
                    if (!newton_ok) { dc_done = true; status = 2; }
                    else if (stage - o.gmin_steps >= o.source_steps) {
                        dc_done = true; status = 0;
                        if (lane == 0) a.ist[(size_t)IS_SRCSTEP * B + ii] = 1;
                    } else { copy_mode = 2; stage++; status = 0; }
                }
                if (dc_done) {
                    gshunt = 0.0;
                    if (o.dc_only) {
                        CB_CTRL_UNROLL for (int k = lane; k < a.O; k += CTRL_LANES) a.y_out[(size_t)k * B + inst] = V(X, a.outputs[k]);
                        phase = PH_DONE;
                        if (lane == 0) atomicAdd(a.done_count, 1);
                    } else {
                        copy_mode = 2;
                        phase = PH_TRAN_INIT;
                    }
                }
            } else if (phase == PH_TRAN && newton_fail && o.fixed_step && !retry) {
                copy_mode = 3;
                retry = 1; it = 0; status = 0;
            } else if (phase == PH_TRAN && newton_fail) {
                nrej++;
                if (o.fixed_step) finish = true;
                else {
                    hprop = h / 8.0;
                    if (hprop < o.dt_min) { status = 3; finish = true; }
                    else { status = 0; begin = true; }
                }
            }
        }
    }
    // ---- output sampling (reads rows owned by other lanes: X = x_n, XN = x_{n-1}, X1 = x_{n-2}, not yet shifted)
    if (out_init) {
        const double* __restrict__ src0 = out_from_x ? X : XN;
        while (sidx < o.nsave && a.saveat[sidx] <= o.t0 + o.teps) {
            CB_CTRL_UNROLL for (int k = lane; k < a.O; k += CTRL_LANES) a.y_out[((size_t)k * o.nsave + sidx) * B + inst] = V(src0, a.outputs[k]);
            sidx++;
        }
    }
    for (int sx = out_from; sx < out_to; sx++) {
        const double ts = a.saveat[sx];
        const bool exact = fabs(ts - t) <= o.teps;
        const int ni = o.method == 0 ? 1 : nh_out;   // history order before a breakpoint hit resets nh
        CB_CTRL_UNROLL for (int k = lane; k < a.O; k += CTRL_LANES) {
            const int u = a.outputs[k];
            const double xn = V(X, u);
            a.y_out[((size_t)k * o.nsave + sx) * B + inst] = exact ? xn : poly_at(ni, ts, t, xn, h1, V(XN, u), h2, V(X1, u));
        }
    }
    // ---- next step attempt: scalars first (mirrors the top of the oracle's step loop)
    double a1 = 0.0, a2 = 0.0;
    if (reinit_begin) {   // backward Euler over h0 = span * 1e-12 "from t0 to t0": beta = -q_dc / h0, Newton starts at the operating point
        tnew = o.t0; h = o.span * 1e-12;
        alpha = DIV(1.0, h); a1 = -alpha;
        method = 0; np = 0; retry = 0;
    }
    if (begin) {
        bool more;
        if (o.fixed_step) {
            more = kstep < o.nfixed;
            if (more) { tnew = o.t0 + (double)(kstep + 1) * o.dt; h = tnew - t; hit_bp = 0; }
        } else {
            more = t < o.t1 - o.teps;
            if (more) {
                while (bpi < a.nbp && a.bp[bpi] <= t + o.teps) bpi++;
                const double tb = bpi < a.nbp ? a.bp[bpi] : o.t1;
                h = fmin(hprop, o.dt_max);
                hit_bp = 0;
                if (t + h >= tb - 1e-3 * h) { h = tb - t; tnew = tb; hit_bp = 1; }
                else if (t + 2.0 * h > tb) { h = 0.5 * (tb - t); tnew = t + h; }
                else tnew = t + h;
            }
        }
        if (!more) { finish = true; begin = false; }
        else {
            method = nh == 0 ? 0 : o.method;
            if (method == 0) { alpha = DIV(1.0, h); a1 = -alpha; }
            else if (method == 1) { alpha = DIV(2.0, h); a1 = -alpha; }
            else {
                const double rho = DIV(h, h1);
                alpha = DIV(1.0 + 2.0 * rho, h * (1.0 + rho));
                a1 = -DIV(1.0 + rho, h);
                a2 = DIV(rho * rho, h * (1.0 + rho));
            }
            np = method == 0 ? (nh < 1 ? nh : 1) : nh;
            it = 0;
            retry = 0;
        }
    }
    __syncthreads();   // output sampling has read the un-shifted history rows of other lanes
    // ---- pass 2 over this lane's unknowns: history shift (accept), DC copies, beta + predictor (begin)
    if (do_accept || copy_mode || begin || reinit_begin) {
            CB_CTRL_UNROLL
        for (int i = lane; i < N; i += CTRL_LANES) {
            double x = V(X, i), xn = V(XN, i), x1 = V(X1, i), x2 = V(X2, i), qn = V(QN, i), q1 = V(Q1, i), qd = V(QD, i);
            if (do_accept) {
                const double qk = QKV(i);
                qd = alpha_old * qk + V(BETA, i);
                x2 = x1; x1 = xn; xn = x;
                q1 = qn; qn = qk;
                V(QD, i) = qd; V(X2, i) = x2; V(X1, i) = x1; V(XN, i) = xn; V(Q1, i) = q1; V(QN, i) = qn;
            } else if (copy_mode == 1) { x = 0.0; xn = 0.0; V(X, i) = x; V(XN, i) = xn; }
            else if (copy_mode == 2) { xn = x; V(XN, i) = xn; }
            else if (copy_mode == 3) { x = xn; V(X, i) = x; }
            else if (copy_mode == 4) { qn = QKV(i); qd = 0.0; V(QN, i) = qn; V(QD, i) = qd; }
            else if (copy_mode == 5) { xn = x; qn = QKV(i); qd = 0.0; V(XN, i) = xn; V(QN, i) = qn; V(QD, i) = qd; }
            if (reinit_begin) V(BETA, i) = a1 * qn;
            if (begin) {
                double beta = a1 * qn;
                if (method == 1) beta -= qd;
                else if (method == 2) beta += a2 * q1;
                V(BETA, i) = beta;
                const double xp = poly_at(np, tnew, t, xn, h1, x1, h2, x2);
                V(XP, i) = xp;
                double lm = np >= 1 ? fabs(xn - x1) * DIV(h, h1) : 0.0;
                if (i < NV) lm = fmin(lm, o.dv_max);
                V(X, i) = xn + fmax(-lm, fmin(lm, xp - xn));
            }
        }
    }
    __syncthreads();   // the fill below reads XN rows written by other lanes
    if (finish) {
        for (; sidx < o.nsave; sidx++)
            CB_CTRL_UNROLL for (int k = lane; k < a.O; k += CTRL_LANES)
                a.y_out[((size_t)k * o.nsave + sidx) * B + inst] = V(XN, a.outputs[k]);
        phase = PH_DONE;
        if (lane == 0) atomicAdd(a.done_count, 1);
    }
    if (live) {
        if (lane == 0) {
            IST(IS_PHASE) = phase; IST(IS_IT) = it; IST(IS_STAGE) = stage; IST(IS_NH) = nh; IST(IS_BPI) = bpi;
            IST(IS_KSTEP) = kstep; IST(IS_STATUS) = status; IST(IS_HITBP) = hit_bp; IST(IS_METHOD) = method;
            IST(IS_NP) = np; IST(IS_SIDX) = sidx; IST(IS_NNEWTON) = nnewton; IST(IS_NACC) = nacc; IST(IS_NREJ) = nrej;
            IST(IS_RETRY) = retry;
            DST(DS_T) = t; DST(DS_TNEW) = tnew; DST(DS_H) = h; DST(DS_H1) = h1; DST(DS_H2) = h2;
            DST(DS_HPROP) = hprop; DST(DS_GSHUNT) = gshunt; DST(DS_LIM) = lim; DST(DS_NRM) = nrm_prev; DST(DS_KAPPA) = kappa;
            a.alpha[inst] = (phase == PH_TRAN || phase == PH_REINIT) ? alpha : 0.0;
            if (phase_in == PH_DC && phase != PH_DC) atomicSub(c.dc_count, 1);
        }
        if (phase != PH_DONE) {
            // source stepping of the operating point: every independent source times stage / source_steps
            const double src_scale = (phase == PH_DC && stage > o.gmin_steps) ? DIV((double)(stage - o.gmin_steps), (double)o.source_steps) : 1.0;
            CB_CTRL_UNROLL for (int w = lane; w < a.nwaves; w += CTRL_LANES)
                c.WV[(size_t)w * B + inst] = src_scale * wave_value(a.waves[w], tnew, phase != PH_TRAN && phase != PH_REINIT, a.params, B, inst);
        }
    }
    // ---- role of every point in the NEXT round + device-wide compaction.  Warp 0 of the CTA (lane 0 of every point)
    //      holds the scalar state.  Each CTA reserves one contiguous range of each list with a single atomicAdd: the
    //      lists are ordered by CTA completion, but inside a range the points ascend, so a warp of the consumer kernels
    //      still reads a few contiguous row segments.  Results do not depend on the order (every point is solved on its
    //      own); an idle point (lock-step schedule) is re-examined every round.
    if (lane == 0) {
        int role = ACT_DONE;
        if (inb && (live ? phase != PH_DONE : act == ACT_IDLE)) {
            // a fresh Jacobian for DC iterations, the first iteration of a step attempt and every vcycle-th iteration
            // (the fixed-step retry of a failed attempt takes full iterations only: chord iterations are what failed)
            const bool want_full = !live ? true : c.mixed ? (phase != PH_TRAN || retry || it % c.vcycle == 0)
                                                          : (phase != PH_TRAN || retry || it == 0);
            role = c.mixed ? (want_full ? ACT_FULL : ACT_ANY)
                           : (c.next_vround ? (want_full ? ACT_IDLE : ACT_ANY) : ACT_FULL);
        }
        if (inb && (live || act == ACT_IDLE)) a.active[inst] = role;
        const unsigned bf = __ballot_sync(0xffffffffu, role == ACT_FULL), ba = __ballot_sync(0xffffffffu, role == ACT_ANY);
        const unsigned bi = FUSED ? __ballot_sync(0xffffffffu, role == ACT_IDLE) : 0u;
        int basef = 0, basea = 0, basei = 0;
        if (FUSED && !c.next_cnt) return;   // blocked mode: a.active[] is all the next round needs
        if (pt == 0) {
            if (bf) basef = atomicAdd(c.next_cnt + 0, __popc(bf));
            if (ba) basea = atomicAdd(c.next_cnt + 1, __popc(ba));
            if (bi) basei = atomicAdd(c.next_cnt + 2, __popc(bi));
        }
        basef = __shfl_sync(0xffffffffu, basef, 0);
        basea = __shfl_sync(0xffffffffu, basea, 0);
        if (FUSED) basei = __shfl_sync(0xffffffffu, basei, 0);
        const unsigned below = (1u << pt) - 1u;
        if (role == ACT_FULL) c.next_full[basef + __popc(bf & below)] = (int)inst;
        if (role == ACT_ANY) c.next_any[basea + __popc(ba & below)] = (int)inst;
        if (FUSED && role == ACT_IDLE) c.next_idle[basei + __popc(bi & below)] = (int)inst;
    }
#undef IST
#undef DST
#undef V
#undef DXV
#undef QKV
}

#ifndef CTRL_MINB
#define CTRL_MINB 4   // resident CTAs per SM of the 8-lane variant (64 registers per thread)
#endif
template <int CTRL_LANES>
__global__ void __launch_bounds__(CTRL_PTS * CTRL_LANES, CTRL_LANES <= 8 ? CTRL_MINB : 1) k_control(const CArgs c) {
    __shared__ unsigned long long s_red[2][CTRL_PTS];
    const int pt = threadIdx.x % CTRL_PTS, lane = threadIdx.x / CTRL_PTS;
    const long long inst = (long long)blockIdx.x * CTRL_PTS + pt;
    const bool inb = inst < c.n.B;
    const int act = inb ? c.n.active[inst] : ACT_DONE;
    if (lane == 0) { s_red[0][pt] = 0ull; s_red[1][pt] = 0ull; }
    const long long ii = inb ? inst : 0;
    const bool live = act == ACT_ANY || act == ACT_FULL;
    const double rmax = live ? c.RMAX[ii] : 0.0, dvmax = live ? c.DVMAX[ii] : 0.0;
    const int bad = live ? c.BAD[ii] : 0;
    __syncthreads();
    control_points<CTRL_LANES, false>(c.n, c.k, inst, inb, act, pt, lane, s_red, c.DX, c.QK, rmax, dvmax, bad, nullptr, nullptr, 0);
}


// ------------------------------------------------------------------------------------------------
// k_lu: assembly + batched static-pivot sparse LU + triangular solves, factors staged in shared memory.
//
// A CTA owns LU_PTS = 32 consecutive sweep points (lane = point, so every global access is one contiguous
// 256-byte row segment and there is no divergence) and keeps their matrices in shared memory as
// vals[entry][lane]: nnz(L+U) entries, the right-hand side in elimination-step order, the charges.  Its LU_W warps
// split the work of each phase by *entry*, not by point:
//   1. assembly: warp w gathers the matrix entries / residual rows w, w + LU_W, ... from the batch-
//      interleaved device outputs (each value is read from HBM exactly once);
//   2. elimination, level by level: pivots whose rows and columns receive no update from one another are
//      one level (the internal nodes of all transistors, for instance).  Per level the warps first invert
//      the level's pivots, then apply its update ops  vals[dst] -= (vals[l] * inv) * vals[u];  ops that
//      hit the same destination are kept on one warp in ascending pivot order, so the factors are
//      bit-identical to a sequential right-looking sweep.  The forward substitution is the same op stream
//      with the right-hand side as an extra column;
//   3. backward substitution, also level-scheduled on the U-row dependency DAG;
//   4. dx, max|dv|, max|r| and the singular / non-finite flag.
// The op lists are warp-uniform int4 streams read through L1; the code is a handful of small loops, so it
// stays in the instruction cache (the generated straight-line k_solve it replaces was 13.6k SASS
// instructions and ran at the cold instruction-fetch rate, ~29 cycles per instruction).
// Index tables of k_lu: every small table (level schedules, pivot / row lists, the U pattern, the linear-part maps) is
// packed as 16-bit entries into ONE blob (ops as four 16-bit positions = 8 bytes).  When the blob fits behind the matrices
// of the group (DFF: 207 KB of matrices + ~12 KB of tables of the 227 KB a CTA may have) every CTA copies it into shared
// memory once, so the ~60 barrier-separated level phases of a group read their schedules at shared-memory latency
// instead of L2 latency (L1 is ~16 KB at this carve-out); otherwise the same code reads the blob from global memory.
// The members are byte offsets into the blob.
struct LuTabs {
    int op_ptr, piv_ptr, piv, ops, brow_ptr, brow, u_ptr, u_pos, u_col, diag_pos, sop_ptr, sops;
    int a_lin;                // per LU entry: bit 15 = node diagonal (gshunt), low bits = index of its linear stamp + 1 (0: none)
    int rl_ptr, rl_lin, rl_col, rs_ptr, rs_wave, row_to_step, col_to_step;
    int item_ptr, sitem_ptr, citem_ptr;
    int mtab;                 // the distinct multipliers of the gather items (doubles)
    int bytes;                // size of the blob (multiple of 16)
};

struct LArgs {
    NArgs n;
    const unsigned char* tab; // the blob in global memory
    LuTabs t;
    // full rounds: ops (l, u, dst, pivot diag) positions into vals, [nlev * LU_W + 1] op_ptr / piv_ptr; backward
    // substitution rows brow by level, [nblev * LU_W + 1] brow_ptr; value-only rounds: sops (l, rhs source, rhs dst,
    // pivot diag), level-scheduled on the L dependency DAG, [nslev * LU_W + 1] sop_ptr
    int nlev, nblev, nslev, pad1_;
    // gather items: x = dev_out row | index into mtab << 20, y = dst position (| second position << 16 for citems);
    // [LU_W + 1] item_ptr per kind in the blob
    const int2* items;
    const int2* sitems;       // gather items of the residual and charge rows only (value-only rounds)
    Lists cur;                // this round's point lists: groups of LU_PTS full-iteration points, then of value-only points
    int* zero_cnt;            // counters of the NEXT round's lists (k_control of this round fills them): zeroed here
    // first-order charge update  q(x + dx) ~ q(x) + C dx:  items (dev_out row of dQ/dV, vals slot of q_row, vals slot of
    // dx_col), grouped by destination row like the gather items
    const int2* citems;
    double* LUF;              // [nnz_lu][B] factors of the last full round: L (unscaled), U, inverted pivots
    const double* WV;
    double *DX, *QK, *RMAX, *DVMAX;
    int* BAD;
    double growth_max;        // pivot-growth bound: an elimination multiplier above it flags the point (BAD bit 1)
    CtlArgs k;                // the control step, when it is fused into this kernel (k_lu<.., true>)
    double* pp_scratch;       // [gridDim.x][N (N + 1)] dense system of the point being re-solved with partial pivoting, or null
    int blocked, pad3_;       // fused kernel only: 1 = no point lists at all, group g = sweep points [32 g, 32 g + 32), see k_lu
    const u16 *e_row, *e_col; // per LU entry: elimination step of its row / of its column (repair pass only; global memory)
};

// The slow path behind the pivot-growth monitor (hard part 3): ONE point's N x N system, dense in
// global scratch as A[N][N + 1] (last column = right-hand side, everything in elimination-step coordinates), solved by the
// whole CTA with PARTIAL pivoting -- row search, swap, rank-1 update per column with three barriers each, then a
// column-oriented backward substitution.  ~2 N barriers and N^3 / 3 multiply-adds: ~0.1 ms for the DFF's 85 unknowns,
// paid only by points whose static pivot order failed numerically.  Returns 0, or 1 when a column has no usable pivot.
// The solution is left in the last column.
__device__ __noinline__ int lu_dense_pp(double* __restrict__ A, const int N, int* s_i, double* s_v) {
    const int T = LU_PTS * LU_W, tid = threadIdx.x, ld = N + 1;
    for (int k = 0; k < N; k++) {
        double best = -1.0;
        int bi = k;
        for (int i = k + tid; i < N; i += T) {
            const double v = fabs(A[(size_t)i * ld + k]);
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best) { best = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { s_v[tid >> 5] = best; s_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int q = 1; q < LU_W; q++) if (s_v[q] > s_v[0]) { s_v[0] = s_v[q]; s_i[0] = s_i[q]; }
        }
        __syncthreads();
        const int p = s_i[0];
        const double pv = s_v[0];
        __syncthreads();   // s_v / s_i are rewritten by the next column's search
        if (!(pv > 0.0) || !isfinite(pv)) return 1;
        if (p != k) {
            for (int j = tid; j <= N; j += T) {
                const double t = A[(size_t)k * ld + j];
                A[(size_t)k * ld + j] = A[(size_t)p * ld + j];
                A[(size_t)p * ld + j] = t;
            }
            __syncthreads();
        }
        const double inv = 1.0 / A[(size_t)k * ld + k];
        const int nr = N - k - 1, nc = N - k;   // rows k + 1 .. N - 1, columns k + 1 .. N
        for (int idx = tid; idx < nr * nc; idx += T) {
            const int i = k + 1 + idx / nc, j = k + 1 + idx % nc;
            A[(size_t)i * ld + j] -= (A[(size_t)i * ld + k] * inv) * A[(size_t)k * ld + j];
        }
        __syncthreads();
    }
    for (int k = N - 1; k >= 0; k--) {
        if (tid == 0) A[(size_t)k * ld + N] /= A[(size_t)k * ld + k];
        __syncthreads();
        const double xk = A[(size_t)k * ld + N];
        for (int i = tid; i < k; i += T) A[(size_t)i * ld + N] -= A[(size_t)i * ld + k] * xk;
        __syncthreads();
    }
    return 0;
}

// One group of LU_PTS points (lane = point) of one kind: SOLVE = value-only iteration with the stored factors.
// Full groups may take a second pass ("repair", when LArgs.pp_scratch is set): the points that the first pass flagged
// (zero / non-finite pivot, elimination multiplier above growth_max, non-finite update) are assembled again and solved
// one by one with lu_dense_pp; their BAD word becomes 4 | 2 ("solved by the slow path": the control step carries on
// with the iteration instead of rejecting it, and keeps the point on full iterations for the rest of the attempt,
// because its stored factors are the ones that failed).
// tb = base of the table blob (shared or global memory).
template <bool SOLVE, bool FUSED, bool REPAIR, bool CANREPAIR>
__device__ __forceinline__ int lu_group(const LArgs& c, double* vals_, const unsigned char* __restrict__ tb,
                                         unsigned long long (*s_red)[LU_PTS], int* s_bad,
                                         const long long inst, const bool on, const int lane, const int w,
                                         int* s_redo, int* s_pi, double* s_pv) {
#define T16(name) ((const u16*)(tb + c.t.name))
    const NArgs& a = c.n;
    // REPAIR: the second pass over a full group (see above), compiled as a function of its own (lu_repair) so that the
    // code of the normal pass is what it would be without it
    const bool redo = REPAIR && s_redo[lane] != 0;   // this lane's point is being re-solved
    const bool st = on && (!REPAIR || redo);         // this thread stores results of its lane's point
    if (REPAIR) __syncthreads();                     // s_redo is read by everybody before anything rewrites shared memory
    // per-point reductions over the workers: max |r|, max |dv| (non-negative, never NaN: fmax drops NaNs -> the bit
    // patterns order like the values) and the singular / non-finite / growth flags, by shared-memory atomics
    if (w == 0) { s_red[0][lane] = 0ull; s_red[1][lane] = 0ull; s_bad[lane] = 0; }
    const long long B = a.B;
    const int N = a.N, NV = a.NV, nnz = a.nnz_lu;
    double* __restrict__ vals = vals_ + lane;
#define VL(i) vals[(size_t)(i) * LU_PTS]
    const double alpha = a.alpha[inst], gshunt = a.dst[(size_t)DS_GSHUNT * B + inst];
    const double* __restrict__ od = a.dev_out + inst;
    const double* __restrict__ X = a.X + inst;
    const u16* __restrict__ row_to_step = T16(row_to_step);
    const u16* __restrict__ col_to_step = T16(col_to_step);
    const u16* __restrict__ rl_ptr = T16(rl_ptr);
    const u16* __restrict__ rl_lin = T16(rl_lin);
    const u16* __restrict__ rl_col = T16(rl_col);
    const double* __restrict__ mtab = (const double*)(tb + c.t.mtab);
    // ---- 1a. linear part (cached / uniform loads only): A = G_lin + alpha C_lin (+ gshunt on node diagonals),
    //          F = -(f_lin + beta) in step order, Q = q_lin in row order
    double* __restrict__ Fv = vals + (size_t)nnz * LU_PTS;
    double* __restrict__ Qv = vals + (size_t)(nnz + N) * LU_PTS;
    if (SOLVE) {
        const double* __restrict__ lf = c.LUF + inst;
#pragma unroll 8
        for (int e = w; e < nnz; e += LU_W) VL(e) = lf[(size_t)e * B];
    } else {
        const u16* __restrict__ alin = T16(a_lin);
#if CB_LU_LIN_UNROLL
#pragma unroll 4
#endif
        for (int e = w; e < nnz; e += LU_W) {
            double v = 0.0;
            const unsigned al = alin[e];
            const int lin = (int)(al & 0x7fffu) - 1;
            if (lin >= 0) {
                const size_t li = (size_t)lin * a.lin_ent_stride + (size_t)inst * a.lin_inst_stride;
                v = a.lin_g[li] + alpha * a.lin_c[li];
            }
            if (al & 0x8000u) v += gshunt;
            VL(e) = v;
        }
    }
    {
        const u16* __restrict__ rs_ptr = T16(rs_ptr);
        const u16* __restrict__ rs_wave = T16(rs_wave);
        for (int i = w; i < N; i += LU_W) {
            double f = 0.0, q = 0.0;
            for (int p = rl_ptr[i]; p < rl_ptr[i + 1]; p++) {
                const size_t li = (size_t)rl_lin[p] * a.lin_ent_stride + (size_t)inst * a.lin_inst_stride;
                const double xc = X[(size_t)rl_col[p] * B];
                f += a.lin_g[li] * xc;
                q += a.lin_c[li] * xc;
            }
            for (int p = rs_ptr[i]; p < rs_ptr[i + 1]; p++) f += a.rs_coef[p] * c.WV[(size_t)rs_wave[p] * B + inst];
            if (i < NV) f += gshunt * X[(size_t)i * B];
            Fv[(size_t)row_to_step[i] * LU_PTS] = -(f + a.BETA[(size_t)i * B + inst]);
            Qv[(size_t)i * LU_PTS] = q;
        }
    }
    __syncthreads();
    // ---- 1b. device outputs: one flat stream of (src row, dst, mult) items, LU_GU independent HBM loads in
    //          flight per warp (the items of the next batch are fetched while this batch's values are on their way);
    //          all items of one destination are on one warp
    {
        const int2* __restrict__ items = SOLVE ? c.sitems : c.items;
        const u16* __restrict__ iptr = SOLVE ? T16(sitem_ptr) : T16(item_ptr);
        int q0 = iptr[w];
        const int q1 = iptr[w + 1];
        int2 it[LU_GU];
        if (q0 < q1) {
#pragma unroll
            for (int u = 0; u < LU_GU; u++) it[u] = __ldg(items + min(q0 + u, q1 - 1));
        }
        for (; q0 < q1; q0 += LU_GU) {
            double v[LU_GU];
            int2 nx[LU_GU];
#pragma unroll
            for (int u = 0; u < LU_GU; u++) v[u] = __ldg(od + (size_t)(it[u].x & 0xfffff) * B);
#pragma unroll
            for (int u = 0; u < LU_GU; u++) nx[u] = __ldg(items + min(q0 + LU_GU + u, q1 - 1));
#pragma unroll
            for (int u = 0; u < LU_GU; u++)
                if (q0 + u < q1) VL(it[u].y) += mtab[(unsigned)it[u].x >> 20] * v[u];
#pragma unroll
            for (int u = 0; u < LU_GU; u++) it[u] = nx[u];
        }
    }
    __syncthreads();
    // ---- 1c. residual rows: b = -(f + alpha q + beta), charges out
    double rmax = 0.0;
    for (int i = w; i < N; i += LU_W) {
        const double q = Qv[(size_t)i * LU_PTS];
        double* bp = Fv + (size_t)row_to_step[i] * LU_PTS;
        const double b = *bp - alpha * q;
        *bp = b;
        rmax = fmax(rmax, fabs(b));
    }
    atomicMax(&s_red[0][lane], (unsigned long long)__double_as_longlong(rmax));
    int bad = 0;
    __syncthreads();
    if (REPAIR) {
        // ---- 2'. repair pass: the flagged points one by one, dense with partial pivoting (lu_dense_pp)
        double* __restrict__ A = c.pp_scratch + (size_t)blockIdx.x * N * (N + 1);
        const u16* __restrict__ e_row = c.e_row;
        const u16* __restrict__ e_col = c.e_col;
        for (int l = 0; l < LU_PTS; l++) {
            if (!s_redo[l]) continue;   // block-uniform
            for (int k = threadIdx.x; k < N * (N + 1); k += LU_PTS * LU_W) A[k] = 0.0;
            __syncthreads();
            for (int e = threadIdx.x; e < nnz; e += LU_PTS * LU_W) A[(size_t)e_row[e] * (N + 1) + e_col[e]] = vals_[(size_t)e * LU_PTS + l];
            for (int r = threadIdx.x; r < N; r += LU_PTS * LU_W) A[(size_t)r * (N + 1) + N] = vals_[(size_t)(nnz + r) * LU_PTS + l];
            __syncthreads();
            const int sing = lu_dense_pp(A, N, s_pi, s_pv);
            for (int r = threadIdx.x; r < N; r += LU_PTS * LU_W) vals_[(size_t)(nnz + r) * LU_PTS + l] = sing ? 0.0 : A[(size_t)r * (N + 1) + N];
            if (threadIdx.x == 0) s_bad[l] = sing ? 3 : 6;   // 6 = growth seen (2) + solved by the slow path (4)
            __syncthreads();
        }
    } else {
    // ---- 2. elimination by levels (fused forward substitution) ------------------------------------------
    {
        const int nlev = SOLVE ? c.nslev : c.nlev;
        const ushort4* __restrict__ ops = (const ushort4*)(tb + (SOLVE ? c.t.sops : c.t.ops));
        const u16* __restrict__ op_ptr = SOLVE ? T16(sop_ptr) : T16(op_ptr);
        const u16* __restrict__ piv_ptr = T16(piv_ptr);
        const u16* __restrict__ piv = T16(piv);
        for (int lv = 0; lv < nlev; lv++) {
            const int slot = lv * LU_W + w;
            if (!SOLVE) {
                for (int p = piv_ptr[slot]; p < piv_ptr[slot + 1]; p++) {
                    const int dp = piv[p];
                    const double d = VL(dp);
                    bad |= !(fabs(d) > 0.0);
                    VL(dp) = 1.0 / d;
                }
                __syncthreads();
            }
            int q0 = op_ptr[slot];
            const int q1 = op_ptr[slot + 1];
            for (; q0 < q1; q0++) {
                const ushort4 cur = ops[q0];
                const double l = VL(cur.x) * VL(cur.w);
                if (!SOLVE) bad |= (fabs(l) > c.growth_max) << 1;   // pivot-growth monitor (BAD bit 1)
                VL(cur.z) -= l * VL(cur.y);
            }
            __syncthreads();
        }
    }
    if (!SOLVE && c.LUF) {   // keep the factors for the value-only rounds that follow
        double* __restrict__ lf = c.LUF + inst;
        if (on)
#pragma unroll 8
            for (int e = w; e < nnz; e += LU_W) lf[(size_t)e * B] = VL(e);
    }
    // ---- 3. backward substitution by levels: x_k = (b_k - sum_j u_kj x_j) * inv_k, in place in the rhs slots
    {
        const u16* __restrict__ brow_ptr = T16(brow_ptr);
        const u16* __restrict__ brow = T16(brow);
        const u16* __restrict__ u_ptr = T16(u_ptr);
        const u16* __restrict__ u_pos = T16(u_pos);
        const u16* __restrict__ u_col = T16(u_col);
        const u16* __restrict__ diag_pos = T16(diag_pos);
        for (int lv = 0; lv < c.nblev; lv++) {
            const int slot = lv * LU_W + w;
            for (int p = brow_ptr[slot]; p < brow_ptr[slot + 1]; p++) {
                const int k = brow[p];
                double acc = VL(nnz + k);
                for (int u = u_ptr[k]; u < u_ptr[k + 1]; u++) acc -= VL(u_pos[u]) * VL(nnz + u_col[u]);
                VL(nnz + k) = acc * VL(diag_pos[k]);
            }
            __syncthreads();
        }
    }
    }   // !REPAIR
    // ---- 3b. charges of the updated iterate to first order: q(x + dx) ~ q(x) + C dx (linear capacitors exactly).
    //          The accepted step keeps these charges, so they must belong to the iterate that is accepted, not to
    //          the point of the last device evaluation (which lies |dx| away).  In value-only rounds dQ/dV is that
    //          of the last full round.
    //          Only the rate-based acceptance tests need it: the plain test accepts when |dx| is below the Newton
    //          tolerance, where the update is negligible.
    if (a.o.rate_test) {
        for (int i = w; i < N; i += LU_W) {
            double dq = 0.0;
            for (int p = rl_ptr[i]; p < rl_ptr[i + 1]; p++) {
                const size_t li = (size_t)rl_lin[p] * a.lin_ent_stride + (size_t)inst * a.lin_inst_stride;
                dq += a.lin_c[li] * VL(nnz + col_to_step[rl_col[p]]);
            }
            Qv[(size_t)i * LU_PTS] += dq;
        }
        __syncthreads();
        {
            const u16* __restrict__ iptr = T16(citem_ptr);
            int q0 = iptr[w];
            const int q1 = iptr[w + 1];
            int2 it[LU_GU];
            if (q0 < q1) {
#pragma unroll
                for (int u = 0; u < LU_GU; u++) it[u] = __ldg(c.citems + min(q0 + u, q1 - 1));
            }
            for (; q0 < q1; q0 += LU_GU) {
                double v[LU_GU];
                int2 nx[LU_GU];
#pragma unroll
                for (int u = 0; u < LU_GU; u++) v[u] = __ldg(od + (size_t)(it[u].x & 0xfffff) * B);
#pragma unroll
                for (int u = 0; u < LU_GU; u++) nx[u] = __ldg(c.citems + min(q0 + LU_GU + u, q1 - 1));
#pragma unroll
                for (int u = 0; u < LU_GU; u++)
                    if (q0 + u < q1) VL(it[u].y & 0xffff) += mtab[(unsigned)it[u].x >> 20] * v[u] * VL((unsigned)it[u].y >> 16);
#pragma unroll
                for (int u = 0; u < LU_GU; u++) it[u] = nx[u];
            }
        }
        __syncthreads();
    }
    if (st && !FUSED)
        for (int i = w; i < N; i += LU_W) c.QK[(size_t)i * B + inst] = Qv[(size_t)i * LU_PTS];
    // ---- 4. update vector and norms ---------------------------------------------------------------------
    double dvm = 0.0;
    for (int i = w; i < N; i += LU_W) {
        const double dx = VL(nnz + col_to_step[i]);
        if (st && !FUSED) c.DX[(size_t)i * B + inst] = dx;
        if (i < NV) dvm = fmax(dvm, fabs(dx));
        bad |= !isfinite(dx);
    }
    atomicMax(&s_red[1][lane], (unsigned long long)__double_as_longlong(dvm));
    if (bad) atomicOr(&s_bad[lane], bad);
    __syncthreads();
    if (!FUSED) {
        if (w == 0 && st) {
            c.RMAX[inst] = __longlong_as_double((long long)s_red[0][lane]);
            c.DVMAX[inst] = __longlong_as_double((long long)s_red[1][lane]);
            c.BAD[inst] = s_bad[lane];
        }
        if (CANREPAIR && !SOLVE && !REPAIR) {
            // a second pass for the points this pass flagged?  (block-uniform answer; the barrier is also the one after
            // which the next group may reuse the shared-memory matrix and the reduction slots)
            const int flagged = on ? s_bad[lane] : 0;
            if (w == 0) s_redo[lane] = flagged != 0;
            return __syncthreads_or(flagged != 0);
        }
        __syncthreads();   // the next group reuses the shared-memory matrix and the reduction slots
    }
    return 0;
#undef VL
#undef T16
}

// The repair pass as a function of its own (see lu_group<.., REPAIR = true>)
__device__ __noinline__ void lu_repair(const LArgs& c, double* vals_, const unsigned char* tb, unsigned long long (*s_red)[LU_PTS],
                                       int* s_bad, const long long inst, const bool on, const int lane, const int w, int* s_redo,
                                       int* s_pi, double* s_pv) {
    lu_group<false, false, true, true>(c, vals_, tb, s_red, s_bad, inst, on, lane, w, s_redo, s_pi, s_pv);
}

// A CTA works through groups g = blockIdx.x, blockIdx.x + gridDim.x, ... of this round's lists: first the groups of
// full-iteration points (assembly + LU + solves, factors stored), then the groups of value-only points (solves with the
// stored factors).  The lists are dense (device-wide compaction by the control step), so every group but the last of each
// kind is full whatever fraction of the sweep points takes part in the round.
// STAGED: the table blob is copied behind the matrices in shared memory (launch with lu_smem + t.bytes dynamic bytes).
// FUSED: the control step of the group's points (control_points: Newton update, convergence, step control, output
//   sampling, next round's lists) runs as the tail of their solve, with the update vector and the charges still in
//   shared memory -- no k_control launch, no DX / QK / RMAX / DVMAX / BAD round trip through HBM.  Idle points of the
//   lock-step schedule are visited through a third list.  The counters zeroed here are those of the round after next
//   (three rotating list buffers), because this kernel itself fills the next round's.
//   Parity-green and 2.7x SLOWER (profiles/probe_r2o.log; ncu profiles/ncu_lu_fused_r2v.json): the groups then append
//   their points to the next lists in GROUP order, and after a few rounds of points moving between the full / value-only /
//   idle lists a group's 32 entries come from many different 32-point blocks -- every row access of k_lu and of the eval
//   kernels (base + inst * 8 bytes per lane) touches 17 sectors per request instead of 4.4.  The stand-alone k_control
//   walks the points in instance order every round and so re-establishes runs of consecutive points; that ordered
//   compaction is what fusion gives up.  Off by default (CB_FUSE=1).
// REPAIRABLE: full groups may call the repair pass (cb_options.pivot_repair); the kernel without it is exactly the
// kernel there was before the repair pass existed (same-box A/B: the repairable one is 1 % / 2 % slower at 16 384 / 2 048
// points, profiles/probe_r2ag.log, although the pass itself never runs on the bench workload).
template <bool STAGED, bool FUSED, bool REPAIRABLE>
__global__ void __launch_bounds__(LU_PTS * LU_W, LU_MINB) k_lu(const __grid_constant__ LArgs c) {
    extern __shared__ __align__(16) double vals_[];
    __shared__ unsigned long long s_red[2][LU_PTS];
    __shared__ unsigned long long s_ctl[2][LU_PTS];
    __shared__ int s_bad[LU_PTS];
    __shared__ int s_redo[LU_PTS];   // repair pass of lu_group: lanes to re-solve, pivot search scratch
    __shared__ int s_pi[LU_W];
    __shared__ double s_pv[LU_W];
    const int lane = threadIdx.x % LU_PTS, w = threadIdx.x / LU_PTS;
    if (blockIdx.x == 0 && threadIdx.x < 3) c.zero_cnt[threadIdx.x] = 0;
    const int nf = c.cur.cnt[0], na = c.cur.cnt[1], ni = FUSED ? c.cur.cnt[2] : 0;
    const int gf = (nf + LU_PTS - 1) / LU_PTS, ga = (na + LU_PTS - 1) / LU_PTS, gi = (ni + LU_PTS - 1) / LU_PTS;
    if (!(FUSED && c.blocked) && (int)blockIdx.x >= gf + ga + gi) return;
    const unsigned char* tb = c.tab;
    if (STAGED) {
        unsigned char* st = (unsigned char*)(vals_ + (size_t)(c.n.nnz_lu + 2 * c.n.N) * LU_PTS);
        const int4* __restrict__ src = (const int4*)c.tab;
        for (int k = threadIdx.x; k < c.t.bytes / 16; k += LU_PTS * LU_W) ((int4*)st)[k] = __ldg(src + k);
        tb = st;
        __syncthreads();
    }
    if (FUSED && c.blocked) {
        // Blocked mode (small batches, where a round is latency-bound and the launch of a separate control kernel is a
        // fifth of it): no point lists.  Group g IS the block of sweep points [32 g, 32 g + 32): every access is a run of
        // consecutive points whatever the participation (which is what the list-based fused mode loses), lanes whose
        // point takes no part in this kind of iteration shadow along without storing, and the control step of the block
        // runs as the tail of its solve.  Costs a value-only round the groups it could have skipped -- nothing while the
        // whole batch is under one wave of CTAs.
        const long long Bn = c.n.B;
        const int nblk = (int)((Bn + LU_PTS - 1) / LU_PTS);
        for (int g = blockIdx.x; g < nblk; g += gridDim.x) {
            const long long k = (long long)g * LU_PTS + lane;
            const bool inb = k < Bn;
            const long long inst = inb ? k : Bn - 1;
            const int act = inb ? c.n.active[inst] : ACT_DONE;
            const int any_f = __syncthreads_or(act == ACT_FULL), any_a = __syncthreads_or(act == ACT_ANY);
            const int any_i = __syncthreads_or(act == ACT_IDLE);
            if (!(any_f | any_a | any_i)) continue;
            for (int kind = 0; kind < 3; kind++) {
                if (kind == 0 ? !any_f : (kind == 1 ? !any_a : (any_f | any_a) != 0)) continue;   // kind 2: a block of idle points only
                const bool last = kind == 2 || (kind == 1) || !any_a;
                const bool on = kind < 2 && act == (kind == 0 ? ACT_FULL : ACT_ANY);
                if (w == 0) { s_ctl[0][lane] = 0ull; s_ctl[1][lane] = 0ull; }
                if (kind == 0) lu_group<false, true, false, false>(c, vals_, tb, s_red, s_bad, inst, on, lane, w, s_redo, s_pi, s_pv);
                else if (kind == 1) lu_group<true, true, false, false>(c, vals_, tb, s_red, s_bad, inst, on, lane, w, s_redo, s_pi, s_pv);
                else __syncthreads();
                const int act_eff = on ? act : ((last && act == ACT_IDLE) ? ACT_IDLE : ACT_DONE);
                const double rmax = on ? __longlong_as_double((long long)s_red[0][lane]) : 0.0;
                const double dvmax = on ? __longlong_as_double((long long)s_red[1][lane]) : 0.0;
                const int bad = on ? s_bad[lane] : 0;
                control_points<LU_W, true>(c.n, c.k, inst, act_eff != ACT_DONE, act_eff, lane, w, s_ctl, nullptr, nullptr, rmax, dvmax, bad,
                                           vals_ + lane, (const u16*)(tb + c.t.col_to_step), c.n.nnz_lu);
                __syncthreads();
            }
        }
        return;
    }
    for (int g = blockIdx.x; g < gf + ga + gi; g += gridDim.x) {
        const int kind = g < gf ? 0 : (g < gf + ga ? 1 : 2);
        const int g0 = (kind == 0 ? g : (kind == 1 ? g - gf : g - gf - ga)) * LU_PTS, n = kind == 0 ? nf : (kind == 1 ? na : ni);
        const int* __restrict__ list = kind == 0 ? c.cur.full : (kind == 1 ? c.cur.any : c.cur.idle);
        const bool on = g0 + lane < n;
        const long long inst = list[on ? g0 + lane : g0];   // idle lanes shadow the group's first point, never store
        if (FUSED && w == 0) { s_ctl[0][lane] = 0ull; s_ctl[1][lane] = 0ull; }
        if (kind == 1) lu_group<true, FUSED, false, false>(c, vals_, tb, s_red, s_bad, inst, on, lane, w, s_redo, s_pi, s_pv);
        else if (kind == 0) {
            const int flagged = lu_group<false, FUSED, false, REPAIRABLE && !FUSED>(c, vals_, tb, s_red, s_bad, inst, on, lane, w, s_redo, s_pi, s_pv);
            if (REPAIRABLE && !FUSED && flagged && c.pp_scratch) lu_repair(c, vals_, tb, s_red, s_bad, inst, on, lane, w, s_redo, s_pi, s_pv);
        }
        else if (FUSED) __syncthreads();   // publishes the zeroed reduction slots
        if (FUSED) {
            const double rmax = kind < 2 ? __longlong_as_double((long long)s_red[0][lane]) : 0.0;
            const double dvmax = kind < 2 ? __longlong_as_double((long long)s_red[1][lane]) : 0.0;
            const int bad = kind < 2 ? s_bad[lane] : 0;
            const int act = on ? (kind == 0 ? ACT_FULL : (kind == 1 ? ACT_ANY : ACT_IDLE)) : ACT_DONE;
            control_points<LU_W, true>(c.n, c.k, inst, on, act, lane, w, s_ctl, nullptr, nullptr, rmax, dvmax, bad, vals_ + lane,
                                       (const u16*)(tb + c.t.col_to_step), c.n.nnz_lu);
            __syncthreads();   // the next group reuses the shared-memory matrix and the reduction slots
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Small-signal analyses (cb_ac / cb_noise; reference ac!/noise!, src/ac.jl:75-190, :257-284).
//
// After the DC operating point one more device evaluation with alpha = 0 leaves G = dI/dV in the J rows of dev_out
// and C = dQ/dV in the rows next to them.  k_ac then factors the complex matrix A = G + j w C of every
// (sweep point, frequency) pair with the SAME static pivot order and fill pattern as the Newton matrix:
//   CTA = AC_PTS consecutive sweep points at one frequency (blockIdx.y); lane = point, AC_W groups of lanes split the
//   work of every phase by matrix entry.  The factors of the AC_PTS systems live in shared memory as double2
//   (re, im), [entry][lane]: consecutive lanes are consecutive sweep points, so the dev_out gathers are coalesced
//   and the shared-memory accesses of a half-warp are one contiguous 256-byte row.
//   1. assembly  (lin_g + j w lin_c for the built-in devices, a_ptr / a_src gather lists for the Verilog-A stamps)
//   2. right-looking elimination, pivot by pivot (two barriers per pivot; L is kept unscaled, pivots inverted)
//   3. AC:    forward / backward substitution with b = -dF/d(eps) (sources with ac_mag), outputs x[out]
//      noise: one adjoint solve per output, A^T y = e_out  (U^T forward, L^T backward), then
//             PSD = sum_k |y[pos_k] - y[neg_k]|^2 pwr_k / f^exp_k  over resistors (4 k T m / R) and the noise sources
//             of the Verilog-A devices (k_evaln_<model> outputs), summed by all groups and reduced through shared memory.
// Substitutions are sequential per system (one group of lanes does them): they are ~nnz complex multiply-adds
// against ~flops/AC_W per group for the elimination plus 2 N barriers.
#define AC_PTS 16
#define AC_W 16
struct NoiseTab { int pos, neg, pwr_row, exp_row; double mult; };   // pos / neg: elimination step of the KCL row or -1
struct ResTab { int pos, neg, pad0, pad1; Pref r; double mult; };
struct AArgs {
    NArgs n;
    const int* u_col;
    const int* a_csrc;        // dev_out row of dQ/dV for every a_src item
    const double* freqs;      // [F]
    const double* ac_rhs;     // [N] in elimination-step order (AC)
    const NoiseTab* ntab;     // Verilog-A noise sources
    const ResTab* rtab;       // resistors
    const double* noise_out;  // [rows][B] outputs of k_evaln_*
    int nnoise, nres, F, pad;
    double temp_val; int temp_col, pad2;
    double* out;              // AC: [O][F][B][2]; noise: [O][F][B]
};

__device__ __forceinline__ double2 cmul(const double2 a, const double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 crcp(const double2 a) {
    // Smith's formula: no overflow / underflow of re^2 + im^2 (entries span ~1e-20 .. 1e+3 over 17 decades of frequency)
    if (fabs(a.x) >= fabs(a.y)) {
        const double r = a.y / a.x, d = a.x + a.y * r;
        return make_double2(1.0 / d, -r / d);
    }
    const double r = a.x / a.y, d = a.x * r + a.y;
    return make_double2(r / d, -1.0 / d);
}

template <bool NOISE>
__global__ void __launch_bounds__(AC_PTS * AC_W) k_ac(const AArgs c) {
    extern __shared__ double2 av_[];
    __shared__ double s_acc[AC_W][AC_PTS];
    const NArgs& a = c.n;
    const long long B = a.B;
    const int lane = threadIdx.x % AC_PTS, w = threadIdx.x / AC_PTS;
    const long long i0 = (long long)blockIdx.x * AC_PTS + lane;
    const bool on = i0 < B;
    const long long inst = on ? i0 : B - 1;   // idle lanes shadow the last point, never store
    const int fi = blockIdx.y;
    const double freq = c.freqs[fi], omega = 6.283185307179586476925286766559 * freq;
    const int N = a.N, nnz = a.nnz_lu;
    double2* __restrict__ vals = av_ + lane;
#define VA(i) vals[(size_t)(i) * AC_PTS]
    const double* __restrict__ od = a.dev_out + inst;
    // ---- 1. assembly
    for (int e = w; e < nnz; e += AC_W) {
        double2 v = make_double2(0.0, 0.0);
        const int lin = a.a_lin[e];
        if (lin >= 0) {
            const size_t li = (size_t)lin * a.lin_ent_stride + (size_t)inst * a.lin_inst_stride;
            v.x = a.lin_g[li];
            v.y = omega * a.lin_c[li];
        }
        for (int q = a.a_ptr[e]; q < a.a_ptr[e + 1]; q++) {
            const double m = a.a_mult[q];
            v.x += m * __ldg(od + (size_t)a.a_src[q] * B);
            v.y += m * omega * __ldg(od + (size_t)c.a_csrc[q] * B);
        }
        VA(e) = v;
    }
    for (int i = w; i < N; i += AC_W) VA(nnz + i) = make_double2(NOISE ? 0.0 : c.ac_rhs[i], 0.0);
    __syncthreads();
    // ---- 2. elimination
    for (int k = 0; k < N; k++) {
        const int dp = a.diag_pos[k];
        const int l0 = a.l_ptr[k], nl = a.l_ptr[k + 1] - l0, u0 = a.u_ptr[k], nu = a.u_ptr[k + 1] - u0;
        if (nl == 0 || nu == 0) {
            if (w == 0) VA(dp) = crcp(VA(dp));
            if (nl > 0) __syncthreads();   // the substitutions read the inverted pivot after the final barrier otherwise
            continue;
        }
        const double2 inv = crcp(VA(dp));
        __syncthreads();                    // every group has read the pivot
        if (w == 0) VA(dp) = inv;
        const int p0 = a.pair_ptr[k], np = nl * nu;
        for (int q = w; q < np; q += AC_W) {
            const int li = q / nu, uj = q - li * nu;
            const double2 l = cmul(VA(a.l_pos[l0 + li]), inv);
            const double2 u = VA(a.u_pos[u0 + uj]);
            const double2 t = cmul(l, u);
            double2 d = VA(a.pair_dst[p0 + q]);
            d.x -= t.x; d.y -= t.y;
            VA(a.pair_dst[p0 + q]) = d;
        }
        __syncthreads();
    }
    __syncthreads();
    if (!NOISE) {
        // ---- 3a. A x = b: forward with M = L D^-1 (unit lower), backward with U
        if (w == 0) {
            for (int k = 0; k < N; k++) {
                const double2 bk = cmul(VA(nnz + k), VA(a.diag_pos[k]));
                for (int p = a.l_ptr[k]; p < a.l_ptr[k + 1]; p++) {
                    const double2 t = cmul(VA(a.l_pos[p]), bk);
                    double2 d = VA(nnz + a.l_row[p]);
                    d.x -= t.x; d.y -= t.y;
                    VA(nnz + a.l_row[p]) = d;
                }
            }
            for (int k = N - 1; k >= 0; k--) {
                double2 acc = VA(nnz + k);
                for (int u = a.u_ptr[k]; u < a.u_ptr[k + 1]; u++) {
                    const double2 t = cmul(VA(a.u_pos[u]), VA(nnz + c.u_col[u]));
                    acc.x -= t.x; acc.y -= t.y;
                }
                VA(nnz + k) = cmul(acc, VA(a.diag_pos[k]));
            }
            if (on)
                for (int o = 0; o < a.O; o++) {
                    const double2 x = VA(nnz + a.col_to_step[a.outputs[o]]);
                    double2* dst = (double2*)c.out + ((size_t)o * c.F + fi) * B + inst;
                    *dst = x;
                }
        }
    } else {
        const double temp_c = c.temp_col >= 0 ? a.params[(size_t)c.temp_col * B + inst] : c.temp_val;
        const double kT4 = 4.0 * 1.380649e-23 * (temp_c + 273.15);
        for (int o = 0; o < a.O; o++) {
            // ---- 3b. A^T y = e_out:  U^T w = e (forward, column-oriented over the rows of U), then M^T y = w
            if (w == 0) {
                for (int i = 0; i < N; i++) VA(nnz + i) = make_double2(0.0, 0.0);
                VA(nnz + a.col_to_step[a.outputs[o]]) = make_double2(1.0, 0.0);
                for (int k = 0; k < N; k++) {
                    const double2 wk = cmul(VA(nnz + k), VA(a.diag_pos[k]));
                    VA(nnz + k) = wk;
                    for (int u = a.u_ptr[k]; u < a.u_ptr[k + 1]; u++) {
                        const double2 t = cmul(VA(a.u_pos[u]), wk);
                        double2 d = VA(nnz + c.u_col[u]);
                        d.x -= t.x; d.y -= t.y;
                        VA(nnz + c.u_col[u]) = d;
                    }
                }
                for (int k = N - 1; k >= 0; k--) {
                    double2 acc = make_double2(0.0, 0.0);
                    for (int p = a.l_ptr[k]; p < a.l_ptr[k + 1]; p++) {
                        const double2 t = cmul(VA(a.l_pos[p]), VA(nnz + a.l_row[p]));
                        acc.x += t.x; acc.y += t.y;
                    }
                    const double2 t = cmul(acc, VA(a.diag_pos[k]));
                    double2 y = VA(nnz + k);
                    y.x -= t.x; y.y -= t.y;
                    VA(nnz + k) = y;
                }
            }
            __syncthreads();
            // ---- 4. PSD: every group sums a slice of the sources
            double acc = 0.0;
            for (int q = w; q < c.nres; q += AC_W) {
                const ResTab r = c.rtab[q];
                double2 h = make_double2(0.0, 0.0);
                if (r.pos >= 0) h = VA(nnz + r.pos);
                if (r.neg >= 0) { const double2 g = VA(nnz + r.neg); h.x -= g.x; h.y -= g.y; }
                acc += (h.x * h.x + h.y * h.y) * (kT4 * r.mult / pv(r.r, a.params, B, inst));
            }
            for (int q = w; q < c.nnoise; q += AC_W) {
                const NoiseTab t = c.ntab[q];
                double2 h = make_double2(0.0, 0.0);
                if (t.pos >= 0) h = VA(nnz + t.pos);
                if (t.neg >= 0) { const double2 g = VA(nnz + t.neg); h.x -= g.x; h.y -= g.y; }
                const double pw = __ldg(c.noise_out + (size_t)t.pwr_row * B + inst);
                const double ex = __ldg(c.noise_out + (size_t)t.exp_row * B + inst);
                acc += (h.x * h.x + h.y * h.y) * t.mult * (ex == 0.0 ? pw : pw / pow(freq, ex));
            }
            s_acc[w][lane] = acc;
            __syncthreads();
            if (w == 0 && on) {
                double tot = 0.0;
#pragma unroll
                for (int k = 0; k < AC_W; k++) tot += s_acc[k][lane];
                c.out[((size_t)o * c.F + fi) * B + inst] = tot;
            }
            __syncthreads();
        }
    }
#undef VA
}

// marks every point as taking part in the next (full) device evaluation with alpha = 0: G and C come out separately
__global__ void k_ac_prepare(long long B, int* active, double* alpha, int* list_full, int* cnt) {
    const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst == 0) { cnt[0] = (int)B; cnt[1] = 0; }
    if (inst >= B) return;
    active[inst] = ACT_FULL;
    alpha[inst] = 0.0;
    list_full[inst] = (int)inst;
}

__global__ void k_init_waves(const NArgs a, double* WV) {
    const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= a.B) return;
    store_waves(a, WV, inst, true, 0.0);
}

// ---- small helper kernels ---------------------------------------------------------------------
__global__ void k_init_state(long long B, int N, int* ist, double* dst, double* alpha, int* active, double* X,
                             double* XN, double* BETA, const double* x0, long long x0_stride, Opts o, int* list_full, int* cnt) {
    const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst == 0) { cnt[0] = (int)B; cnt[1] = 0; }   // first round: every point, full iteration (identity list)
    if (inst >= B) return;
    list_full[inst] = (int)inst;
    for (int k = 0; k < IS_COUNT; k++) ist[(size_t)k * B + inst] = 0;
    for (int k = 0; k < DS_COUNT; k++) dst[(size_t)k * B + inst] = 0.0;
    ist[(size_t)IS_PHASE * B + inst] = o.skip_dc && !o.dc_only ? PH_TRAN_INIT : PH_DC;
    ist[(size_t)IS_STAGE * B + inst] = -1;
    dst[(size_t)DS_T * B + inst] = o.t0;
    dst[(size_t)DS_HPROP * B + inst] = o.dt > 0.0 ? o.dt : o.span * 1e-5;
    dst[(size_t)DS_KAPPA * B + inst] = o.kappa0;
    alpha[inst] = 0.0;
    active[inst] = ACT_FULL;   // the first iteration of every point is a full one (fresh Jacobian, factors stored)
    for (int i = 0; i < N; i++) {
        const double v = x0 ? (x0_stride ? x0[(size_t)i * x0_stride + inst] : x0[i]) : 0.0;
        X[(size_t)i * B + inst] = v;
        XN[(size_t)i * B + inst] = v;
        BETA[(size_t)i * B + inst] = 0.0;
    }
}

struct LinContrib { Pref p; double coef; int entry, recip, is_c, pad; };

// per-point values of the linear (R, C, L, E, G, incidence) stamps: lin_g / lin_c [nlin][Bl]
__global__ void k_lin_setup(long long Bl, long long B, int nlin, int ncontrib, const LinContrib* lc, const double* params,
                            double* lin_g, double* lin_c) {
    const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= Bl) return;
    for (int e = 0; e < nlin; e++) { lin_g[(size_t)e * Bl + inst] = 0.0; lin_c[(size_t)e * Bl + inst] = 0.0; }
    for (int c = 0; c < ncontrib; c++) {
        const LinContrib& k = lc[c];
        double v = pv(k.p, params, B, inst);
        if (k.recip) v = 1.0 / v;
        v *= k.coef;
        if (k.is_c) lin_c[(size_t)k.entry * Bl + inst] += v;
        else lin_g[(size_t)k.entry * Bl + inst] += v;
    }
}

// FP64 FMA throughput microbenchmark (roofline denominator of the device-evaluation kernels;
// MEASURED_PEAKS.json carries no FP64 figure): 8 independent DFMA chains per thread.
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double b, double c) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0,
           a6 = a0 + 6.0, a7 = a0 + 7.0;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// direct-method sensitivities: the first B entries of the value-only list, no full-iteration points
__global__ void k_list_identity_any(long long B, int* list_any, int* cnt) {
    const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst == 0) { cnt[0] = 0; cnt[1] = (int)B; }
    if (inst < B) list_any[inst] = (int)inst;
}
// sens[o][b] += sign * dx[outputs[o]][b] / (2 step[b])
__global__ void k_sens_accum(long long B, int O, const int* outputs, const double* DX, const double* step, double sign, double* sens) {
    const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= B) return;
    const double w = sign / (2.0 * step[inst]);
    for (int o = 0; o < O; o++) sens[(size_t)o * B + inst] += w * DX[(size_t)outputs[o] * B + inst];
}

__global__ void k_gather_rows(long long B, int n, const int* idx, const double* src, double* dst) {
    const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= B) return;
    for (int k = 0; k < n; k++) dst[(size_t)k * B + inst] = src[(size_t)idx[k] * B + inst];
}

}  // namespace cbk
