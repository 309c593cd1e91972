// cedarb200.cu -- C-ABI shared library of the B200 batched circuit-sweep engine
// (include/cedarb200.h).  Host side: deep copy of the flat circuit, symbolic analysis,
// NVRTC compilation of the host-generated device models for sm_100a (+ cubin cache), plan
// (device memory) management and the lock-step round driver.  Device side: kernels.cuh.
//
// There is NO CPU fallback anywhere in this file: without a CUDA device cb_plan_create fails
// with CB_ERR_NO_DEVICE.
#include "../../include/cedarb200.h"

#include <cuda_runtime.h>
#include <nvrtc.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX: ranges around the phases of a call (tracing)
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
// RAII NVTX range: the phases of every entry point show up by name in Nsight Systems / Nsight Compute (--nvtx)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define CB_NBUF 3   // rotating buffers of the per-round point lists (solve())
#include "spice_front.hpp"
#include "symbolic.hpp"
#include "va_prelude.h"

using namespace cbk;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(CB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));         \
    } while (0)

// ------------------------------------------------------------------------------------------
struct WaveH {
    int kind, has_dc;
    cb_pref dc;
    std::vector<double> t;
    std::vector<cb_pref> y;
    cb_pref v[7];
    double ac_mag = 0.0;
};
struct ModelH {
    std::string name;
    int nterm, nparam, ncache, nj;
    std::vector<int> jrow, jcol;
    int ncache_n = 0;
    bool linear = false;                     // bias-independent Jacobian (cb_va_model.linear)
    std::vector<int> noise_pos, noise_neg;   // noise sources (terminal indices, -1 = ground)
    int nout() const { return 2 * nterm + 2 * nj; }   // I | Q | J = dI/dV + alpha dQ/dV | C = dQ/dV
};
struct VaInstH {
    int model;
    std::vector<int> term;
    std::vector<cb_pref> par;
    std::vector<uint8_t> given;
    double mult;
};

struct cb_circuit {
    int N = 0, NV = 0, P = 0;
    std::vector<cb_device> devs;
    std::vector<WaveH> waves;
    std::vector<ModelH> models;
    std::vector<VaInstH> insts;
    std::vector<int> outputs;
    std::string cuda_source;
    std::vector<char> cubin;
    bool compiled = false;
    cb::Symbolic sym;
    // assembly tables (host copies)
    std::vector<int> a_ptr, a_src, a_lin;
    std::vector<int> a_csrc;   // dev_out row of dQ/dV next to every a_src row (small-signal analyses)
    std::vector<double> a_mult;
    std::vector<uint8_t> a_diag;
    std::vector<int> ri_ptr, ri_src, rq_ptr, rq_src, rl_ptr, rl_col, rl_lin, rs_ptr, rs_wave;
    std::vector<double> ri_mult, rq_mult, rs_coef;
    // charge-update items: q_row += mult * C[src row of dev_out] * dx[col]
    std::vector<int> cq_row, cq_src, cq_col;
    std::vector<double> cq_mult;
    std::vector<LinContrib> lin_contrib;
    int nlin = 0;
    bool lin_swept = false;
    std::vector<uint8_t> lte_mask;
    // per-model device instance lists and slot offsets
    std::vector<std::vector<int>> model_insts;  // model -> global inst ids
    std::vector<long long> out_off, cache_off;  // per model, in slots
    long long total_out = 0, total_cache = 0;
    std::vector<long long> inst_out_base;       // per inst: first slot
};

extern "C" int cb_version(void) { return CB_ABI_VERSION; }
extern "C" const char* cb_last_error(void) { return g_err.c_str(); }

extern "C" size_t cb_options_size(void) { return sizeof(cb_options); }
extern "C" size_t cb_stats_size(void) { return sizeof(cb_stats); }
extern "C" int cb_options_init(cb_options* o, size_t caller_sizeof) {
    if (!o) return fail(CB_ERR_INVALID, "null argument");
    if (caller_sizeof != sizeof(cb_options))
        return fail(CB_ERR_INVALID, "cb_options layout mismatch: the caller's struct has " + std::to_string(caller_sizeof) +
                                        " bytes, this library's has " + std::to_string(sizeof(cb_options)) + " (ABI version " +
                                        std::to_string(CB_ABI_VERSION) + ")");
    cb_options_default(o);
    return CB_OK;
}
// every entry point that takes options: refuse a struct that was not initialised by this library's layout
static int check_options(const cb_options* o) {
    if (!o) return fail(CB_ERR_INVALID, "null options");
    if (o->struct_size != sizeof(cb_options) || o->abi_version != CB_ABI_VERSION)
        return fail(CB_ERR_INVALID, "cb_options was not initialised by cb_options_default / cb_options_init of this library (size " +
                                        std::to_string(o->struct_size) + " / ABI " + std::to_string(o->abi_version) + ", expected " +
                                        std::to_string(sizeof(cb_options)) + " / " + std::to_string(CB_ABI_VERSION) + ")");
    return CB_OK;
}

extern "C" void cb_options_default(cb_options* o) {
    std::memset(o, 0, sizeof(*o));
    o->struct_size = (uint32_t)sizeof(cb_options); o->abi_version = CB_ABI_VERSION;
    o->mixed_rounds = 0; o->source_steps = 10; o->t0_reinit = 1; o->pivot_repair = 0; o->pivot_growth_max = 1e14;
    o->temp.value = 27.0; o->temp.col = -1;
    o->gmin.value = 1e-12; o->gmin.col = -1;
    o->reltol = 1e-3; o->vabstol = 1e-6; o->iabstol = 1e-12;
    o->nr_reltol = 1e-7; o->nr_vabstol = 1e-10; o->nr_iabstol = 1e-13;
    o->dc_abstol = 1e-10; o->dv_max = 0.5;
    o->max_newton_dc = 200; o->max_newton_tran = 20;
    o->method = CB_METHOD_TRAP; o->fixed_step = 0;
    o->dt = 0; o->dt_min = 1e-18; o->dt_max = 0;
    o->gmin_steps = 10; o->skip_dc = 0;
}

extern "C" int cb_circuit_create(const cb_flat_circuit* f, cb_circuit** out) {
    if (!f || !out) return fail(CB_ERR_INVALID, "null argument");
    if (f->n_unknowns <= 0 || f->n_nodes > f->n_unknowns) return fail(CB_ERR_INVALID, "bad unknown counts");
    auto c = std::make_unique<cb_circuit>();
    c->N = f->n_unknowns; c->NV = f->n_nodes; c->P = f->n_params;
    auto chk = [&](int idx) { return idx >= -1 && idx < c->N; };
    auto chkp = [&](const cb_pref& p) { return p.col >= -1 && p.col < f->n_params; };
    for (int i = 0; i < f->n_devices; i++) {
        const cb_device& d = f->devices[i];
        for (int k = 0; k < 4; k++) if (!chk(d.n[k])) return fail(CB_ERR_INVALID, "device node index out of range");
        if (!chk(d.branch) || !chkp(d.value)) return fail(CB_ERR_INVALID, "device branch/param index out of range");
        if ((d.kind == CB_DEV_VSRC || d.kind == CB_DEV_ISRC) && (d.wave < 0 || d.wave >= f->n_waves))
            return fail(CB_ERR_INVALID, "source without waveform");
        if ((d.kind == CB_DEV_L || d.kind == CB_DEV_VSRC || d.kind == CB_DEV_VCVS) && d.branch < 0)
            return fail(CB_ERR_INVALID, "voltage-defined device without branch unknown");
        c->devs.push_back(d);
    }
    for (int i = 0; i < f->n_waves; i++) {
        const cb_wave& w = f->waves[i];
        WaveH h;
        h.kind = w.kind; h.has_dc = w.has_dc; h.dc = w.dc;
        if (w.kind == CB_W_PWL) {
            if (w.npts <= 0) return fail(CB_ERR_INVALID, "empty PWL");
            h.t.assign(w.t, w.t + w.npts);
            h.y.assign(w.y, w.y + w.npts);
        }
        for (int k = 0; k < 7; k++) h.v[k] = w.v[k];
        h.ac_mag = w.ac_mag;
        // every parameter reference of a waveform must name a row of params[P][B] (or -1): they are dereferenced on the device
        if (!chkp(w.dc)) return fail(CB_ERR_INVALID, "waveform dc parameter column out of range");
        for (const cb_pref& y : h.y) if (!chkp(y)) return fail(CB_ERR_INVALID, "PWL value parameter column out of range");
        if (w.kind == CB_W_PULSE || w.kind == CB_W_SIN)
            for (int k = 0; k < 7; k++) if (!chkp(w.v[k])) return fail(CB_ERR_INVALID, "waveform parameter column out of range");
        if (w.kind == CB_W_PULSE) {
            for (int k = 2; k < 7; k++)
                if (w.v[k].col >= 0) return fail(CB_ERR_INVALID, "PULSE timing parameters cannot be swept");
            // a period <= 0 (or NaN) means "no period" as in SPICE: one pulse (fmod(t, 0) is NaN, and the breakpoint
            // collection would never terminate)
            if (!(h.v[6].value > 0.0)) h.v[6].value = INFINITY;
            for (int k = 2; k < 6; k++)
                if (!(h.v[k].value >= 0.0) || !std::isfinite(h.v[k].value)) return fail(CB_ERR_INVALID, "PULSE td / tr / tf / pw must be finite and >= 0");
        }
        c->waves.push_back(std::move(h));
    }
    for (int i = 0; i < f->n_va_models; i++) {
        const cb_va_model& m = f->va_models[i];
        ModelH h;
        h.name = m.name ? m.name : "";
        h.nterm = m.nterm; h.nparam = m.nparam; h.ncache = m.ncache; h.nj = m.nj;
        h.jrow.assign(m.jrow, m.jrow + m.nj);
        h.jcol.assign(m.jcol, m.jcol + m.nj);
        h.ncache_n = m.ncache_n;
        if (m.n_noise > 0) {
            if (!m.noise_pos || !m.noise_neg) return fail(CB_ERR_INVALID, "noise source tables missing");
            h.noise_pos.assign(m.noise_pos, m.noise_pos + m.n_noise);
            h.noise_neg.assign(m.noise_neg, m.noise_neg + m.n_noise);
            for (int k = 0; k < m.n_noise; k++)
                if (h.noise_pos[k] < -1 || h.noise_pos[k] >= m.nterm || h.noise_neg[k] < -1 || h.noise_neg[k] >= m.nterm)
                    return fail(CB_ERR_INVALID, "noise source terminal out of range");
        }
        h.linear = m.linear != 0;
        c->models.push_back(std::move(h));
    }
    for (int i = 0; i < f->n_va_insts; i++) {
        const cb_va_inst& v = f->va_insts[i];
        if (v.model < 0 || v.model >= f->n_va_models) return fail(CB_ERR_INVALID, "VA instance model index");
        const ModelH& m = c->models[v.model];
        VaInstH h;
        h.model = v.model; h.mult = v.mult;
        h.term.assign(v.term, v.term + m.nterm);
        for (int t : h.term) if (!chk(t)) return fail(CB_ERR_INVALID, "VA terminal index out of range");
        h.par.assign(v.par, v.par + m.nparam);
        h.given.assign(v.given, v.given + m.nparam);
        for (int k = 0; k < m.nparam; k++) {
            if (!h.given[k]) { h.par[k].col = -1; continue; }   // only meaningful where given: never dereferenced otherwise
            if (!chkp(h.par[k])) return fail(CB_ERR_INVALID, "VA instance parameter column out of range");
        }
        c->insts.push_back(std::move(h));
    }
    c->outputs.assign(f->outputs, f->outputs + f->n_outputs);
    for (int o : c->outputs) if (o < 0 || o >= c->N) return fail(CB_ERR_INVALID, "output index out of range");
    *out = c.release();
    return CB_OK;
}


// ---- "flatckt" files: a flat circuit + its generated device code, written by the host front end
//      (cedarsim.jl_b200/flat.py save_flatckt, `python -m cedarsim.jl_b200.flatten`) so that a binding without a
//      front end of its own (the Julia extension, ext/CedarSimB200Ext.jl) needs no struct packing: little-endian,
//      "CBFC" u32 version | 8 x i32 counts | cb_device[] | waves | models | instances | outputs | u64 len, CUDA C.
namespace {
struct Reader {
    const unsigned char* p; size_t n, off = 0; bool ok = true;
    template <class T> T get() { T v{}; if (off + sizeof(T) > n) { ok = false; return v; } std::memcpy(&v, p + off, sizeof(T)); off += sizeof(T); return v; }
    template <class T> void vec(std::vector<T>& v, size_t cnt) {
        if (cnt > (n - std::min(n, off)) / sizeof(T)) { ok = false; return; }
        v.resize(cnt); if (cnt) std::memcpy(v.data(), p + off, cnt * sizeof(T)); off += cnt * sizeof(T);
    }
};
}  // namespace

// ---- native netlist front end (spice_front.hpp): deck text + sweep values -> flat circuit + params[P][B] ----------------
static int netlist_flatten(bool spectre, const char* text, const char* base_dir, const char* const* sweep_names, int n_sweep,
                           const double* sweep_values, int64_t n_inst, const char* const* output_names, int n_outputs,
                           cb_netlist** out) {
    if (!text || !out || n_sweep < 0 || n_outputs < 0 || n_inst < 1 || (n_sweep > 0 && (!sweep_names || !sweep_values)) ||
        (n_outputs > 0 && !output_names))
        return fail(CB_ERR_INVALID, "null / negative argument");
    try {
        auto h = std::make_unique<cb_netlist>();
        if (spectre) sf::parse_spectre_into(h->nl, text, base_dir ? base_dir : "", 0);
        else sf::parse_into(h->nl, text, true, base_dir ? base_dir : "", 0);
        std::vector<std::string> names, outs;
        for (int k = 0; k < n_sweep; k++) names.push_back(sweep_names[k] ? sweep_names[k] : "");
        for (int k = 0; k < n_outputs; k++) outs.push_back(output_names[k] ? output_names[k] : "");
        sf::Flattener(h->nl, h->fc).run(names, sweep_values, n_inst, outs);
        sf::Flat& fc = h->fc;
        h->unknown_names = fc.node_names;
        h->unknown_names.insert(h->unknown_names.end(), fc.branch_names.begin(), fc.branch_names.end());
        if (fc.outputs.empty() && n_outputs == 0)
            for (size_t k = 0; k < h->unknown_names.size(); k++) fc.outputs.push_back((int32_t)k);   // default: every unknown
        if (h->unknown_names.empty()) return fail(CB_ERR_INVALID, "the deck has no unknowns");
        const size_t P = fc.columns.size();
        h->params.resize(P * (size_t)n_inst);
        for (size_t k = 0; k < P; k++) std::copy(fc.columns[k].begin(), fc.columns[k].end(), h->params.begin() + k * (size_t)n_inst);
        cb_flat_circuit& f = h->flat;
        f.n_unknowns = (int32_t)h->unknown_names.size();
        f.n_nodes = (int32_t)fc.node_names.size();
        f.n_params = (int32_t)P;
        f.n_devices = (int32_t)fc.devices.size();
        f.devices = fc.devices.data();
        f.n_waves = (int32_t)fc.waves.size();
        f.waves = fc.waves.data();
        f.n_va_models = 0; f.va_models = nullptr; f.n_va_insts = 0; f.va_insts = nullptr;
        f.n_outputs = (int32_t)fc.outputs.size();
        f.outputs = fc.outputs.data();
        *out = h.release();
        return CB_OK;
    } catch (const std::exception& e) {
        return fail(CB_ERR_INVALID, std::string("netlist: ") + e.what());
    }
}

extern "C" int cb_netlist_flatten(const char* text, const char* base_dir, const char* const* sweep_names, int n_sweep,
                                  const double* sweep_values, int64_t n_inst, const char* const* output_names, int n_outputs,
                                  cb_netlist** out) {
    return netlist_flatten(false, text, base_dir, sweep_names, n_sweep, sweep_values, n_inst, output_names, n_outputs, out);
}
extern "C" int cb_netlist_flatten_spectre(const char* text, const char* base_dir, const char* const* sweep_names, int n_sweep,
                                          const double* sweep_values, int64_t n_inst, const char* const* output_names,
                                          int n_outputs, cb_netlist** out) {
    return netlist_flatten(true, text, base_dir, sweep_names, n_sweep, sweep_values, n_inst, output_names, n_outputs, out);
}

extern "C" int cb_netlist_circuit(cb_netlist* nl, cb_circuit** out) {
    if (!nl || !out) return fail(CB_ERR_INVALID, "null argument");
    return cb_circuit_create(&nl->flat, out);
}
extern "C" const cb_flat_circuit* cb_netlist_flat(const cb_netlist* nl) { return nl ? &nl->flat : nullptr; }
extern "C" int32_t cb_netlist_n_unknowns(const cb_netlist* nl) { return nl ? nl->flat.n_unknowns : -1; }
extern "C" int32_t cb_netlist_n_params(const cb_netlist* nl) { return nl ? nl->flat.n_params : -1; }
extern "C" const double* cb_netlist_params(const cb_netlist* nl) { return nl ? nl->params.data() : nullptr; }
extern "C" const char* cb_netlist_unknown_name(const cb_netlist* nl, int32_t i) {
    return (nl && i >= 0 && i < (int32_t)nl->unknown_names.size()) ? nl->unknown_names[(size_t)i].c_str() : nullptr;
}
extern "C" const char* cb_netlist_param_name(const cb_netlist* nl, int32_t i) {
    return (nl && i >= 0 && i < (int32_t)nl->fc.param_names.size()) ? nl->fc.param_names[(size_t)i].c_str() : nullptr;
}
extern "C" int32_t cb_netlist_unknown(const cb_netlist* nl, const char* name) { return (nl && name) ? nl->fc.unknown(name) : -1; }
extern "C" double cb_netlist_option(const cb_netlist* nl, const char* name, double dflt) {
    if (!nl || !name) return dflt;
    auto it = nl->fc.options.find(sf::lower(name));
    return it == nl->fc.options.end() ? dflt : it->second;
}
extern "C" void cb_netlist_destroy(cb_netlist* nl) { delete nl; }

extern "C" int cb_circuit_load(const char* path, cb_circuit** out) {
    if (!path || !out) return fail(CB_ERR_INVALID, "null argument");
    std::vector<unsigned char> buf;
    {
        FILE* f = std::fopen(path, "rb");
        if (!f) return fail(CB_ERR_INVALID, std::string("cannot open ") + path);
        std::fseek(f, 0, SEEK_END); const long sz = std::ftell(f); std::fseek(f, 0, SEEK_SET);
        buf.resize(sz > 0 ? (size_t)sz : 0);
        const size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
        std::fclose(f);
        if (got != buf.size()) return fail(CB_ERR_INVALID, std::string("short read of ") + path);
    }
    Reader r{buf.data(), buf.size()};
    if (r.get<uint32_t>() != 0x43464243u /* "CBFC" */ || r.get<uint32_t>() != 1u) return fail(CB_ERR_INVALID, "not a flatckt v1 file");
    cb_flat_circuit fc{};
    fc.n_unknowns = r.get<int32_t>(); fc.n_nodes = r.get<int32_t>(); fc.n_params = r.get<int32_t>(); fc.n_devices = r.get<int32_t>();
    fc.n_waves = r.get<int32_t>(); fc.n_va_models = r.get<int32_t>(); fc.n_va_insts = r.get<int32_t>(); fc.n_outputs = r.get<int32_t>();
    if (!r.ok || fc.n_devices < 0 || fc.n_waves < 0 || fc.n_va_models < 0 || fc.n_va_insts < 0 || fc.n_outputs < 0)
        return fail(CB_ERR_INVALID, "flatckt: bad header");
    std::vector<cb_device> devs; r.vec(devs, (size_t)fc.n_devices);
    std::vector<cb_wave> waves((size_t)fc.n_waves);
    std::vector<std::vector<double>> wt((size_t)fc.n_waves);
    std::vector<std::vector<cb_pref>> wy((size_t)fc.n_waves);
    for (int i = 0; i < fc.n_waves && r.ok; i++) {
        cb_wave& w = waves[i]; w = cb_wave{};
        w.kind = r.get<int32_t>(); w.has_dc = r.get<int32_t>(); w.dc = r.get<cb_pref>(); w.npts = r.get<int32_t>();
        if (w.npts < 0) { r.ok = false; break; }
        r.vec(wt[i], (size_t)w.npts); r.vec(wy[i], (size_t)w.npts);
        for (int k = 0; k < 7; k++) w.v[k] = r.get<cb_pref>();
        w.ac_mag = r.get<double>();
        w.t = wt[i].data(); w.y = wy[i].data();
    }
    std::vector<cb_va_model> models((size_t)fc.n_va_models);
    std::vector<std::string> mname((size_t)fc.n_va_models);
    std::vector<std::vector<int32_t>> mjr(models.size()), mjc(models.size()), mnp(models.size()), mnn(models.size());
    for (size_t i = 0; i < models.size() && r.ok; i++) {
        cb_va_model& m = models[i]; m = cb_va_model{};
        std::vector<char> nm; r.vec(nm, r.get<uint32_t>()); mname[i].assign(nm.begin(), nm.end());
        m.nterm = r.get<int32_t>(); m.nparam = r.get<int32_t>(); m.ncache = r.get<int32_t>(); m.nj = r.get<int32_t>();
        if (m.nterm < 0 || m.nparam < 0 || m.nj < 0) { r.ok = false; break; }
        r.vec(mjr[i], (size_t)m.nj); r.vec(mjc[i], (size_t)m.nj);
        m.n_noise = r.get<int32_t>(); m.ncache_n = r.get<int32_t>();
        if (m.n_noise < 0) { r.ok = false; break; }
        r.vec(mnp[i], (size_t)m.n_noise); r.vec(mnn[i], (size_t)m.n_noise);
        m.linear = r.get<int32_t>();
        m.name = mname[i].c_str(); m.jrow = mjr[i].data(); m.jcol = mjc[i].data(); m.noise_pos = mnp[i].data(); m.noise_neg = mnn[i].data();
    }
    std::vector<cb_va_inst> insts((size_t)fc.n_va_insts);
    std::vector<std::vector<int32_t>> it(insts.size());
    std::vector<std::vector<cb_pref>> ip(insts.size());
    std::vector<std::vector<uint8_t>> ig(insts.size());
    for (size_t i = 0; i < insts.size() && r.ok; i++) {
        cb_va_inst& v = insts[i]; v = cb_va_inst{};
        v.model = r.get<int32_t>();
        if (v.model < 0 || v.model >= fc.n_va_models) { r.ok = false; break; }
        r.vec(it[i], (size_t)models[v.model].nterm); r.vec(ip[i], (size_t)models[v.model].nparam); r.vec(ig[i], (size_t)models[v.model].nparam);
        v.mult = r.get<double>();
        v.term = it[i].data(); v.par = ip[i].data(); v.given = ig[i].data();
    }
    std::vector<int32_t> outs; r.vec(outs, (size_t)fc.n_outputs);
    const uint64_t slen = r.get<uint64_t>();
    std::vector<char> src; r.vec(src, (size_t)slen);
    if (!r.ok) return fail(CB_ERR_INVALID, "flatckt: truncated or corrupt file");
    fc.devices = devs.data(); fc.waves = waves.data(); fc.va_models = models.data(); fc.va_insts = insts.data(); fc.outputs = outs.data();
    cb_circuit* c = nullptr;
    int rc = cb_circuit_create(&fc, &c);
    if (rc != CB_OK) return rc;
    if (!src.empty()) c->cuda_source.assign(src.data(), src.size());
    *out = c;
    return CB_OK;
}

extern "C" int cb_circuit_set_cuda_source(cb_circuit* c, const char* src, size_t len) {
    if (!c || !src) return fail(CB_ERR_INVALID, "null argument");
    c->cuda_source.assign(src, len);
    c->compiled = false;
    return CB_OK;
}

// slots of one per-instance cache row as the device code lays it out: whole 32-byte sectors (va_prelude.h, NCACHE_P)
static long long pad4(int n) { return ((long long)std::max(1, n) + 3) / 4 * 4; }

static uint64_t fnv1a(const std::string& s) {
    uint64_t h = 1469598103934665603ULL;
    for (unsigned char ch : s) { h ^= ch; h *= 1099511628211ULL; }
    return h;
}

static int build_tables(cb_circuit* c) {
    const int N = c->N;
    std::vector<cb::PatternEntry> pat;
    for (int i = 0; i < c->NV; i++) pat.push_back({i, i, 0});  // gshunt slot, not a pivot by itself
    // ---- linear devices: pattern + contributions per (row, col)
    std::map<std::pair<int, int>, int> lin_index;
    auto lin_of = [&](int r, int col) {
        auto key = std::make_pair(r, col);
        auto it = lin_index.find(key);
        if (it != lin_index.end()) return it->second;
        int idx = (int)lin_index.size();
        lin_index[key] = idx;
        return idx;
    };
    auto add = [&](int r, int col, int cls, cb_pref p, double coef, bool recip, bool is_c) {
        if (r < 0 || col < 0) return;
        pat.push_back({r, col, cls});
        LinContrib k;
        k.p.value = p.value; k.p.col = p.col; k.p.pad = 0;
        k.coef = coef; k.entry = lin_of(r, col); k.recip = recip; k.is_c = is_c; k.pad = 0;
        if (p.col >= 0) c->lin_swept = true;
        c->lin_contrib.push_back(k);
    };
    const cb_pref one = {1.0, -1, 0};
    std::vector<std::vector<std::pair<int, double>>> src_rows(N);
    c->lte_mask.assign(N, 0);
    for (int i = 0; i < c->NV; i++) c->lte_mask[i] = 1;
    for (const cb_device& d : c->devs) {
        const int p = d.n[0], n = d.n[1], cp = d.n[2], cn = d.n[3], b = d.branch;
        const double m = d.mult;
        switch (d.kind) {
            case CB_DEV_R:
                add(p, p, 2, d.value, m, true, false); add(p, n, 1, d.value, -m, true, false);
                add(n, p, 1, d.value, -m, true, false); add(n, n, 2, d.value, m, true, false);
                break;
            case CB_DEV_C:
                add(p, p, 1, d.value, m, false, true); add(p, n, 1, d.value, -m, false, true);
                add(n, p, 1, d.value, -m, false, true); add(n, n, 1, d.value, m, false, true);
                break;
            case CB_DEV_L:
                add(p, b, 3, one, m, false, false); add(n, b, 3, one, -m, false, false);
                add(b, p, 3, one, 1.0, false, false); add(b, n, 3, one, -1.0, false, false);
                add(b, b, 0, d.value, -1.0, false, true);
                c->lte_mask[b] = 1;
                break;
            case CB_DEV_VSRC:
                add(p, b, 3, one, m, false, false); add(n, b, 3, one, -m, false, false);
                add(b, p, 3, one, 1.0, false, false); add(b, n, 3, one, -1.0, false, false);
                src_rows[b].push_back({d.wave, -1.0});
                break;
            case CB_DEV_ISRC:
                if (p >= 0) src_rows[p].push_back({d.wave, m});
                if (n >= 0) src_rows[n].push_back({d.wave, -m});
                break;
            case CB_DEV_VCVS:
                add(p, b, 3, one, m, false, false); add(n, b, 3, one, -m, false, false);
                add(b, p, 3, one, 1.0, false, false); add(b, n, 3, one, -1.0, false, false);
                add(b, cp, 1, d.value, -1.0, false, false); add(b, cn, 1, d.value, 1.0, false, false);
                break;
            case CB_DEV_VCCS:
                add(p, cp, 1, d.value, m, false, false); add(p, cn, 1, d.value, -m, false, false);
                add(n, cp, 1, d.value, -m, false, false); add(n, cn, 1, d.value, m, false, false);
                break;
            default:
                return fail(CB_ERR_INVALID, "unknown device kind");
        }
    }
    c->nlin = (int)lin_index.size();
    // ---- VA devices: slot numbering [model][device][slot]
    const int nm = (int)c->models.size();
    c->model_insts.assign(nm, {});
    for (int i = 0; i < (int)c->insts.size(); i++) c->model_insts[c->insts[i].model].push_back(i);
    c->out_off.assign(nm, 0); c->cache_off.assign(nm, 0);
    c->total_out = 0; c->total_cache = 0;
    c->inst_out_base.assign(c->insts.size(), 0);
    for (int m = 0; m < nm; m++) {
        c->out_off[m] = c->total_out;
        c->cache_off[m] = c->total_cache;
        for (size_t k = 0; k < c->model_insts[m].size(); k++)
            c->inst_out_base[c->model_insts[m][k]] = c->total_out + (long long)k * c->models[m].nout();
        c->total_out += (long long)c->model_insts[m].size() * c->models[m].nout();
        c->total_cache += (long long)c->model_insts[m].size() * pad4(c->models[m].ncache);
    }
    for (size_t i = 0; i < c->insts.size(); i++) {
        const VaInstH& v = c->insts[i];
        const ModelH& m = c->models[v.model];
        for (int j = 0; j < m.nj; j++) {
            const int r = v.term[m.jrow[j]], col = v.term[m.jcol[j]];
            if (r >= 0 && col >= 0) pat.push_back({r, col, m.jrow[j] == m.jcol[j] ? 2 : 1});
        }
    }
    if (!cb::analyze(N, pat, c->sym)) return fail(CB_ERR_SINGULAR, c->sym.error);
    const cb::Symbolic& S = c->sym;
    // ---- J assembly lists per LU entry
    std::vector<std::vector<std::pair<int, double>>> ent_src(S.nnz_lu);
    c->a_lin.assign(S.nnz_lu, -1);
    c->a_diag.assign(S.nnz_lu, 0);
    for (const auto& kv : lin_index) c->a_lin[S.pos_of_orig.at(kv.first)] = kv.second;
    for (int i = 0; i < c->NV; i++) c->a_diag[S.pos_of_orig.at({i, i})] = 1;
    std::vector<std::vector<std::pair<int, double>>> rowI(N), rowQ(N);
    for (size_t i = 0; i < c->insts.size(); i++) {
        const VaInstH& v = c->insts[i];
        const ModelH& m = c->models[v.model];
        const long long base = c->inst_out_base[i];
        for (int k = 0; k < m.nterm; k++) {
            if (v.term[k] < 0) continue;
            rowI[v.term[k]].push_back({(int)(base + k), v.mult});
            rowQ[v.term[k]].push_back({(int)(base + m.nterm + k), v.mult});
        }
        for (int j = 0; j < m.nj; j++) {
            const int r = v.term[m.jrow[j]], col = v.term[m.jcol[j]];
            if (r < 0 || col < 0) continue;
            ent_src[S.pos_of_orig.at({r, col})].push_back({(int)(base + 2 * m.nterm + j), v.mult});
            c->cq_row.push_back(r); c->cq_src.push_back((int)(base + 2 * m.nterm + m.nj + j)); c->cq_col.push_back(col);
            c->cq_mult.push_back(v.mult);
        }
    }
    auto flatten = [](const std::vector<std::vector<std::pair<int, double>>>& src, std::vector<int>& ptr,
                      std::vector<int>& idx, std::vector<double>& mult) {
        ptr.assign(1, 0); idx.clear(); mult.clear();
        for (const auto& l : src) {
            for (const auto& p : l) { idx.push_back(p.first); mult.push_back(p.second); }
            ptr.push_back((int)idx.size());
        }
    };
    flatten(ent_src, c->a_ptr, c->a_src, c->a_mult);
    {   // the dQ/dV row of a Jacobian stamp sits nj rows after its J row (OUT_J in va_prelude.h)
        std::map<int, int> shift;
        for (size_t i = 0; i < c->insts.size(); i++) {
            const ModelH& m = c->models[c->insts[i].model];
            for (int j = 0; j < m.nj; j++) shift[(int)(c->inst_out_base[i] + 2 * m.nterm + j)] = m.nj;
        }
        c->a_csrc.resize(c->a_src.size());
        for (size_t q = 0; q < c->a_src.size(); q++) c->a_csrc[q] = c->a_src[q] + shift.at(c->a_src[q]);
    }
    flatten(rowI, c->ri_ptr, c->ri_src, c->ri_mult);
    flatten(rowQ, c->rq_ptr, c->rq_src, c->rq_mult);
    flatten(src_rows, c->rs_ptr, c->rs_wave, c->rs_coef);
    c->rl_ptr.assign(1, 0); c->rl_col.clear(); c->rl_lin.clear();
    {
        std::vector<std::vector<std::pair<int, int>>> rows(N);
        for (const auto& kv : lin_index) rows[kv.first.first].push_back({kv.first.second, kv.second});
        for (int i = 0; i < N; i++) {
            for (auto& p : rows[i]) { c->rl_col.push_back(p.first); c->rl_lin.push_back(p.second); }
            c->rl_ptr.push_back((int)c->rl_col.size());
        }
    }
    return CB_OK;
}


// ------------------------------------------------------------------------------------------
// Circuit-specialised straight-line kernel k_solve: residual + Jacobian assembly from the device
// outputs, static-pivot sparse LU (right-looking, fused forward substitution) and backward
// substitution, all on named scalars -- one thread per sweep point, every global access a
// coalesced [k][B] row, no index loads, no barriers.  Generated from the symbolic analysis; linear
// stamp values that are not swept are baked in as literals so that e.g. the +-1 incidence entries
// of voltage sources fold away.
static std::string dlit(double v) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.17g", v);
    std::string s(buf);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    if (s.find("inf") != std::string::npos || s.find("nan") != std::string::npos) return "(0.0/0.0)";
    return v < 0 ? "(" + s + ")" : s;
}

static void host_lin_values(const cb_circuit* c, std::vector<double>& g, std::vector<double>& cc) {
    g.assign(std::max(1, c->nlin), 0.0);
    cc.assign(std::max(1, c->nlin), 0.0);
    for (const LinContrib& k : c->lin_contrib) {
        double v = k.p.value;
        if (k.recip) v = 1.0 / v;
        v *= k.coef;
        (k.is_c ? cc : g)[k.entry] += v;
    }
}

static std::string gen_solve_source(const cb_circuit* c) {
    const cb::Symbolic& S = c->sym;
    const int N = c->N, NV = c->NV;
    std::vector<double> lg, lc;
    host_lin_values(c, lg, lc);
    std::ostringstream o;
    o << "\n// ---- generated: assembly + static-pivot LU + solve for this circuit (N=" << N << ", nnz(L+U)=" << S.nnz_lu << ")\n";
    o << "struct SArgs { long long B; const double* X; const double* alpha; const double* gshunt; const double* BETA;\n"
         "  const double* dev_out; const double* lin_g; const double* lin_c; const double* WV; const int* active;\n"
         "  double* DX; double* QK; double* RMAX; int* BAD; double* DVMAX; };\n";
    o << "extern \"C\" __global__ void __launch_bounds__(64, 1) k_solve(SArgs a) {\n";
    o << "  const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;\n";
    o << "  if (inst >= a.B) return;\n  if (a.active[inst] != 1 && a.active[inst] != 2) return;\n";
    o << "  const size_t B = (size_t)a.B;\n";
    o << "  const double al = a.alpha[inst], gs = a.gshunt[inst];\n";
    o << "  const double* __restrict__ od = a.dev_out + inst;\n";
    o << "#define OD(s) __ldg(od + (size_t)(s) * B)\n";
    for (int i = 0; i < N; i++) o << "  const double x" << i << " = a.X[" << i << " * B + inst];\n";
    for (size_t w = 0; w < c->waves.size(); w++) o << "  const double w" << w << " = a.WV[" << w << " * B + inst];\n";
    auto LG = [&](int l) { return c->lin_swept ? "lg" + std::to_string(l) : dlit(lg[l]); };
    auto LC = [&](int l) { return c->lin_swept ? "lc" + std::to_string(l) : dlit(lc[l]); };
    if (c->lin_swept)
        for (int l = 0; l < c->nlin; l++)
            o << "  const double lg" << l << " = a.lin_g[" << l << " * B + inst], lc" << l << " = a.lin_c[" << l << " * B + inst];\n";
    // Row-by-row (up-looking) elimination: row i is assembled from the device outputs right before
    // it is eliminated against the finished rows k < i, so only the U parts of finished rows stay
    // live.  Every entry receives its updates in the same ascending-k order as the right-looking
    // schedule of k_newton, so both kernels produce identical factors.
    std::map<std::pair<int, int>, int> pos;   // (row step, col step) -> LU position
    std::vector<std::vector<int>> row_cols(N);
    for (int e = 0; e < S.nnz_lu; e++) { pos[{S.lu_i[e], S.lu_j[e]}] = e; row_cols[S.lu_i[e]].push_back(S.lu_j[e]); }
    for (auto& rc : row_cols) std::sort(rc.begin(), rc.end());
    auto entry_expr = [&](int e) {
        std::ostringstream v;
        bool h = false;
        const int lin = c->a_lin[e];
        if (lin >= 0) {
            const bool g0 = !c->lin_swept && lg[lin] == 0.0, c0 = !c->lin_swept && lc[lin] == 0.0;
            if (!g0) { v << LG(lin); h = true; }
            if (!c0) { v << (h ? " + " : "") << "al * " << LC(lin); h = true; }
        }
        if (c->a_diag[e]) { v << (h ? " + " : "") << "gs"; h = true; }
        for (int p = c->a_ptr[e]; p < c->a_ptr[e + 1]; p++) {
            v << (h ? " + " : "");
            if (c->a_mult[p] != 1.0) v << dlit(c->a_mult[p]) << " * ";
            v << "OD(" << c->a_src[p] << ")";
            h = true;
        }
        return h ? v.str() : std::string("0.0");
    };
    o << "  double rmax = 0.0;\n  int bad = 0;\n";
    for (int si = 0; si < N; si++) {
        const int i = S.prow[si];   // original row eliminated at step si
        // residual of this row
        std::ostringstream f, q;
        bool hf = false, hq = false;
        for (int p = c->rl_ptr[i]; p < c->rl_ptr[i + 1]; p++) {
            const int l = c->rl_lin[p], col = c->rl_col[p];
            if (c->lin_swept || lg[l] != 0.0) { f << (hf ? " + " : "") << LG(l) << " * x" << col; hf = true; }
            if (c->lin_swept || lc[l] != 0.0) { q << (hq ? " + " : "") << LC(l) << " * x" << col; hq = true; }
        }
        for (int p = c->ri_ptr[i]; p < c->ri_ptr[i + 1]; p++) {
            f << (hf ? " + " : "");
            if (c->ri_mult[p] != 1.0) f << dlit(c->ri_mult[p]) << " * ";
            f << "OD(" << c->ri_src[p] << ")";
            hf = true;
        }
        for (int p = c->rq_ptr[i]; p < c->rq_ptr[i + 1]; p++) {
            q << (hq ? " + " : "");
            if (c->rq_mult[p] != 1.0) q << dlit(c->rq_mult[p]) << " * ";
            q << "OD(" << c->rq_src[p] << ")";
            hq = true;
        }
        for (int p = c->rs_ptr[i]; p < c->rs_ptr[i + 1]; p++) {
            f << (hf ? " + " : "") << dlit(c->rs_coef[p]) << " * w" << c->rs_wave[p];
            hf = true;
        }
        if (i < NV) { f << (hf ? " + " : "") << "gs * x" << i; hf = true; }
        o << "  const double q" << i << " = " << (hq ? q.str() : "0.0") << ";\n";
        o << "  a.QK[" << i << " * B + inst] = q" << i << ";\n";
        o << "  const double r" << i << " = " << (hf ? f.str() : "0.0") << (hq ? " + al * q" + std::to_string(i) : "")
          << " + a.BETA[" << i << " * B + inst];\n";
        o << "  rmax = fmax(rmax, fabs(r" << i << "));\n";
        o << "  double b" << si << " = -r" << i << ";\n";
        // entries of this row
        for (int j : row_cols[si]) o << "  double e" << pos[{si, j}] << " = " << entry_expr(pos[{si, j}]) << ";\n";
        // eliminate against finished rows
        for (int k : row_cols[si]) {
            if (k >= si) break;
            o << "  { const double l = e" << pos[{si, k}] << " * inv" << k << ";";
            const int up = S.u_ptr[k], nU = S.u_ptr[k + 1] - up;
            for (int uj = 0; uj < nU; uj++) o << " e" << pos[{si, S.u_col[up + uj]}] << " -= l * e" << S.u_pos[up + uj] << ";";
            o << " b" << si << " -= l * b" << k << "; }\n";
        }
        const int d = S.diag_pos[si];
        o << "  bad |= !(fabs(e" << d << ") > 0.0);\n";
        o << "  const double inv" << si << " = 1.0 / e" << d << ";\n";
    }
    o << "  a.RMAX[inst] = rmax;\n  double dvm = 0.0;\n";
    // ---- backward substitution (row oriented)
    for (int k = N - 1; k >= 0; k--) {
        o << "  const double s" << k << " = (b" << k;
        const int up = S.u_ptr[k], nU = S.u_ptr[k + 1] - up;
        for (int uj = 0; uj < nU; uj++) o << " - e" << S.u_pos[up + uj] << " * s" << S.u_col[up + uj];
        o << ") * inv" << k << ";\n";
        o << "  a.DX[" << S.pcol[k] << " * B + inst] = s" << k << ";\n";
        if (S.pcol[k] < NV) o << "  dvm = fmax(dvm, fabs(s" << k << "));\n";
        o << "  bad |= !isfinite(s" << k << ");\n";
    }
    o << "  a.DVMAX[inst] = dvm;\n";
    o << "  a.BAD[inst] = bad;\n#undef OD\n}\n";
    return o.str();
}

static int nvrtc_compile(cb_circuit* c, const char* cache_dir) {
    std::string full = std::string(CB_VA_PRELUDE) + "\n" + c->cuda_source + gen_solve_source(c);
    if (const char* dump = std::getenv("CB_DUMP_SRC")) {
        std::ofstream df(dump);
        df << full;
    }
    // experiment knobs (they change the cache key): CB_MAXREG=<n>, CB_NVRTC_DEFS="-DX=1 -DY"
    std::string maxreg = std::getenv("CB_MAXREG") ? std::string("--maxrregcount=") + std::getenv("CB_MAXREG") : "";
    std::vector<std::string> extra;
    if (const char* d = std::getenv("CB_NVRTC_DEFS")) {
        std::string ds(d);
        std::replace(ds.begin(), ds.end(), ',', ' ');
        std::istringstream is(ds);
        std::string tok;
        while (is >> tok) extra.push_back(tok);
    }
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17", "-default-device", "-w"};
    if (!maxreg.empty()) opts.push_back(maxreg.c_str());
    for (const std::string& e : extra) opts.push_back(e.c_str());
    std::string key;
    {
        std::string tag = full + "|sm_100a|v1|" + maxreg;
        for (const std::string& e : extra) tag += "|" + e;
        char buf[64];
        std::snprintf(buf, sizeof buf, "%016llx", (unsigned long long)fnv1a(tag));
        key = buf;
    }
    std::string path;
    if (cache_dir && *cache_dir) {
        path = std::string(cache_dir) + "/cb_" + key + ".cubin";
        std::ifstream in(path, std::ios::binary);
        if (in) {
            c->cubin.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
            if (!c->cubin.empty()) return CB_OK;
        }
    }
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, full.c_str(), "cb_models.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
        return fail(CB_ERR_NVRTC, "nvrtcCreateProgram failed");
    nvrtcResult r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        nvrtcGetProgramLogSize(prog, &n);
        std::string log(n, 0);
        nvrtcGetProgramLog(prog, &log[0]);
        nvrtcDestroyProgram(&prog);
        if (log.size() > 6000) log = log.substr(0, 3000) + "\n...\n" + log.substr(log.size() - 3000);
        return fail(CB_ERR_NVRTC, "NVRTC: " + log);
    }
    size_t sz = 0;
    nvrtcGetCUBINSize(prog, &sz);
    c->cubin.resize(sz);
    nvrtcGetCUBIN(prog, c->cubin.data());
    nvrtcDestroyProgram(&prog);
    if (!path.empty()) {
        std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
        std::ofstream outf(tmp, std::ios::binary);
        outf.write(c->cubin.data(), (std::streamsize)c->cubin.size());
        outf.close();
        std::rename(tmp.c_str(), path.c_str());
    }
    return CB_OK;
}

extern "C" int cb_circuit_compile(cb_circuit* c, const char* cache_dir, double* compile_seconds) {
    NvtxRange nvtx_("cb:circuit_compile (symbolic analysis, NVRTC)");
    if (!c) return fail(CB_ERR_INVALID, "null circuit");
    auto t0 = std::chrono::steady_clock::now();
    c->lin_contrib.clear();
    c->cq_row.clear(); c->cq_src.clear(); c->cq_col.clear(); c->cq_mult.clear();
    c->lin_swept = false;
    int rc = build_tables(c);
    if (rc != CB_OK) return rc;
    if (!c->models.empty() && c->cuda_source.empty())
        return fail(CB_ERR_STATE, "circuit has Verilog-A devices but no CUDA source was set");
    rc = nvrtc_compile(c, cache_dir);
    if (rc != CB_OK) return rc;
    c->compiled = true;
    if (compile_seconds)
        *compile_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return CB_OK;
}

extern "C" int cb_circuit_lu_info(const cb_circuit* c, int32_t* nnz_a, int32_t* nnz_lu, int64_t* lu_flops) {
    if (!c || !c->compiled) return fail(CB_ERR_STATE, "circuit not compiled");
    if (nnz_a) *nnz_a = c->sym.nnz_a;
    if (nnz_lu) *nnz_lu = c->sym.nnz_lu;
    if (lu_flops) *lu_flops = c->sym.flops;
    return CB_OK;
}

// ------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct cb_plan {
    // A plan over many sweep points is a set of LANES: ordinary single-stream plans over contiguous sub-ranges of the
    // points, each driven by its own host thread.  The lock-step round loop of one lane leaves SMs idle in the tail
    // wave of every kernel and runs latency-bound kernels (k_lu, k_control: one or two CTAs per SM) back to back with
    // the instruction-bound device evaluation; with several lanes in flight those gaps are filled by the other lanes'
    // kernels (+12 % points/s at 4 lanes on the 16 384-point DFF sweep).  lanes.empty() = this is a single lane.
    std::vector<cb_plan*> lanes;
    long long lane_off = 0;            // first sweep point of this lane within its parent
    long long device_points = 0;       // lane: sweep points of all lanes of its parent on this lane's GPU (0 = a plan of its own)
    bool multi_device = false;         // parent: lanes on more than one GPU (host-array entry points only)
    double* d_params_all = nullptr;    // parent: [P][B] buffer handed out by cb_plan_device_params
    bool params_all_dirty = false;
    double* d_y_all = nullptr;         // parent: contiguous results of the *_device entry points
    size_t y_all_capacity = 0;
    double* d_xout_all = nullptr;
    int* d_status_all = nullptr;
    cb_circuit* c = nullptr;
    long long B = 0, Bpad = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaLibrary_t lib = nullptr;
    std::vector<unsigned> eval_threads;
    std::vector<size_t> eval_smem;
    std::vector<cudaKernel_t> k_setup, k_eval;
    // value-only variants (null when a model has none; then every round is a full round)
    std::vector<cudaKernel_t> k_setupv, k_evalv;
    std::vector<unsigned> evalv_threads;
    std::vector<size_t> evalv_smem;
    std::vector<long long> cachev_off;
    double* d_cachev = nullptr;
    // uniform cache slots (va_prelude.h CACHE_LDU): per variant (0 full, 1 value-only, 2 noise) and model the slot count and a
    // table of [B][nuni] doubles (only the first row is used unless temperature / gmin are swept)
    std::vector<int> nuni[3];
    std::vector<double*> d_uni[3];
    // small-signal analyses (cb_ac / cb_noise): tables built on first use
    std::vector<cudaKernel_t> k_setupn, k_evaln;
    std::vector<unsigned> evaln_threads;
    std::vector<size_t> evaln_smem;
    std::vector<long long> cachen_off, noise_off;   // per model: cache slots / output rows before it
    double *d_cachen = nullptr, *d_noise_out = nullptr, *d_ac_rhs = nullptr;
    bool ac_ready = false, noise_ready = false, setupn_valid = false;
    AArgs aa{};
    size_t ac_smem = 0;
    bool have_v = false;     // all models have a value-only variant and the solve kernel is k_lu
    long long cachev_slots = 0;
    bool setupv_valid = false;
    cudaKernel_t k_solve = nullptr;
    bool lu = false;         // solve kernel = hand-written shared-memory batched LU (k_lu), else the generated k_solve
    LArgs la{};
    size_t lu_smem = 0;
    double* d_pp_scratch = nullptr;          // [SMs][N (N + 1)]: dense systems of the points k_lu re-solves with partial pivoting
    bool lu_staged = false;                  // k_lu's table blob lives in shared memory behind the matrices
    std::vector<unsigned char> lu_blob;
    int* d_dc_count = nullptr;
    // point lists of the rounds (device-wide compaction, kernels.cuh): [parity][kind][B] and counters [parity][2]
    int* d_lists = nullptr;
    int* d_cnt = nullptr;
    int num_sms = 148;
    double *d_DX = nullptr, *d_QK = nullptr, *d_RMAX = nullptr, *d_WV = nullptr, *d_DVMAX = nullptr;
    int* d_BAD = nullptr;
    std::vector<void*> allocs;
    NArgs na{};
    // device arrays
    double* d_params = nullptr;
    double* d_cache = nullptr;
    double* d_dev_out = nullptr;
    double* d_y = nullptr;
    size_t y_capacity = 0;
    double* d_saveat = nullptr;
    size_t saveat_capacity = 0;
    double* d_bp = nullptr;
    size_t bp_capacity = 0;
    double* d_xout = nullptr;
    int* d_done = nullptr;
    int* h_done = nullptr;  // pinned
    // per-model device tables
    std::vector<int*> d_term;
    std::vector<double*> d_par_val;
    std::vector<int*> d_par_col;
    std::vector<uint8_t*> d_given;
    LinContrib* d_lin_contrib = nullptr;
    double *d_lin_g = nullptr, *d_lin_c = nullptr;
    int* d_outputs = nullptr;
    bool params_set = false;
    bool setup_valid = false;
    double* d_x0 = nullptr;
    long long x0_stride = 0;
    bool have_x0 = false;
    bool timing = false;
    cb_pref last_temp{}, last_gmin{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // the eval kernels of different device models of one round are independent: models 1.. run on side streams,
    // forked from / joined into `stream` with events (a lane's launches are often a partial wave)
    cudaStream_t mstream[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool fork_models = true;
    cudaStream_t ustream[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // mixed rounds: up to 8 eval kernels in flight
    cudaEvent_t uev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

    template <class T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(1, count) * sizeof(T));
        if (e != cudaSuccess) return fail(CB_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        allocs.push_back(p);
        *out = (T*)p;
        return CB_OK;
    }
    template <class T>
    int upload(T** out, const std::vector<T>& v) {
        int rc = alloc(out, v.size());
        if (rc != CB_OK) return rc;
        if (!v.empty()) CUDA_TRY(cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        return CB_OK;
    }
};

// Level schedule of the static-pivot LU for k_lu (see kernels.cuh).  Pivot k is in level
// 1 + max(level of every earlier pivot m that updates row k or column k); the ops of one level only
// conflict through accumulation into a common destination, and those stay on one warp in pivot order.
struct LuSchedule {
    std::vector<int4> ops;
    std::vector<int> op_ptr, piv, piv_ptr, brow, brow_ptr;
    int nlev = 0, nblev = 0;
};

static void build_lu_schedule(const cb::Symbolic& S, int W, LuSchedule& out) {
    const int N = S.N, nnz = S.nnz_lu;
    std::vector<int> level(N, 0), blevel(N, 0);
    for (int m = 0; m < N; m++) {
        for (int li = S.l_ptr[m]; li < S.l_ptr[m + 1]; li++) level[S.l_row[li]] = std::max(level[S.l_row[li]], level[m] + 1);
        for (int uj = S.u_ptr[m]; uj < S.u_ptr[m + 1]; uj++) level[S.u_col[uj]] = std::max(level[S.u_col[uj]], level[m] + 1);
    }
    for (int k = N - 1; k >= 0; k--)
        for (int uj = S.u_ptr[k]; uj < S.u_ptr[k + 1]; uj++) blevel[k] = std::max(blevel[k], blevel[S.u_col[uj]] + 1);
    out.nlev = N ? *std::max_element(level.begin(), level.end()) + 1 : 0;
    out.nblev = N ? *std::max_element(blevel.begin(), blevel.end()) + 1 : 0;
    out.op_ptr.assign(1, 0); out.piv_ptr.assign(1, 0); out.brow_ptr.assign(1, 0);
    for (int lv = 0; lv < out.nlev; lv++) {
        // ops of this level grouped by destination
        std::map<int, std::vector<int4>> by_dst;
        std::vector<int> pivots;
        for (int k = 0; k < N; k++) {
            if (level[k] != lv) continue;
            pivots.push_back(k);
            const int lp = S.l_ptr[k], nL = S.l_ptr[k + 1] - lp, up = S.u_ptr[k], nU = S.u_ptr[k + 1] - up, pp = S.pair_ptr[k];
            for (int li = 0; li < nL; li++) {
                const int lpos = S.l_pos[lp + li];
                for (int uj = 0; uj < nU; uj++) {
                    const int dst = S.pair_dst[pp + li * nU + uj];
                    by_dst[dst].push_back(make_int4(lpos, S.u_pos[up + uj], dst, S.diag_pos[k]));
                }
                const int dst = nnz + S.l_row[lp + li];
                by_dst[dst].push_back(make_int4(lpos, nnz + k, dst, S.diag_pos[k]));
            }
        }
        std::vector<std::pair<int, int>> groups;   // (size, dst), largest first onto the least loaded warp
        for (auto& kv : by_dst) groups.push_back({(int)kv.second.size(), kv.first});
        std::sort(groups.begin(), groups.end(), [](auto& x, auto& y) { return x.first != y.first ? x.first > y.first : x.second < y.second; });
        std::vector<std::vector<int4>> per(W);
        for (auto& g : groups) {
            int best = 0;
            for (int w = 1; w < W; w++) if (per[w].size() < per[best].size()) best = w;
            per[best].insert(per[best].end(), by_dst[g.second].begin(), by_dst[g.second].end());   // ascending pivot order
        }
        for (int w = 0; w < W; w++) {
            out.ops.insert(out.ops.end(), per[w].begin(), per[w].end());
            out.op_ptr.push_back((int)out.ops.size());
            for (size_t q = w; q < pivots.size(); q += W) out.piv.push_back(S.diag_pos[pivots[q]]);
            out.piv_ptr.push_back((int)out.piv.size());
        }
    }
    for (int lv = 0; lv < out.nblev; lv++) {
        std::vector<int> rows;
        for (int k = 0; k < N; k++) if (blevel[k] == lv) rows.push_back(k);
        std::sort(rows.begin(), rows.end(), [&](int x, int y) {
            const int nx = S.u_ptr[x + 1] - S.u_ptr[x], ny = S.u_ptr[y + 1] - S.u_ptr[y];
            return nx != ny ? nx > ny : x < y;
        });
        std::vector<std::vector<int>> per(W);
        std::vector<int> load(W, 0);
        for (int k : rows) {
            int best = 0;
            for (int w = 1; w < W; w++) if (load[w] < load[best]) best = w;
            per[best].push_back(k);
            load[best] += 1 + S.u_ptr[k + 1] - S.u_ptr[k];
        }
        for (int w = 0; w < W; w++) {
            out.brow.insert(out.brow.end(), per[w].begin(), per[w].end());
            out.brow_ptr.push_back((int)out.brow.size());
        }
    }
}

// Forward substitution with stored factors (value-only rounds): op (l_rk, rhs_k, rhs_r, diag_k) runs at the level
// of its source k = 1 + max level of the rows that update rhs_k; ops into one destination stay on one warp in
// ascending pivot order, so the result is bit-identical to the fused sweep of the full factorisation.
static void build_fwd_schedule(const cb::Symbolic& S, int W, std::vector<int4>& ops, std::vector<int>& op_ptr, int& nlev) {
    const int N = S.N, nnz = S.nnz_lu;
    std::vector<int> level(N, 0);
    for (int k = 0; k < N; k++)
        for (int li = S.l_ptr[k]; li < S.l_ptr[k + 1]; li++) level[S.l_row[li]] = std::max(level[S.l_row[li]], level[k] + 1);
    nlev = N ? *std::max_element(level.begin(), level.end()) + 1 : 0;
    ops.clear();
    op_ptr.assign(1, 0);
    for (int lv = 0; lv < nlev; lv++) {
        std::map<int, std::vector<int4>> by_dst;
        for (int k = 0; k < N; k++) {
            if (level[k] != lv) continue;
            for (int li = S.l_ptr[k]; li < S.l_ptr[k + 1]; li++) {
                const int dst = nnz + S.l_row[li];
                by_dst[dst].push_back(make_int4(S.l_pos[li], nnz + k, dst, S.diag_pos[k]));
            }
        }
        std::vector<std::pair<int, int>> groups;
        for (auto& kv : by_dst) groups.push_back({(int)kv.second.size(), kv.first});
        std::sort(groups.begin(), groups.end(), [](auto& x, auto& y) { return x.first != y.first ? x.first > y.first : x.second < y.second; });
        std::vector<std::vector<int4>> per(W);
        for (auto& g : groups) {
            int best = 0;
            for (int w = 1; w < W; w++) if (per[w].size() < per[best].size()) best = w;
            per[best].insert(per[best].end(), by_dst[g.second].begin(), by_dst[g.second].end());
        }
        for (int w = 0; w < W; w++) {
            ops.insert(ops.end(), per[w].begin(), per[w].end());
            op_ptr.push_back((int)ops.size());
        }
    }
}

// Shared-memory carve-out (percent of the SM's maximum) of a generated eval kernel: what its resident CTAs need
// (meta[3] CTAs of meta[1] bytes of cache ring + static + the driver's 1 KB), so that the rest of the 256 KB is L1 --
// the eval kernels spill to local memory and are bound by those reloads (ncu: long-scoreboard stalls).  The driver's
// default keeps the maximum shared-memory configuration (233 KB, ~28 KB of L1) whatever the kernel uses.
// CB_EVAL_CARVEOUT=<percent> overrides, -1 = driver default.
static int eval_carveout(const int* meta) {
    if (const char* cv = std::getenv("CB_EVAL_CARVEOUT")) return std::atoi(cv);
    const long long need = (long long)std::max(1, meta[3]) * ((long long)meta[1] + 2048);
    return (int)std::min<long long>(100, (need * 100 + 233471) / 233472);
}

static int plan_create1(cb_circuit* c, int64_t n_inst, int device_id, cb_plan** out) {
    if (!c || !out || n_inst <= 0) return fail(CB_ERR_INVALID, "bad argument");
    if (!c->compiled) return fail(CB_ERR_STATE, "circuit not compiled");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(CB_ERR_NO_DEVICE, "no CUDA device available; this engine has no CPU fallback");
    if (device_id < 0 || device_id >= ndev) return fail(CB_ERR_INVALID, "device id out of range");
    CUDA_TRY(cudaSetDevice(device_id));
    // an error return below releases the streams, events and device allocations made so far
    struct PlanDeleter { void operator()(cb_plan* q) const { cb_plan_destroy(q); } };
    std::unique_ptr<cb_plan, PlanDeleter> p(new cb_plan());
    p->c = c; p->B = n_inst; p->device = device_id;
    const long long B = n_inst;
    const int N = c->N;
    const cb::Symbolic& S = c->sym;
    CUDA_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&p->ev0));
    CUDA_TRY(cudaEventCreate(&p->ev1));
    CUDA_TRY(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    for (size_t m = 1; m < c->models.size() && m < 8; m++) {
        CUDA_TRY(cudaStreamCreateWithFlags(&p->mstream[m], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&p->ev_join[m], cudaEventDisableTiming));
    }
    for (size_t k = 1; k < 2 * c->models.size() && k < 8; k++) {
        CUDA_TRY(cudaStreamCreateWithFlags(&p->ustream[k], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&p->uev[k], cudaEventDisableTiming));
    }
    if (const char* e = std::getenv("CB_EVAL_FORK")) p->fork_models = std::atoi(e) != 0;
    int rc;
#define TRY(x) do { rc = (x); if (rc != CB_OK) return rc; } while (0)
    {
        CUDA_TRY(cudaLibraryLoadData(&p->lib, c->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
        CUDA_TRY(cudaLibraryGetKernel(&p->k_solve, p->lib, "k_solve"));
        for (const ModelH& m : c->models) {
            cudaKernel_t ks, ke;
            CUDA_TRY(cudaLibraryGetKernel(&ks, p->lib, ("k_setup_" + m.name).c_str()));
            CUDA_TRY(cudaLibraryGetKernel(&ke, p->lib, ("k_eval_" + m.name).c_str()));
            {   // block size and dynamic shared memory (cache ring) the generated kernel was compiled for
                void* dmeta = nullptr;
                size_t msz = 0;
                int meta[6] = {128, 0, 0, 0, 1, 0};
                CUDA_TRY(cudaLibraryGetGlobal(&dmeta, &msz, p->lib, ("va_meta_" + m.name).c_str()));
                CUDA_TRY(cudaMemcpy(meta, dmeta, std::min(msz, sizeof meta), cudaMemcpyDeviceToHost));
                p->nuni[0].push_back(std::max(1, meta[4]));
                p->eval_threads.push_back((unsigned)meta[0]);
                p->eval_smem.push_back((size_t)meta[1]);
                if (meta[1] > 48 * 1024)
                    CUDA_TRY(cudaFuncSetAttribute((const void*)ke, cudaFuncAttributeMaxDynamicSharedMemorySize, meta[1]));
                CUDA_TRY(cudaFuncSetAttribute((const void*)ke, cudaFuncAttributePreferredSharedMemoryCarveout, eval_carveout(meta)));
            }
            p->k_setup.push_back(ks);
            p->k_eval.push_back(ke);
            cudaKernel_t ksv = nullptr, kev = nullptr;
            int metav[6] = {128, 0, 0, 0, 1, 0};
            if (cudaLibraryGetKernel(&kev, p->lib, ("k_evalv_" + m.name).c_str()) == cudaSuccess &&
                cudaLibraryGetKernel(&ksv, p->lib, ("k_setupv_" + m.name).c_str()) == cudaSuccess) {
                void* dmeta = nullptr;
                size_t msz = 0;
                CUDA_TRY(cudaLibraryGetGlobal(&dmeta, &msz, p->lib, ("va_metav_" + m.name).c_str()));
                CUDA_TRY(cudaMemcpy(metav, dmeta, std::min(msz, sizeof metav), cudaMemcpyDeviceToHost));
                if (metav[1] > 48 * 1024)
                    CUDA_TRY(cudaFuncSetAttribute((const void*)kev, cudaFuncAttributeMaxDynamicSharedMemorySize, metav[1]));
                CUDA_TRY(cudaFuncSetAttribute((const void*)kev, cudaFuncAttributePreferredSharedMemoryCarveout, eval_carveout(metav)));
            } else {
                (void)cudaGetLastError();
                ksv = kev = nullptr;
            }
            {   // noise variant (optional)
                cudaKernel_t ksn = nullptr, ken = nullptr;
                int metan[6] = {128, 0, 0, 0, 1, 0};
                if (!m.noise_pos.empty() && cudaLibraryGetKernel(&ken, p->lib, ("k_evaln_" + m.name).c_str()) == cudaSuccess &&
                    cudaLibraryGetKernel(&ksn, p->lib, ("k_setupn_" + m.name).c_str()) == cudaSuccess) {
                    void* dmeta = nullptr;
                    size_t msz = 0;
                    CUDA_TRY(cudaLibraryGetGlobal(&dmeta, &msz, p->lib, ("va_metan_" + m.name).c_str()));
                    CUDA_TRY(cudaMemcpy(metan, dmeta, std::min(msz, sizeof metan), cudaMemcpyDeviceToHost));
                    if (metan[1] > 48 * 1024)
                        CUDA_TRY(cudaFuncSetAttribute((const void*)ken, cudaFuncAttributeMaxDynamicSharedMemorySize, metan[1]));
                } else {
                    (void)cudaGetLastError();
                    ksn = ken = nullptr;
                }
                p->k_setupn.push_back(ksn);
                p->k_evaln.push_back(ken);
                p->evaln_threads.push_back((unsigned)metan[0]);
                p->evaln_smem.push_back((size_t)metan[1]);
                p->nuni[2].push_back(std::max(1, metan[4]));
            }
            p->k_setupv.push_back(ksv);
            p->k_evalv.push_back(kev);
            p->evalv_threads.push_back((unsigned)metav[0]);
            p->evalv_smem.push_back((size_t)metav[1]);
            p->cachev_off.push_back(metav[2]);   // slots per instance for now; turned into offsets below
            p->nuni[1].push_back(std::max(1, metav[4]));
        }
    }
    NArgs& a = p->na;
    a.B = B; a.N = N; a.NV = c->NV; a.nnz_lu = S.nnz_lu; a.O = (int)c->outputs.size();
    a.nwaves = (int)c->waves.size();
    // symbolic + assembly tables
    int* tmp;
#define UP(field, vec) do { TRY(p->upload(&tmp, vec)); a.field = tmp; } while (0)
    UP(diag_pos, S.diag_pos); UP(l_ptr, S.l_ptr); UP(l_pos, S.l_pos); UP(l_row, S.l_row);
    UP(u_ptr, S.u_ptr); UP(u_pos, S.u_pos); UP(pair_ptr, S.pair_ptr); UP(pair_dst, S.pair_dst);
    UP(uc_ptr, S.uc_ptr); UP(uc_pos, S.uc_pos); UP(uc_row, S.uc_row);
    UP(row_to_step, S.row_to_step); UP(col_to_step, S.col_to_step);
    UP(a_ptr, c->a_ptr); UP(a_src, c->a_src); UP(a_lin, c->a_lin);
    UP(ri_ptr, c->ri_ptr); UP(ri_src, c->ri_src); UP(rq_ptr, c->rq_ptr); UP(rq_src, c->rq_src);
    UP(rl_ptr, c->rl_ptr); UP(rl_col, c->rl_col); UP(rl_lin, c->rl_lin);
    UP(rs_ptr, c->rs_ptr); UP(rs_wave, c->rs_wave);
#undef UP
    double* dtmp;
    TRY(p->upload(&dtmp, c->a_mult)); a.a_mult = dtmp;
    TRY(p->upload(&dtmp, c->ri_mult)); a.ri_mult = dtmp;
    TRY(p->upload(&dtmp, c->rq_mult)); a.rq_mult = dtmp;
    TRY(p->upload(&dtmp, c->rs_coef)); a.rs_coef = dtmp;
    uint8_t* btmp;
    TRY(p->upload(&btmp, c->a_diag)); a.a_diag = btmp;
    TRY(p->upload(&btmp, c->lte_mask)); a.lte_mask = btmp;
    TRY(p->upload(&p->d_outputs, c->outputs)); a.outputs = p->d_outputs;
    // waves
    {
        std::vector<WaveDev> wd(c->waves.size());
        for (size_t i = 0; i < c->waves.size(); i++) {
            const WaveH& w = c->waves[i];
            WaveDev d{};
            d.kind = w.kind; d.has_dc = w.has_dc;
            d.dc = {w.dc.value, w.dc.col, 0};
            d.npts = (int)w.t.size();
            if (!w.t.empty()) {
                double* dt_; Pref* dy_;
                TRY(p->upload(&dt_, w.t));
                std::vector<Pref> ys(w.y.size());
                for (size_t k = 0; k < ys.size(); k++) ys[k] = {w.y[k].value, w.y[k].col, 0};
                TRY(p->upload(&dy_, ys));
                d.t = dt_; d.y = dy_;
            }
            for (int k = 0; k < 7; k++) d.v[k] = {w.v[k].value, w.v[k].col, 0};
            wd[i] = d;
        }
        WaveDev* dw;
        TRY(p->upload(&dw, wd));
        a.waves = dw;
    }
    // linear stamp values
    TRY(p->upload(&p->d_lin_contrib, c->lin_contrib));
    {
        const long long Bl = c->lin_swept ? B : 1;
        TRY(p->alloc(&p->d_lin_g, (size_t)std::max(1, c->nlin) * Bl));
        TRY(p->alloc(&p->d_lin_c, (size_t)std::max(1, c->nlin) * Bl));
        a.lin_g = p->d_lin_g; a.lin_c = p->d_lin_c;
        a.lin_inst_stride = c->lin_swept ? 1 : 0;
        a.lin_ent_stride = Bl;
    }
    // state
    TRY(p->alloc(&p->d_params, (size_t)std::max(1, c->P) * B));
    a.params = p->d_params;
    double** vecs[] = {&a.X, &a.XN, &a.X1, &a.X2, &a.XP, &a.QN, &a.Q1, &a.QD, &a.BETA};
    for (double** v : vecs) TRY(p->alloc(v, (size_t)N * B));
    TRY(p->alloc(&a.alpha, (size_t)B));
    TRY(p->alloc(&a.dst, (size_t)DS_COUNT * B));
    TRY(p->alloc(&a.ist, (size_t)IS_COUNT * B));
    TRY(p->alloc(&a.active, (size_t)B));
    p->Bpad = (B + 127) / 128 * 128;   // the per-instance caches are laid out in blocks of 128 points (va_prelude.h)
    TRY(p->alloc(&p->d_cache, (size_t)std::max<long long>(1, c->total_cache) * p->Bpad));
    {   // value-only variants: usable when every model with instances has one
        bool all = !c->insts.empty();
        long long total = 0;
        for (size_t m = 0; m < c->models.size(); m++) {
            const long long per = pad4((int)p->cachev_off[m]);
            p->cachev_off[m] = total;
            if (c->model_insts[m].empty()) continue;
            if (!p->k_evalv[m]) all = false;
            total += (long long)c->model_insts[m].size() * per;
        }
        p->have_v = all;
        p->cachev_slots = total;
    }
    TRY(p->alloc(&p->d_dev_out, (size_t)std::max<long long>(1, c->total_out) * B));
    a.dev_out = p->d_dev_out;
    TRY(p->alloc(&p->d_xout, (size_t)std::max<size_t>(1, c->outputs.size()) * B));
    TRY(p->alloc(&p->d_done, 1));
    a.done_count = p->d_done;
    CUDA_TRY(cudaMallocHost((void**)&p->h_done, 8 * sizeof(int)));
    // per-model tables
    for (size_t m = 0; m < c->models.size(); m++) {
        const ModelH& M = c->models[m];
        std::vector<int> term, pcol;
        std::vector<double> pval;
        std::vector<uint8_t> giv;
        for (int gi : c->model_insts[m]) {
            const VaInstH& v = c->insts[gi];
            term.insert(term.end(), v.term.begin(), v.term.end());
            for (int k = 0; k < M.nparam; k++) {
                pval.push_back(v.par[k].value);
                pcol.push_back(v.given[k] ? v.par[k].col : -1);
                giv.push_back(v.given[k]);
            }
        }
        int* dt_; int* dc_; double* dv_; uint8_t* dg_;
        TRY(p->upload(&dt_, term)); TRY(p->upload(&dc_, pcol)); TRY(p->upload(&dv_, pval)); TRY(p->upload(&dg_, giv));
        p->d_term.push_back(dt_); p->d_par_col.push_back(dc_); p->d_par_val.push_back(dv_); p->d_given.push_back(dg_);
    }
    for (int v = 0; v < 3; v++)
        for (size_t m = 0; m < c->models.size(); m++) {
            double* du = nullptr;
            TRY(p->alloc(&du, (size_t)p->nuni[v][m] * B));
            p->d_uni[v].push_back(du);
        }
    // solve kernel: shared-memory batched LU (k_lu) when the factors of LU_PTS points fit one SM, else the generated
    // straight-line k_solve (no value-only rounds then)
    {
        int max_smem = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device_id));
        const char* env = std::getenv("CB_NEWTON");
        p->lu_smem = (size_t)(S.nnz_lu + 2 * N) * LU_PTS * sizeof(double);
        const size_t lu_static = (size_t)(2 * sizeof(double) + sizeof(int)) * LU_PTS + 2048;   // reduction slots + the driver's reserve
        p->lu = !(env && std::string(env) == "gen") && p->lu_smem + lu_static <= (size_t)max_smem;
        if (!p->lu) p->have_v = false;
        if (p->lu) {
            LuSchedule sch;
            build_lu_schedule(S, LU_W, sch);
            LArgs& la = p->la;
            la.nlev = sch.nlev; la.nblev = sch.nblev;
            std::vector<int4> sops;
            std::vector<int> sop_ptr;
            int nslev = 0;
            build_fwd_schedule(S, LU_W, sops, sop_ptr, nslev);
            la.nslev = nslev;
            // the table blob (kernels.cuh LuTabs): 16-bit entries, every table 16-byte aligned
            std::vector<unsigned char>& blob = p->lu_blob;
            bool blob_ok = true;
            auto add16 = [&](const std::vector<int>& v) {
                const int off = (int)blob.size();
                for (int x : v) {
                    if (x < 0 || x > 65535) blob_ok = false;
                    const unsigned short u = (unsigned short)x;
                    blob.push_back((unsigned char)(u & 0xff)); blob.push_back((unsigned char)(u >> 8));
                }
                while (blob.size() % 16) blob.push_back(0);
                return off;
            };
            auto add_ops = [&](const std::vector<int4>& v) {
                std::vector<int> flat;
                for (const int4& o : v) { flat.push_back(o.x); flat.push_back(o.y); flat.push_back(o.z); flat.push_back(o.w); }
                return add16(flat);
            };
            LuTabs& t = la.t;
            t.op_ptr = add16(sch.op_ptr); t.piv_ptr = add16(sch.piv_ptr); t.piv = add16(sch.piv); t.ops = add_ops(sch.ops);
            t.brow_ptr = add16(sch.brow_ptr); t.brow = add16(sch.brow);
            t.u_ptr = add16(S.u_ptr); t.u_pos = add16(S.u_pos); t.u_col = add16(S.u_col); t.diag_pos = add16(S.diag_pos);
            t.sop_ptr = add16(sop_ptr); t.sops = add_ops(sops);
            {
                std::vector<int> al(S.nnz_lu);
                for (int e = 0; e < S.nnz_lu; e++) {
                    if (c->a_lin[e] + 1 >= 0x8000) blob_ok = false;
                    al[e] = ((c->a_lin[e] + 1) & 0x7fff) | (c->a_diag[e] ? 0x8000 : 0);
                }
                t.a_lin = add16(al);
            }
            t.rl_ptr = add16(c->rl_ptr); t.rl_lin = add16(c->rl_lin); t.rl_col = add16(c->rl_col);
            t.rs_ptr = add16(c->rs_ptr); t.rs_wave = add16(c->rs_wave);
            t.row_to_step = add16(S.row_to_step); t.col_to_step = add16(S.col_to_step);
            {   // elimination step of the row / column of every LU entry (repair pass: scatter into the dense system)
                std::vector<int> er(S.nnz_lu, 0), ec(S.nnz_lu, 0);
                for (int k = 0; k < N; k++) {
                    er[S.diag_pos[k]] = k; ec[S.diag_pos[k]] = k;
                    for (int li = S.l_ptr[k]; li < S.l_ptr[k + 1]; li++) { er[S.l_pos[li]] = S.l_row[li]; ec[S.l_pos[li]] = k; }
                    for (int uj = S.u_ptr[k]; uj < S.u_ptr[k + 1]; uj++) { er[S.u_pos[uj]] = k; ec[S.u_pos[uj]] = S.u_col[uj]; }
                }
                std::vector<unsigned short> er16(er.begin(), er.end()), ec16(ec.begin(), ec.end());
                unsigned short* d16;
                TRY(p->upload(&d16, er16)); la.e_row = d16;
                TRY(p->upload(&d16, ec16)); la.e_col = d16;
            }
            if (std::getenv("CB_DEBUG"))
                std::fprintf(stderr, "k_lu schedule: N=%d nnz=%d levels=%d back-levels=%d ops=%zu fwd-levels=%d fwd-ops=%zu smem=%zu\n", N,
                             S.nnz_lu, sch.nlev, sch.nblev, sch.ops.size(), nslev, sops.size(), p->lu_smem);
            // gather items, grouped by destination and balanced over the warps; the value-only list holds the
            // residual and charge rows only
            // items are (dev_out row | multiplier code << 20, destination | second position << 16); the distinct multipliers
            // (+-1, +-m of the instances, source coefficients) are a small table of doubles in the blob
            std::vector<double> mtab;
            auto mcode = [&](double m) {
                for (size_t k = 0; k < mtab.size(); k++) if (std::memcmp(&mtab[k], &m, sizeof m) == 0) return (int)k;
                mtab.push_back(m);
                return (int)mtab.size() - 1;
            };
            auto item = [&](int src, int dst, int aux, double m) {
                const int code = mcode(m);
                if (src < 0 || src >= (1 << 20) || code >= (1 << 12) || dst < 0 || dst > 65535 || aux < 0 || aux > 65535) blob_ok = false;
                return make_int2((int)((unsigned)src | ((unsigned)code << 20)), (int)((unsigned)dst | ((unsigned)aux << 16)));
            };
            auto balance = [&](std::map<int, std::vector<int2>>& by_dst, std::vector<int2>& items, std::vector<int>& iptr) {
                std::vector<std::pair<int, int>> groups;   // (size, dst), largest first onto the least loaded warp
                for (auto& kv : by_dst) groups.push_back({(int)kv.second.size(), kv.first});
                std::sort(groups.begin(), groups.end(), [](auto& x, auto& y) { return x.first != y.first ? x.first > y.first : x.second < y.second; });
                std::vector<std::vector<int2>> per(LU_W);
                for (auto& g : groups) {
                    int best = 0;
                    for (int w = 1; w < LU_W; w++) if (per[w].size() < per[best].size()) best = w;
                    per[best].insert(per[best].end(), by_dst[g.second].begin(), by_dst[g.second].end());
                }
                items.clear(); iptr.assign(1, 0);
                for (int w = 0; w < LU_W; w++) { items.insert(items.end(), per[w].begin(), per[w].end()); iptr.push_back((int)items.size()); }
            };
            int2* ditems;
            // gather items, grouped by destination and balanced over the warps; the value-only list holds the
            // residual and charge rows only
            for (int pass = 0; pass < 2; pass++) {
                std::map<int, std::vector<int2>> by_dst;
                if (pass == 0)
                    for (int e = 0; e < S.nnz_lu; e++)
                        for (int q = c->a_ptr[e]; q < c->a_ptr[e + 1]; q++) by_dst[e].push_back(item(c->a_src[q], e, 0, c->a_mult[q]));
                for (int i = 0; i < N; i++) {
                    for (int q = c->ri_ptr[i]; q < c->ri_ptr[i + 1]; q++)
                        by_dst[S.nnz_lu + S.row_to_step[i]].push_back(item(c->ri_src[q], S.nnz_lu + S.row_to_step[i], 0, -c->ri_mult[q]));
                    for (int q = c->rq_ptr[i]; q < c->rq_ptr[i + 1]; q++)
                        by_dst[S.nnz_lu + N + i].push_back(item(c->rq_src[q], S.nnz_lu + N + i, 0, c->rq_mult[q]));
                }
                std::vector<int2> items;
                std::vector<int> iptr;
                balance(by_dst, items, iptr);
                TRY(p->upload(&ditems, items));
                if (pass == 0) { la.items = ditems; t.item_ptr = add16(iptr); } else { la.sitems = ditems; t.sitem_ptr = add16(iptr); }
            }
            {   // charge-update items, all items of one row on one warp
                std::map<int, std::vector<int2>> by_row;
                for (size_t q = 0; q < c->cq_row.size(); q++)
                    by_row[c->cq_row[q]].push_back(item(c->cq_src[q], S.nnz_lu + N + c->cq_row[q], S.nnz_lu + S.col_to_step[c->cq_col[q]], c->cq_mult[q]));
                std::vector<int2> items;
                std::vector<int> iptr;
                balance(by_row, items, iptr);
                TRY(p->upload(&ditems, items)); la.citems = ditems;
                t.citem_ptr = add16(iptr);
            }
            {
                t.mtab = (int)blob.size();
                const unsigned char* mb = (const unsigned char*)mtab.data();
                blob.insert(blob.end(), mb, mb + mtab.size() * sizeof(double));
                while (blob.size() % 16) blob.push_back(0);
            }
            t.bytes = (int)blob.size();
            if (!blob_ok) { p->lu = false; p->have_v = false; }   // an index beyond 16 bits (never at sizes whose matrices fit one SM): generated k_solve
            else {
                unsigned char* dblob;
                TRY(p->upload(&dblob, blob)); la.tab = dblob;
                // stage the blob in shared memory when it fits behind the matrices
                p->lu_staged = p->lu_smem + blob.size() + lu_static <= (size_t)max_smem && !std::getenv("CB_LU_NOSTAGE");
                if (p->lu_staged) p->lu_smem += blob.size();
                if (std::getenv("CB_DEBUG")) std::fprintf(stderr, "k_lu tables: %zu bytes, staged=%d\n", blob.size(), (int)p->lu_staged);
                if (p->lu_staged) {
                    CUDA_TRY(cudaFuncSetAttribute(k_lu<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->lu_smem));
                    CUDA_TRY(cudaFuncSetAttribute(k_lu<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->lu_smem));
                    CUDA_TRY(cudaFuncSetAttribute(k_lu<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->lu_smem));
                } else {
                    CUDA_TRY(cudaFuncSetAttribute(k_lu<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->lu_smem));
                    CUDA_TRY(cudaFuncSetAttribute(k_lu<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->lu_smem));
                    CUDA_TRY(cudaFuncSetAttribute(k_lu<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->lu_smem));
                }
            }
            la.LUF = nullptr;
        }
        TRY(p->alloc(&p->d_DX, (size_t)N * B));
        TRY(p->alloc(&p->d_QK, (size_t)N * B));
        TRY(p->alloc(&p->d_RMAX, (size_t)B));
        TRY(p->alloc(&p->d_DVMAX, (size_t)B));
        TRY(p->alloc(&p->d_BAD, (size_t)B));
        TRY(p->alloc(&p->d_WV, (size_t)std::max(1, a.nwaves) * B));
        TRY(p->alloc(&p->d_dc_count, 1));
        TRY(p->alloc(&p->d_lists, (size_t)CB_NBUF * 3 * B));   // rotating buffers of [full | value-only | idle] point lists
        TRY(p->alloc(&p->d_cnt, CB_NBUF * 4));
        {
            cudaDeviceProp prop;
            CUDA_TRY(cudaGetDeviceProperties(&prop, device_id));
            p->num_sms = std::max(1, prop.multiProcessorCount);
        }
        if (p->lu) {   // one dense system per k_lu CTA (the grid never exceeds the SM count), no more than the batch has groups
            const long long ctas = std::min<long long>((B + LU_PTS - 1) / LU_PTS + 1, (long long)p->num_sms);
            TRY(p->alloc(&p->d_pp_scratch, (size_t)ctas * N * (N + 1)));
        }
        a.scratch = nullptr;
        a.sm_stride = 0;
    }
#undef TRY
    *out = p.release();
    return CB_OK;
}

// `pitch` = row length (in points) of the caller's [k][B_total] arrays; this lane owns columns [0, p->B) of `params`
static int set_params1(cb_plan* p, const double* params, long long pitch, cudaMemcpyKind kind) {
    if (!p) return fail(CB_ERR_INVALID, "null plan");
    CUDA_TRY(cudaSetDevice(p->device));
    if (p->c->P > 0) {
        if (!params) return fail(CB_ERR_INVALID, "params is null but the circuit has swept parameters");
        CUDA_TRY(cudaMemcpy2DAsync(p->d_params, (size_t)p->B * sizeof(double), params, (size_t)pitch * sizeof(double),
                                   (size_t)p->B * sizeof(double), (size_t)p->c->P, kind, p->stream));
    }
    p->params_set = true;
    p->setup_valid = false;
    p->setupn_valid = false;
    return CB_OK;
}

static int set_x01(cb_plan* p, const double* x0, int per_point, long long pitch) {
    if (!p) return fail(CB_ERR_INVALID, "null plan");
    CUDA_TRY(cudaSetDevice(p->device));
    if (!x0) { p->have_x0 = false; return CB_OK; }
    if (!p->d_x0) {
        int rc = p->alloc(&p->d_x0, (size_t)p->c->N * p->B);
        if (rc != CB_OK) return rc;
    }
    if (per_point)
        CUDA_TRY(cudaMemcpy2D(p->d_x0, (size_t)p->B * sizeof(double), x0, (size_t)pitch * sizeof(double),
                              (size_t)p->B * sizeof(double), (size_t)p->c->N, cudaMemcpyHostToDevice));
    else
        CUDA_TRY(cudaMemcpy(p->d_x0, x0, (size_t)p->c->N * sizeof(double), cudaMemcpyHostToDevice));
    p->x0_stride = per_point ? p->B : 0;
    p->have_x0 = true;
    return CB_OK;
}

extern "C" int cb_plan_device_params(cb_plan* p, double** d_params) {
    if (!p || !d_params) return fail(CB_ERR_INVALID, "null argument");
    if (p->multi_device) return fail(CB_ERR_INVALID, "device-resident entry points need a single-device plan");
    if (!p->lanes.empty()) {   // one [P][B] buffer for the caller; scattered to the lanes before the next solve
        CUDA_TRY(cudaSetDevice(p->device));
        if (!p->d_params_all)
            CUDA_TRY(cudaMalloc((void**)&p->d_params_all, (size_t)std::max(1, p->c->P) * p->B * sizeof(double)));
        *d_params = p->d_params_all;
        p->params_all_dirty = true;
        p->params_set = true;
        return CB_OK;
    }
    *d_params = p->d_params;
    p->params_set = true;   // caller writes the device buffer directly
    p->setup_valid = false;
    p->setupn_valid = false;
    return CB_OK;
}

// layout must match struct VaArgs in va_prelude.h
struct VaArgsH {
    long long B; const double* x; const double* alpha; const int* list; const double* cache; double* out;
    const int* term; const double* params; const double* par_val; const int* par_col; const uint8_t* given;
    double temp_val; double gmin_val; int temp_col; int gmin_col; const int* count;
    double* uni; int uni_per_inst; int want; const int* active;
};

// which point list a device-evaluation launch runs over: this round's full-iteration or value-only points
static void fill_va_args(cb_plan* p, size_t m, const cb_options* opt, void* out_args, bool value_only = false, int parity = 0) {
    VaArgsH* a = (VaArgsH*)out_args;
    const cb_circuit* c = p->c;
    a->B = p->B; a->x = p->na.X; a->alpha = p->na.alpha;
    a->list = p->d_lists + ((size_t)parity * 3 + (value_only ? 1 : 0)) * p->B;   // parity = list buffer of the round
    a->count = p->d_cnt + parity * 4 + (value_only ? 1 : 0);
    a->cache = value_only ? p->d_cachev + (size_t)p->cachev_off[m] * p->Bpad : p->d_cache + (size_t)c->cache_off[m] * p->Bpad;
    a->out = p->d_dev_out + (size_t)c->out_off[m] * p->B;
    a->term = p->d_term[m]; a->params = p->d_params; a->par_val = p->d_par_val[m]; a->par_col = p->d_par_col[m];
    a->given = p->d_given[m];
    a->temp_val = opt->temp.value; a->gmin_val = opt->gmin.value;
    a->temp_col = opt->temp.col; a->gmin_col = opt->gmin.col;
    a->uni = p->d_uni[value_only ? 1 : 0][m];
    a->uni_per_inst = (opt->temp.col >= 0 || opt->gmin.col >= 0) ? 1 : 0;
    a->want = 0; a->active = nullptr;   // list mode (solve() switches its launches to blocked mode for small batches)
}

// (the repairable variant, k_lu<., ., true>, when the arguments carry a repair scratch: cb_options.pivot_repair)
static inline void launch_lu(cb_plan* p, unsigned grid, cudaStream_t st, const LArgs& la, bool fused = false) {
    const bool rep = la.pp_scratch != nullptr && !fused;
    if (p->lu_staged) {
        if (fused) k_lu<true, true, false><<<grid, LU_PTS * LU_W, p->lu_smem, st>>>(la);
        else if (rep) k_lu<true, false, true><<<grid, LU_PTS * LU_W, p->lu_smem, st>>>(la);
        else k_lu<true, false, false><<<grid, LU_PTS * LU_W, p->lu_smem, st>>>(la);
    } else {
        if (fused) k_lu<false, true, false><<<grid, LU_PTS * LU_W, p->lu_smem, st>>>(la);
        else if (rep) k_lu<false, false, true><<<grid, LU_PTS * LU_W, p->lu_smem, st>>>(la);
        else k_lu<false, false, false><<<grid, LU_PTS * LU_W, p->lu_smem, st>>>(la);
    }
}

static int run_setup(cb_plan* p, const cb_options* opt) {
    NvtxRange nvtx_("cb:setup (linear stamps, device-model caches)");
    const cb_circuit* c = p->c;
    if (p->setup_valid && std::memcmp(&p->last_temp, &opt->temp, sizeof(cb_pref)) == 0 &&
        std::memcmp(&p->last_gmin, &opt->gmin, sizeof(cb_pref)) == 0)
        return CB_OK;
    const long long B = p->B;
    // linear stamp values
    {
        const long long Bl = c->lin_swept ? B : 1;
        const int threads = 128;
        k_lin_setup<<<(unsigned)((Bl + threads - 1) / threads), threads, 0, p->stream>>>(
            Bl, B, c->nlin, (int)c->lin_contrib.size(), p->d_lin_contrib, p->d_params, p->d_lin_g, p->d_lin_c);
        CUDA_TRY(cudaGetLastError());
    }
    for (size_t m = 0; m < c->models.size(); m++) {
        if (c->model_insts[m].empty()) continue;
        char args[256];
        fill_va_args(p, m, opt, args);
        void* kargs[] = {args};
        dim3 grid((unsigned)((B + 127) / 128), (unsigned)c->model_insts[m].size());
        CUDA_TRY(cudaLaunchKernel((const void*)p->k_setup[m], grid, dim3(128), kargs, 0, p->stream));
    }
    p->setup_valid = true;
    p->setupv_valid = false;
    p->setupn_valid = false;   // the noise variant's cache depends on temp / gmin as well
    p->last_temp = opt->temp; p->last_gmin = opt->gmin;
    return CB_OK;
}

// value-only variants: their own bias-independent cache, filled on first use after a parameter change
static int run_setupv(cb_plan* p, const cb_options* opt) {
    NvtxRange nvtx_("cb:setup value-only caches");
    const cb_circuit* c = p->c;
    if (p->setupv_valid) return CB_OK;
    for (size_t m = 0; m < c->models.size(); m++) {
        if (c->model_insts[m].empty()) continue;
        char args[256];
        fill_va_args(p, m, opt, args, true);
        void* kargs[] = {args};
        dim3 grid((unsigned)((p->B + 127) / 128), (unsigned)c->model_insts[m].size());
        CUDA_TRY(cudaLaunchKernel((const void*)p->k_setupv[m], grid, dim3(128), kargs, 0, p->stream));
    }
    p->setupv_valid = true;
    return CB_OK;
}

static int collect_breakpoints(const cb_circuit* c, double t0, double t1, std::vector<double>& out) {
    std::vector<double> bp;
    for (const WaveH& w : c->waves) {
        if (w.kind == CB_W_PULSE && w.v[6].value > 0.0 && std::isfinite(w.v[6].value) && t1 / w.v[6].value > 4e6)
            return fail(CB_ERR_INVALID, "PULSE period gives more than 4e6 periods (16e6 breakpoints) inside the transient span");
        if (w.kind == CB_W_PWL) for (double t : w.t) bp.push_back(t);
        else if (w.kind == CB_W_PULSE) {
            const double td = w.v[2].value, tr = w.v[3].value, tf = w.v[4].value, pw = w.v[5].value, per = w.v[6].value;
            const double ts[4] = {td, td + tr, td + tr + pw, td + tr + pw + tf};
            for (int k = 0; k < 4; k++) {
                if (!std::isfinite(ts[k])) continue;
                if (!(per > 0.0) || std::isinf(per)) bp.push_back(ts[k]);
                else for (double base = 0.0; base + ts[k] <= t1; base += per) bp.push_back(base + ts[k]);
            }
        } else if (w.kind == CB_W_SIN) bp.push_back(w.v[3].value);
    }
    bp.push_back(t1);
    std::sort(bp.begin(), bp.end());
    out.clear();
    const double tiny = 1e-12 * std::max(std::fabs(t1), std::fabs(t1 - t0));
    for (double b : bp) {
        if (b <= t0 + tiny || b > t1 + tiny) continue;
        if (!out.empty() && b - out.back() <= tiny) continue;
        out.push_back(b);
    }
    return CB_OK;
}

static int solve(cb_plan* p, const cb_options* opt, bool dc_only, double t0, double t1, const double* saveat,
                 int64_t nsave, cb_stats* stats) {
    NvtxRange nvtx_(dc_only ? "cb:solve dc" : "cb:solve tran");
    cb_circuit* c = p->c;
    if (!p->params_set && c->P > 0) return fail(CB_ERR_STATE, "cb_plan_set_params was not called");
    CUDA_TRY(cudaSetDevice(p->device));
    const long long B = p->B;
    NArgs& a = p->na;
    for (int k = 0; k < 2; k++) {
        const cb_pref& pr = k == 0 ? opt->temp : opt->gmin;
        if (pr.col >= c->P) return fail(CB_ERR_INVALID, "temp/gmin column out of range");
    }
    Opts& o = a.o;
    o.reltol = opt->reltol; o.vabstol = opt->vabstol; o.iabstol = opt->iabstol;
    o.nr_reltol = opt->nr_reltol; o.nr_vabstol = opt->nr_vabstol; o.nr_iabstol = opt->nr_iabstol;
    o.dc_abstol = opt->dc_abstol;
    // the Newton voltage-step limit only exists for nonlinear (Verilog-A) devices
    {   // the voltage-step limit is for nonlinear devices only
        bool nonlinear = false;
        for (const VaInstH& vi : c->insts) nonlinear = nonlinear || !c->models[vi.model].linear;
        o.dv_max = nonlinear ? opt->dv_max : 1e300;
    }
    o.dt = opt->dt; o.dt_min = opt->dt_min; o.t0 = t0; o.t1 = t1;
    o.span = t1 - t0;
    o.teps = 1e-12 * std::max(std::fabs(t1), o.span);
    o.dt_max = opt->dt_max > 0 ? opt->dt_max : o.span / 50.0;
    o.max_newton_dc = opt->max_newton_dc; o.max_newton_tran = opt->max_newton_tran;
    o.method = opt->method; o.fixed_step = opt->fixed_step; o.gmin_steps = opt->gmin_steps;
    o.skip_dc = opt->skip_dc; o.dc_only = dc_only ? 1 : 0;
    o.source_steps = std::max(0, opt->source_steps);
    o.reinit = 0;   // t0 re-initialisation only when a source's DC value may differ from its transient value at t0
    if (opt->t0_reinit && !dc_only)
        for (const WaveH& w : c->waves) if (w.has_dc && w.kind != CB_W_DC) o.reinit = 1;
    o.rate_test = p->lu ? opt->nr_rate_test : 0;   // needs the charge update of k_lu
    // e_1 / tol ~ (e_0 / tol)^2 tol / (2 Vt): kappa = 20/V x (Newton tolerance of a 1 V signal); learnt values
    // never drop below 1/30 of it
    o.kappa0 = 20.0 * (opt->nr_reltol + opt->nr_vabstol);
    o.kappa_floor = o.kappa0 / 30.0;
    o.nsave = dc_only ? 0 : nsave;
    o.nfixed = 0;
    if (!dc_only) {
        if (!(t1 > t0)) return fail(CB_ERR_INVALID, "tran needs t1 > t0");
        if (opt->fixed_step) {
            if (!(opt->dt > 0)) return fail(CB_ERR_INVALID, "fixed-step mode needs dt > 0");
            o.nfixed = (long long)std::llround(o.span / opt->dt);
        }
        // outputs, saveat, breakpoints
        const size_t ycount = (size_t)a.O * (size_t)nsave * (size_t)B;
        if (ycount > p->y_capacity) {
            if (p->d_y) cudaFree(p->d_y);
            p->d_y = nullptr;
            CUDA_TRY(cudaMalloc((void**)&p->d_y, std::max<size_t>(1, ycount) * sizeof(double)));
            p->y_capacity = ycount;
        }
        if ((size_t)nsave > p->saveat_capacity) {
            if (p->d_saveat) cudaFree(p->d_saveat);
            CUDA_TRY(cudaMalloc((void**)&p->d_saveat, std::max<size_t>(1, nsave) * sizeof(double)));
            p->saveat_capacity = nsave;
        }
        if (nsave > 0)
            CUDA_TRY(cudaMemcpyAsync(p->d_saveat, saveat, nsave * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        std::vector<double> bp;
        { const int rcb = collect_breakpoints(c, t0, t1, bp); if (rcb != CB_OK) return rcb; }
        if (bp.size() > p->bp_capacity) {
            if (p->d_bp) cudaFree(p->d_bp);
            CUDA_TRY(cudaMalloc((void**)&p->d_bp, std::max<size_t>(1, bp.size()) * sizeof(double)));
            p->bp_capacity = bp.size();
        }
        if (!bp.empty())
            CUDA_TRY(cudaMemcpy(p->d_bp, bp.data(), bp.size() * sizeof(double), cudaMemcpyHostToDevice));
        a.bp = p->d_bp; a.nbp = (int)bp.size();
        a.saveat = p->d_saveat;
        a.y_out = p->d_y;
    } else {
        a.y_out = p->d_xout;
        a.bp = nullptr; a.nbp = 0; a.saveat = nullptr;
    }
    CUDA_TRY(cudaEventRecord(p->ev0, p->stream));
    int rc = run_setup(p, opt);
    if (rc != CB_OK) return rc;
    CUDA_TRY(cudaMemsetAsync(p->d_done, 0, sizeof(int), p->stream));
    CUDA_TRY(cudaMemsetAsync(p->d_cnt, 0, CB_NBUF * 4 * sizeof(int), p->stream));
    k_init_state<<<(unsigned)((B + 127) / 128), 128, 0, p->stream>>>(B, a.N, a.ist, a.dst, a.alpha, a.active, a.X, a.XN, a.BETA,
                                                                          p->have_x0 ? p->d_x0 : nullptr, p->x0_stride, o,
                                                                          p->d_lists, p->d_cnt);
    CUDA_TRY(cudaGetLastError());

    // ---- round schedule.  Every round: the eval kernels of the points that take a FULL iteration (k_eval_*: currents,
    // charges, Jacobian stamps) and of the points that take a VALUE-ONLY iteration (k_evalv_*: derivative-free, ~1/2 of
    // the instructions), k_lu over both lists (refactor + solve / solve with the stored factors), k_control (Newton
    // update, step control, and the point lists of the next round).  Two schedules:
    //   mixed rounds (default with value_rounds > 0): every unfinished point iterates in EVERY round and follows its own
    //     cycle: full iteration, value_rounds chord iterations, full, ...  No point ever idles, the host makes no
    //     scheduling decision, and the launches stay dense through the device-wide compaction.
    //   lock-step rounds (mixed_rounds = 0): one full round, then value_rounds value-only rounds in which the points that
    //     need a fresh Jacobian idle; all rounds are full while any point is still in its DC phase.
    // Both produce the same iterates.  A window of rounds is captured in a CUDA graph (argument buffers alternate with the
    // round parity) and replayed; the host polls two device counters once per window and keeps one window queued ahead,
    // so launch and polling latency stay off the critical path however small the batch is.
    const bool timing = p->timing || std::getenv("CB_TIMING") != nullptr;
    double t_eval = 0, t_newton = 0, t_evalv = 0, t_newtonv = 0;
    int64_t rounds = 0, vrounds = 0, launches = 0;
    const int64_t max_rounds = std::getenv("CB_MAX_ROUNDS") ? std::atoll(std::getenv("CB_MAX_ROUNDS")) : (int64_t)1 << 40;

    struct SArgsH {
        long long B; const double* X; const double* alpha; const double* gshunt; const double* BETA; const double* dev_out;
        const double* lin_g; const double* lin_c; const double* WV; const int* active; double* DX; double* QK; double* RMAX; int* BAD;
        double* DVMAX;
    } sargs{B, a.X, a.alpha, a.dst + (size_t)DS_GSHUNT * B, a.BETA, a.dev_out, a.lin_g, a.lin_c, p->d_WV, a.active,
            p->d_DX, p->d_QK, p->d_RMAX, p->d_BAD, p->d_DVMAX};
    k_init_waves<<<(unsigned)((B + 127) / 128), 128, 0, p->stream>>>(a, p->d_WV);
    CUDA_TRY(cudaGetLastError());
    {
        const int dc0 = (o.skip_dc && !o.dc_only) ? 0 : (int)B;
        CUDA_TRY(cudaMemcpyAsync(p->d_dc_count, &dc0, sizeof(int), cudaMemcpyHostToDevice, p->stream));
        p->h_done[1] = dc0;
    }
    const int v_rounds = std::getenv("CB_VROUNDS") ? std::max(0, std::atoi(std::getenv("CB_VROUNDS"))) : std::max(0, opt->value_rounds);
    const bool use_v = p->have_v && !dc_only && v_rounds > 0;
    if (use_v && !p->la.LUF) {   // first use: value-only cache and factor storage
        int rc2 = p->alloc(&p->d_cachev, (size_t)std::max<long long>(1, p->cachev_slots) * p->Bpad);
        if (rc2 == CB_OK) rc2 = p->alloc(&p->la.LUF, (size_t)c->sym.nnz_lu * B);
        if (rc2 != CB_OK) return rc2;
    }
    if (use_v) { int rc2 = run_setupv(p, opt); if (rc2 != CB_OK) return rc2; }
    int mixed = use_v && opt->mixed_rounds != 0;
    if (const char* e = std::getenv("CB_MIXED")) mixed = use_v && std::atoi(e) != 0;
    // per-parity argument blocks
    struct VaArgBuf { char b[256]; };
    std::vector<VaArgBuf> vargs_store[CB_NBUF], vargs_v_store[CB_NBUF];   // per list buffer, one argument block per device model
    for (int par = 0; par < CB_NBUF; par++) { vargs_store[par].resize(c->models.size()); vargs_v_store[par].resize(c->models.size()); }
    auto vargs = [&](int par, size_t m) { return (void*)vargs_store[par][m].b; };
    auto vargs_v = [&](int par, size_t m) { return (void*)vargs_v_store[par][m].b; };
    // The control step as the tail of k_lu (kernels.cuh, k_lu<.., true>) instead of a k_control launch per round
    // (off by default: measured 2.7x SLOWER because the lists lose their runs of consecutive points, profiles/probe_r2o.log,
    // ncu_lu_fused_r2v.json -- see DESIGN.md section 4; CB_FUSE=1 selects it)
    bool fused = false;
    if (const char* e = std::getenv("CB_FUSE")) fused = p->lu && std::atoi(e) != 0;
    // Blocked mode of the fused kernel for SMALL batches (kernels.cuh, k_lu): no point lists, group = 32 consecutive points,
    // control step as the tail of the solve.  While the whole GPU runs at most 4 096 points every k_lu launch is under one
    // wave of CTAs whatever the participation, and the k_control launch it saves is a fifth of a latency-bound round.
    bool blocked = p->lu && std::max(B, p->device_points) <= 4096 && !timing && !opt->pivot_repair;   // (the repair pass lives in the list-based kernel)
    if (const char* e = std::getenv("CB_BLOCKED")) blocked = p->lu && std::atoi(e) != 0 && !opt->pivot_repair;
    if (blocked) fused = true;
    // Round r reads list buffer r % 3 and fills buffer (r + 1) % 3; its k_lu zeroes the counters of the buffer that is
    // filled next: (r + 1) % 3 when k_control fills it after k_lu, (r + 2) % 3 when k_lu fills (r + 1) % 3 itself.
    CArgs cargs[CB_NBUF];
    LArgs largs[CB_NBUF];
    for (int par = 0; par < CB_NBUF; par++) {
        for (size_t m = 0; m < c->models.size(); m++) {
            fill_va_args(p, m, opt, vargs(par, m), false, par);
            if (use_v) fill_va_args(p, m, opt, vargs_v(par, m), true, par);
            if (blocked) {
                VaArgsH* va = (VaArgsH*)vargs(par, m);
                va->list = nullptr; va->active = a.active; va->want = ACT_FULL;
                if (use_v) { va = (VaArgsH*)vargs_v(par, m); va->list = nullptr; va->active = a.active; va->want = ACT_ANY; }
            }
        }
        const int nxt = (par + 1) % CB_NBUF, nxt2 = (par + 2) % CB_NBUF;
        int* lists_next = p->d_lists + (size_t)nxt * 3 * B;
        cargs[par] = CArgs{a, p->d_DX, p->d_QK, p->d_RMAX, p->d_DVMAX, p->d_BAD,
                           CtlArgs{p->d_WV, mixed, 0, p->d_dc_count, v_rounds + 1, 0, lists_next, lists_next + B, lists_next + 2 * B,
                                   p->d_cnt + nxt * 4}};
        largs[par] = p->la;
        largs[par].n = a; largs[par].WV = p->d_WV; largs[par].DX = p->d_DX; largs[par].QK = p->d_QK; largs[par].RMAX = p->d_RMAX;
        largs[par].DVMAX = p->d_DVMAX; largs[par].BAD = p->d_BAD;
        largs[par].LUF = use_v ? p->la.LUF : nullptr;
        largs[par].cur = Lists{p->d_lists + (size_t)par * 3 * B, p->d_lists + (size_t)par * 3 * B + B, p->d_lists + (size_t)par * 3 * B + 2 * B,
                               p->d_cnt + par * 4};
        largs[par].zero_cnt = p->d_cnt + (fused ? nxt2 : nxt) * 4;
        largs[par].k = cargs[par].k;
        largs[par].blocked = blocked ? 1 : 0;
        if (blocked) largs[par].k.next_cnt = nullptr;   // no lists: a.active[] carries the roles
        largs[par].growth_max = opt->pivot_growth_max > 0.0 ? opt->pivot_growth_max : 1e300;
        largs[par].pp_scratch = (opt->pivot_repair && !fused) ? p->d_pp_scratch : nullptr;
    }
    int n_live_models = 0;
    for (size_t m = 0; m < c->models.size(); m++) n_live_models += !c->model_insts[m].empty();
    // k_control: 32 lanes per point (1024-thread CTAs: one per SM) only while the whole GPU runs a small batch, where a
    // round is latency-bound (+4 % at 2 048 points); with several busy lanes such CTAs crowd out the other lanes' kernels
    // (-9 % at 4 lanes x 4 096 points, profiles/probe_r2g.log, probe_r2h.log)
    int ctrl_lanes = std::max(B, p->device_points) <= 4096 && B <= 2048 ? 32 : 8;
    if (const char* e = std::getenv("CB_CTRL_LANES")) ctrl_lanes = std::atoi(e) == 32 ? 32 : 8;
    const unsigned lu_grid = (unsigned)std::min<long long>((B + LU_PTS - 1) / LU_PTS + 1, (long long)p->num_sms);
    // one round on the plan's streams; `vround`: lock-step value-only round (only the value-only list is non-empty);
    // `next_v`: lock-step schedule of the next round; ev: optional timing events {start, after eval, after evalv, end}
    auto launch_round = [&](int par, bool full_kernels, bool value_kernels, int next_v, cudaEvent_t* ev) -> int {
        if (ev) cudaEventRecord(ev[0], p->stream);
        const bool fork = (p->fork_models && (n_live_models > 1 || (full_kernels && value_kernels)) && !ev);
        if (fork) CUDA_TRY(cudaEventRecord(p->ev_fork, p->stream));
        int slot = 0;
        for (int kind = 0; kind < 2; kind++) {          // 0: full evaluation of the full list, 1: value-only of the other
            if (kind == 1 && ev) cudaEventRecord(ev[1], p->stream);
            if (kind == 0 ? !full_kernels : !value_kernels) continue;
            for (size_t m = 0; m < c->models.size(); m++) {
                if (c->model_insts[m].empty()) continue;
                void* kargs[] = {kind ? vargs_v(par, m) : vargs(par, m)};
                const unsigned eval_threads = kind ? p->evalv_threads[m] : p->eval_threads[m];
                dim3 grid((unsigned)((B + eval_threads - 1) / eval_threads), (unsigned)c->model_insts[m].size());
                cudaStream_t ms = (fork && slot > 0 && slot < 8 && p->ustream[slot]) ? p->ustream[slot] : p->stream;
                if (ms != p->stream) CUDA_TRY(cudaStreamWaitEvent(ms, p->ev_fork, 0));
                CUDA_TRY(cudaLaunchKernel((const void*)(kind ? p->k_evalv[m] : p->k_eval[m]), grid, dim3(eval_threads), kargs,
                                          kind ? p->evalv_smem[m] : p->eval_smem[m], ms));
                if (ms != p->stream) {
                    CUDA_TRY(cudaEventRecord(p->uev[slot], ms));
                    CUDA_TRY(cudaStreamWaitEvent(p->stream, p->uev[slot], 0));
                }
                slot++;
                launches++;
            }
        }
        if (ev) cudaEventRecord(ev[2], p->stream);
        if (p->lu) {
            LArgs la = largs[par];
            la.k.next_vround = next_v;
            launch_lu(p, lu_grid, p->stream, la, fused);
        } else {
            void* sargs_ptr[] = {&sargs};
            CUDA_TRY(cudaMemsetAsync(p->d_cnt + ((par + 1) % CB_NBUF) * 4, 0, 4 * sizeof(int), p->stream));
            CUDA_TRY(cudaLaunchKernel((const void*)p->k_solve, dim3((unsigned)((B + 63) / 64)), dim3(64), sargs_ptr, 0, p->stream));
        }
        launches++;
        if (!fused) {
            CArgs ca = cargs[par];
            ca.k.next_vround = next_v;
            if (ctrl_lanes == 32) k_control<32><<<(unsigned)((B + CTRL_PTS - 1) / CTRL_PTS), CTRL_PTS * 32, 0, p->stream>>>(ca);
            else k_control<8><<<(unsigned)((B + CTRL_PTS - 1) / CTRL_PTS), CTRL_PTS * 8, 0, p->stream>>>(ca);
            launches++;
        }
        if (ev) cudaEventRecord(ev[3], p->stream);
        return CB_OK;
    };
    // lock-step schedule: kind of round r (0-based round index since the transient phase began)
    auto is_vround = [&](int64_t since_tran) { return use_v && !mixed && since_tran % (v_rounds + 1) != 0; };

    NvtxRange nvtx_rounds_("cb:rounds (eval / k_lu / control)");
    const bool use_graph = !timing && !(std::getenv("CB_NOGRAPH") && std::atoi(std::getenv("CB_NOGRAPH")) != 0);
    int window = std::getenv("CB_POLL") ? std::max(2, std::atoi(std::getenv("CB_POLL"))) : 24;
    window = ((window + CB_NBUF * (v_rounds + 1) - 1) / (CB_NBUF * (v_rounds + 1))) * (CB_NBUF * (v_rounds + 1));   // whole cycles of the schedule and of the list buffers
    bool done = false;
    int par = 0;
    if (use_graph) {
        // two graphs: the DC phase of the lock-step schedule (all rounds full) and the steady pattern.  Mixed rounds
        // need only one (every round launches both kinds; an empty list costs an empty launch).
        cudaGraphExec_t gexec[2] = {nullptr, nullptr};
        auto capture = [&](bool dc_pattern, cudaGraphExec_t* out) -> int {
            cudaGraph_t g = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
            int rcl = CB_OK;
            for (int r = 0; r < window && rcl == CB_OK; r++) {
                const bool v = !dc_pattern && is_vround(r), nv = !dc_pattern && is_vround(r + 1);
                rcl = launch_round(r % CB_NBUF, mixed || !v, mixed ? use_v : v, nv, nullptr);
            }
            cudaError_t e = cudaStreamEndCapture(p->stream, &g);
            if (rcl != CB_OK) { if (g) cudaGraphDestroy(g); return rcl; }
            if (e != cudaSuccess) return fail(CB_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
            e = cudaGraphInstantiate(out, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) return fail(CB_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
            return CB_OK;
        };
        const int64_t launches_before = launches;
        rc = capture(false, &gexec[0]);
        const int64_t per_window = launches - launches_before;
        launches = launches_before;
        if (rc == CB_OK && !mixed && use_v) { rc = capture(true, &gexec[1]); launches = launches_before; }
        if (rc != CB_OK) { for (auto g : gexec) if (g) cudaGraphExecDestroy(g); return rc; }
        // keep one window queued behind the one being polled: the counters read after window k decide about window k + 2
        int queued = 0;
        cudaEvent_t evw[2];
        cudaEventCreateWithFlags(&evw[0], cudaEventDisableTiming); cudaEventCreateWithFlags(&evw[1], cudaEventDisableTiming);
        int* h_snap = p->h_done + 2;   // pinned: [done, dc] snapshots of the two windows in flight (h_done has 8 ints)
        bool dc_phase = p->h_done[1] > 0;
        cudaError_t gerr = cudaSuccess;
        int64_t k = 0;
        while (!done && rounds < max_rounds && gerr == cudaSuccess) {
            while (queued < 2 && gerr == cudaSuccess) {
                const bool dcg = dc_phase && gexec[1];
                gerr = cudaGraphLaunch(dcg ? gexec[1] : gexec[0], p->stream);
                if (gerr != cudaSuccess) break;
                const int sl = (int)((k + queued) & 1);
                cudaMemcpyAsync(h_snap + 2 * sl, p->d_done, sizeof(int), cudaMemcpyDeviceToHost, p->stream);
                cudaMemcpyAsync(h_snap + 2 * sl + 1, p->d_dc_count, sizeof(int), cudaMemcpyDeviceToHost, p->stream);
                gerr = cudaEventRecord(evw[sl], p->stream);
                rounds += window; launches += per_window;
                if (mixed) vrounds += window;   // every round carries value-only iterations
                else if (!dcg && use_v) vrounds += (int64_t)window / (v_rounds + 1) * v_rounds;
                queued++;
            }
            if (gerr != cudaSuccess) break;
            const int sl = (int)(k & 1);
            gerr = cudaEventSynchronize(evw[sl]);
            queued--; k++;
            done = h_snap[2 * sl] >= B;
            dc_phase = h_snap[2 * sl + 1] > 0;
        }
        if (gerr == cudaSuccess) gerr = cudaStreamSynchronize(p->stream);
        cudaEventDestroy(evw[0]); cudaEventDestroy(evw[1]);
        for (auto g : gexec) if (g) cudaGraphExecDestroy(g);
        if (gerr != cudaSuccess) return fail(CB_ERR_CUDA, std::string("round graph: ") + cudaGetErrorString(gerr));
        p->h_done[0] = h_snap[2 * ((k - 1) & 1)];
    } else {
        std::vector<cudaEvent_t> evs;
        int64_t since_tran = 0;
        while (!done && rounds < max_rounds) {
            const bool dc_phase = p->h_done[1] > 0;
            for (int r = 0; r < window; r++) {
                const bool v = !dc_phase && is_vround(since_tran), nv = !dc_phase && is_vround(since_tran + 1);
                cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
                if (timing) for (auto& e : ev) { cudaEventCreate(&e); evs.push_back(e); }
                rc = launch_round(par, mixed || !v, mixed ? use_v : v, nv, timing ? ev : nullptr);
                if (rc != CB_OK) return rc;
                par = (par + 1) % CB_NBUF;
                rounds++;
                vrounds += (mixed || v) ? 1 : 0;
                if (!dc_phase) since_tran++;
            }
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(p->h_done, p->d_done, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
            CUDA_TRY(cudaMemcpyAsync(p->h_done + 1, p->d_dc_count, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
            CUDA_TRY(cudaStreamSynchronize(p->stream));
            done = p->h_done[0] >= B;
            // timing: e0 | k_eval_* | e1 | k_evalv_* | e2 | k_lu, k_control | e3  (one stream: nothing overlaps).  k_lu and
            // k_control serve both kinds of iteration; their time is split in proportion to the eval kernels' times.
            for (size_t q = 0; q + 3 < evs.size(); q += 4) {
                float ms[3] = {0, 0, 0};
                for (int z = 0; z < 3; z++) cudaEventElapsedTime(&ms[z], evs[q + z], evs[q + z + 1]);
                t_eval += ms[0] * 1e-3; t_evalv += ms[1] * 1e-3;
                const double wv = (ms[0] + ms[1]) > 0 ? ms[1] / (ms[0] + ms[1]) : 0.0;
                t_newton += ms[2] * 1e-3 * (1.0 - wv); t_newtonv += ms[2] * 1e-3 * wv;
                for (int z = 0; z < 4; z++) cudaEventDestroy(evs[q + z]);
            }
            evs.clear();
        }
    }
    CUDA_TRY(cudaEventRecord(p->ev1, p->stream));
    CUDA_TRY(cudaEventSynchronize(p->ev1));
    if (!done) return fail(CB_ERR_STATE, "round limit reached before all points finished");
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        float ms = 0;
        cudaEventElapsedTime(&ms, p->ev0, p->ev1);
        stats->solve_seconds = ms * 1e-3;
        stats->rounds = rounds; stats->kernel_launches = launches + 2 + (int64_t)c->models.size();
        stats->eval_seconds = t_eval; stats->newton_seconds = t_newton;
        stats->value_rounds = vrounds; stats->evalv_seconds = t_evalv; stats->newtonv_seconds = t_newtonv;
        {
            std::vector<int> nf((size_t)B);
            CUDA_TRY(cudaMemcpy(nf.data(), a.ist + (size_t)IS_NFULL * B, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost));
            for (long long i = 0; i < B; i++) stats->full_iters += nf[i];
        }
        {
            std::vector<int> sg((size_t)2 * B);
            CUDA_TRY(cudaMemcpy(sg.data(), a.ist + (size_t)IS_SRCSTEP * B, (size_t)2 * B * sizeof(int), cudaMemcpyDeviceToHost));
            for (long long i = 0; i < B; i++) { stats->dc_source_stepped += sg[i]; stats->pivot_fallbacks += sg[(size_t)B + i]; }
        }
        std::vector<int> cnt((size_t)3 * B);
        CUDA_TRY(cudaMemcpy(cnt.data(), a.ist + (size_t)IS_NNEWTON * B, (size_t)3 * B * sizeof(int), cudaMemcpyDeviceToHost));
        for (long long i = 0; i < B; i++) {
            stats->newton_iters += cnt[i];
            stats->steps_accepted += cnt[(size_t)B + i];
            stats->steps_rejected += cnt[(size_t)2 * B + i];
        }
        stats->lu_factors = stats->newton_iters;
    }
    return CB_OK;
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int dc_device1(cb_plan* p, const cb_options* opt, double** d_x_out, int32_t** d_status, cb_stats* stats) {
    if (!p || !opt) return fail(CB_ERR_INVALID, "null argument");
    int rc = solve(p, opt, true, 0.0, 1.0, nullptr, 0, stats);
    if (rc != CB_OK) return rc;
    if (d_x_out) *d_x_out = p->d_xout;
    if (d_status) *d_status = p->na.ist + (size_t)IS_STATUS * p->B;
    return CB_OK;
}

// copy rows of a lane's [rows][B] device array into columns [0, B) of the caller's [rows][pitch] array
static cudaError_t rows_to_host(void* dst, const void* src, size_t rows, long long B, long long pitch, size_t elem) {
    if (rows == 0) return cudaSuccess;
    return cudaMemcpy2D(dst, (size_t)pitch * elem, src, (size_t)B * elem, (size_t)B * elem, rows, cudaMemcpyDeviceToHost);
}

static int dc1(cb_plan* p, const cb_options* opt, double* x_out, double* x_full, int32_t* status, cb_stats* stats, long long pitch) {
    NvtxRange nvtx_("cb:dc");
    if (!p || !opt) return fail(CB_ERR_INVALID, "null argument");
    int rc = solve(p, opt, true, 0.0, 1.0, nullptr, 0, stats);
    if (rc != CB_OK) return rc;
    const double t = now_s();
    const long long B = p->B;
    if (x_out && p->na.O > 0) CUDA_TRY(rows_to_host(x_out, p->d_xout, (size_t)p->na.O, B, pitch, sizeof(double)));
    if (x_full) CUDA_TRY(rows_to_host(x_full, p->na.X, (size_t)p->na.N, B, pitch, sizeof(double)));
    if (status) CUDA_TRY(cudaMemcpy(status, p->na.ist + (size_t)IS_STATUS * B, B * sizeof(int), cudaMemcpyDeviceToHost));
    if (stats) stats->d2h_seconds = now_s() - t;
    return CB_OK;
}

// ---- forward sensitivities of the operating point, direct method (reference test/sensitivity.jl:31-41:
// ODEForwardSensitivityProblem over the ParamSim's parameters).  After the operating point x* has been found and the
// Jacobian J(x*) factored (the factors k_lu stores for its value-only iterations), a direction d -- the caller's
// parameter matrices P+ and P- = the sweep's parameters with one swept quantity moved by +-step -- costs two chord
// updates with those factors and NO new nonlinear solve:
//       dx+- = -J^-1 F(x*; P+-)           (k_setupv + k_evalv with the moved parameters, k_lu with the stored factors)
//       dx*/dp = (dx+ - dx-) / (2 step)   (F(x*; P) cancels; the error is O(step^2) from the devices' own curvature in p)
// i.e. dF/dp comes from central differences of the DEVICE equations at fixed x*, the linear algebra is one forward /
// backward substitution per direction with the Newton factors.  The stencil form (sweeps.sensitivities_ with
// method="stencil": every direction re-solved as extra sweep points) stays as the check.
static int sens_dc1(cb_plan* p, const cb_options* opt, int64_t n_dir, const double* pp, const double* pm, const double* step,
                    double* x_out, double* sens_out, int32_t* status, cb_stats* stats, long long pitch) {
    cb_circuit* c = p->c;
    if (!p->lu || (!p->have_v && !c->insts.empty()))
        return fail(CB_ERR_STATE, "direct sensitivities need the shared-memory solver and value-only variants of every device model");
    int rc = dc1(p, opt, x_out, nullptr, status, stats, pitch);
    if (rc != CB_OK) return rc;
    const long long B = p->B;
    const int O = p->na.O, P = c->P;
    NArgs a = p->na;
    if (!p->la.LUF) {
        rc = p->alloc(&p->la.LUF, (size_t)c->sym.nnz_lu * B);
        if (rc == CB_OK && p->have_v && !p->d_cachev) rc = p->alloc(&p->d_cachev, (size_t)std::max<long long>(1, p->cachev_slots) * p->Bpad);
        if (rc != CB_OK) return rc;
    } else if (p->have_v && !p->d_cachev) {
        rc = p->alloc(&p->d_cachev, (size_t)std::max<long long>(1, p->cachev_slots) * p->Bpad);
        if (rc != CB_OK) return rc;
    }
    double *d_pert = nullptr, *d_step = nullptr, *d_sens = nullptr, *d_lg = nullptr, *d_lc = nullptr;
    auto cleanup = [&]() { cudaFree(d_pert); cudaFree(d_step); cudaFree(d_sens); cudaFree(d_lg); cudaFree(d_lc); };
    const long long Bl = c->lin_swept ? B : 1;
    if (cudaMalloc((void**)&d_pert, (size_t)std::max(1, P) * B * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&d_step, (size_t)B * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&d_sens, (size_t)std::max(1, O) * B * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&d_lg, (size_t)std::max(1, c->nlin) * Bl * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&d_lc, (size_t)std::max(1, c->nlin) * Bl * sizeof(double)) != cudaSuccess) {
        cleanup();
        return fail(CB_ERR_CUDA, "cudaMalloc of the sensitivity scratch failed");
    }
    cudaStream_t st = p->stream;
    const unsigned gB = (unsigned)((B + 127) / 128);
    LArgs la = p->la;
    la.WV = p->d_WV; la.DX = p->d_DX; la.QK = p->d_QK; la.RMAX = p->d_RMAX; la.DVMAX = p->d_DVMAX; la.BAD = p->d_BAD;
    la.cur = Lists{p->d_lists, p->d_lists + B, p->d_lists + 2 * B, p->d_cnt};
    la.zero_cnt = p->d_cnt + 4;
    la.growth_max = 1e300;
    la.pp_scratch = nullptr;
    a.o.rate_test = 0;
    const unsigned lu_grid = (unsigned)std::min<long long>((B + LU_PTS - 1) / LU_PTS + 1, (long long)p->num_sms);
    // 1. factors of J(x*): one full iteration of every point at the solution (alpha = 0: the DC Jacobian)
    k_ac_prepare<<<gB, 128, 0, st>>>(B, a.active, a.alpha, p->d_lists, p->d_cnt);
    for (size_t m = 0; m < c->models.size(); m++) {
        if (c->model_insts[m].empty()) continue;
        char args[256];
        fill_va_args(p, m, opt, args);
        void* kargs[] = {args};
        dim3 grid((unsigned)((B + p->eval_threads[m] - 1) / p->eval_threads[m]), (unsigned)c->model_insts[m].size());
        CUDA_TRY(cudaLaunchKernel((const void*)p->k_eval[m], grid, dim3(p->eval_threads[m]), kargs, p->eval_smem[m], st));
    }
    k_init_waves<<<gB, 128, 0, st>>>(a, p->d_WV);
    la.n = a;
    launch_lu(p, lu_grid, st, la);
    CUDA_TRY(cudaGetLastError());
    // 2. two chord updates per direction with the moved parameters
    a.params = d_pert;
    if (c->lin_swept) { a.lin_g = d_lg; a.lin_c = d_lc; }
    la.n = a;
    cudaError_t err = cudaSuccess;
    for (int64_t d = 0; d < n_dir && err == cudaSuccess; d++) {
        cudaMemsetAsync(d_sens, 0, (size_t)std::max(1, O) * B * sizeof(double), st);
        cudaMemcpyAsync(d_step, step + (size_t)d * pitch, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, st);
        for (int sg = 0; sg < 2 && err == cudaSuccess; sg++) {
            const double* src = (sg == 0 ? pp : pm) + (size_t)d * std::max(1, P) * pitch;
            if (P > 0)
                cudaMemcpy2DAsync(d_pert, (size_t)B * sizeof(double), src, (size_t)pitch * sizeof(double), (size_t)B * sizeof(double),
                                  (size_t)P, cudaMemcpyHostToDevice, st);
            if (c->lin_swept)
                k_lin_setup<<<gB, 128, 0, st>>>(Bl, B, c->nlin, (int)c->lin_contrib.size(), p->d_lin_contrib, d_pert, d_lg, d_lc);
            k_init_waves<<<gB, 128, 0, st>>>(a, p->d_WV);
            k_list_identity_any<<<gB, 128, 0, st>>>(B, p->d_lists + B, p->d_cnt);
            for (size_t m = 0; m < c->models.size(); m++) {
                if (c->model_insts[m].empty()) continue;
                char args[256];
                fill_va_args(p, m, opt, args, true);
                ((VaArgsH*)args)->params = d_pert;
                void* kargs[] = {args};
                dim3 gs((unsigned)((B + 127) / 128), (unsigned)c->model_insts[m].size());
                cudaLaunchKernel((const void*)p->k_setupv[m], gs, dim3(128), kargs, 0, st);
                dim3 grid((unsigned)((B + p->evalv_threads[m] - 1) / p->evalv_threads[m]), (unsigned)c->model_insts[m].size());
                cudaLaunchKernel((const void*)p->k_evalv[m], grid, dim3(p->evalv_threads[m]), kargs, p->evalv_smem[m], st);
            }
            launch_lu(p, lu_grid, st, la);
            k_sens_accum<<<gB, 128, 0, st>>>(B, O, a.outputs, p->d_DX, d_step, sg == 0 ? 1.0 : -1.0, d_sens);
            err = cudaGetLastError();
        }
        if (err == cudaSuccess) err = cudaStreamSynchronize(st);
        if (err == cudaSuccess && O > 0)
            err = rows_to_host(sens_out + (size_t)d * O * pitch, d_sens, (size_t)O, B, pitch, sizeof(double));
    }
    cleanup();
    p->setupv_valid = false;   // the value-only cache now holds the last moved parameters
    if (err != cudaSuccess) return fail(CB_ERR_CUDA, std::string("sensitivity pass: ") + cudaGetErrorString(err));
    if (stats) { stats->kernel_launches += 4 + n_dir * 2 * (4 + 2 * (int64_t)c->models.size()); stats->lu_factors += B; }
    return CB_OK;
}

static int tran_device1(cb_plan* p, double t0, double t1, const double* saveat, int64_t n_save,
                        const cb_options* opt, double** d_y_out, int32_t** d_status, cb_stats* stats) {
    if (!p || !opt || (n_save > 0 && !saveat)) return fail(CB_ERR_INVALID, "null argument");
    for (int64_t k = 1; k < n_save; k++)
        if (!(saveat[k] >= saveat[k - 1])) return fail(CB_ERR_INVALID, "saveat must be ascending");
    int rc = solve(p, opt, false, t0, t1, saveat, n_save, stats);
    if (rc != CB_OK) return rc;
    if (d_y_out) *d_y_out = p->d_y;
    if (d_status) *d_status = p->na.ist + (size_t)IS_STATUS * p->B;
    return CB_OK;
}

static int tran1(cb_plan* p, double t0, double t1, const double* saveat, int64_t n_save, const cb_options* opt,
                 double* y_out, int32_t* status, cb_stats* stats, long long pitch) {
    int rc = tran_device1(p, t0, t1, saveat, n_save, opt, nullptr, nullptr, stats);
    if (rc != CB_OK) return rc;
    NvtxRange nvtx_("cb:tran results to host");
    const double t = now_s();
    const long long B = p->B;
    if (y_out && n_save > 0 && p->na.O > 0)
        CUDA_TRY(rows_to_host(y_out, p->d_y, (size_t)p->na.O * n_save, B, pitch, sizeof(double)));
    if (status) CUDA_TRY(cudaMemcpy(status, p->na.ist + (size_t)IS_STATUS * B, B * sizeof(int), cudaMemcpyDeviceToHost));
    if (stats) stats->d2h_seconds = now_s() - t;
    return CB_OK;
}

// ------------------------------------------------------------------------------------------
// Small-signal analyses about the DC operating point (reference: ac! / noise!, src/ac.jl:75-190).
static int ac_tables(cb_plan* p, bool noise) {
    cb_circuit* c = p->c;
    const cb::Symbolic& S = c->sym;
    const int N = c->N;
    int rc;
#define TRY(x) do { rc = (x); if (rc != CB_OK) return rc; } while (0)
    if (!p->ac_ready) {
        int max_smem = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device));
        p->ac_smem = (size_t)(S.nnz_lu + N) * AC_PTS * sizeof(double2);
        if (p->ac_smem + sizeof(double) * AC_W * AC_PTS + 1024 > (size_t)max_smem)
            return fail(CB_ERR_INVALID, "circuit too large for the shared-memory complex LU of cb_ac / cb_noise");
        CUDA_TRY(cudaFuncSetAttribute(k_ac<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->ac_smem));
        CUDA_TRY(cudaFuncSetAttribute(k_ac<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->ac_smem));
        int* ti;
        TRY(p->upload(&ti, S.u_col)); p->aa.u_col = ti;
        TRY(p->upload(&ti, c->a_csrc)); p->aa.a_csrc = ti;
        // b = -dF/d(eps): sources are dc + eps |ac| (src/simpledevices.jl:292-294, :331-333)
        std::vector<double> rhs(N, 0.0);
        for (const cb_device& d : c->devs) {
            if (d.wave < 0) continue;
            const double ac = c->waves[d.wave].ac_mag;
            if (d.kind == CB_DEV_VSRC) rhs[S.row_to_step[d.branch]] += ac;
            else if (d.kind == CB_DEV_ISRC) {
                if (d.n[0] >= 0) rhs[S.row_to_step[d.n[0]]] -= d.mult * ac;
                if (d.n[1] >= 0) rhs[S.row_to_step[d.n[1]]] += d.mult * ac;
            }
        }
        TRY(p->upload(&p->d_ac_rhs, rhs)); p->aa.ac_rhs = p->d_ac_rhs;
        p->ac_ready = true;
    }
    if (noise && !p->noise_ready) {
        std::vector<ResTab> rt;
        for (const cb_device& d : c->devs) {
            if (d.kind != CB_DEV_R) continue;
            ResTab r{};
            r.pos = d.n[0] < 0 ? -1 : S.row_to_step[d.n[0]];
            r.neg = d.n[1] < 0 ? -1 : S.row_to_step[d.n[1]];
            r.r.value = d.value.value; r.r.col = d.value.col; r.r.pad = 0;
            r.mult = d.mult;
            rt.push_back(r);
        }
        std::vector<NoiseTab> nt;
        long long rows = 0, slots = 0;
        p->noise_off.assign(c->models.size(), 0);
        p->cachen_off.assign(c->models.size(), 0);
        for (size_t m = 0; m < c->models.size(); m++) {
            const ModelH& M = c->models[m];
            p->noise_off[m] = rows;
            p->cachen_off[m] = slots;
            const int K = (int)M.noise_pos.size();
            if (K == 0 || c->model_insts[m].empty()) continue;
            if (!p->k_evaln[m]) return fail(CB_ERR_STATE, "model " + M.name + " has noise sources but the CUDA source has no noise variant");
            for (size_t k = 0; k < c->model_insts[m].size(); k++) {
                const VaInstH& v = c->insts[c->model_insts[m][k]];
                for (int s = 0; s < K; s++) {
                    NoiseTab t{};
                    const int tp = M.noise_pos[s] < 0 ? -1 : v.term[M.noise_pos[s]];
                    const int tn = M.noise_neg[s] < 0 ? -1 : v.term[M.noise_neg[s]];
                    if (tp == tn) continue;
                    t.pos = tp < 0 ? -1 : S.row_to_step[tp];
                    t.neg = tn < 0 ? -1 : S.row_to_step[tn];
                    t.pwr_row = (int)(rows + (long long)k * 2 * K + s);
                    t.exp_row = t.pwr_row + K;
                    t.mult = v.mult;
                    nt.push_back(t);
                }
            }
            rows += (long long)c->model_insts[m].size() * 2 * K;
            slots += (long long)c->model_insts[m].size() * pad4(M.ncache_n);
        }
        ResTab* dr; NoiseTab* dn;
        TRY(p->upload(&dr, rt)); TRY(p->upload(&dn, nt));
        p->aa.rtab = dr; p->aa.ntab = dn; p->aa.nres = (int)rt.size(); p->aa.nnoise = (int)nt.size();
        TRY(p->alloc(&p->d_noise_out, (size_t)std::max<long long>(1, rows) * p->B));
        TRY(p->alloc(&p->d_cachen, (size_t)std::max<long long>(1, slots) * p->Bpad));
        p->aa.noise_out = p->d_noise_out;
        p->noise_ready = true;
    }
#undef TRY
    return CB_OK;
}

static int small_signal(cb_plan* p, bool noise, const double* freqs, int64_t F, const cb_options* opt, double* out,
                        int32_t* status, cb_stats* stats, long long pitch) {
    NvtxRange nvtx_(noise ? "cb:noise" : "cb:ac");
    if (!p || !opt || !freqs || F <= 0 || !out) return fail(CB_ERR_INVALID, "null argument");
    for (int64_t k = 0; k < F; k++)
        if (!(freqs[k] > 0.0) || !std::isfinite(freqs[k])) return fail(CB_ERR_INVALID, "frequencies must be positive");
    if (F > 65535) return fail(CB_ERR_INVALID, "more than 65535 frequencies in one call");
    cb_circuit* c = p->c;
    int rc = solve(p, opt, true, 0.0, 1.0, nullptr, 0, stats);   // operating point, X left on the device
    if (rc != CB_OK) return rc;
    rc = ac_tables(p, noise);
    if (rc != CB_OK) return rc;
    const long long B = p->B;
    NArgs& a = p->na;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    CUDA_TRY(cudaEventRecord(e0, p->stream));
    // linearisation: one full device evaluation at the operating point with alpha = 0 -> G rows and C rows of dev_out
    k_ac_prepare<<<(unsigned)((B + 127) / 128), 128, 0, p->stream>>>(B, a.active, a.alpha, p->d_lists, p->d_cnt);
    CUDA_TRY(cudaGetLastError());
    int64_t launches = 1;
    for (size_t m = 0; m < c->models.size(); m++) {
        if (c->model_insts[m].empty()) continue;
        char args[256];
        fill_va_args(p, m, opt, args);
        void* kargs[] = {args};
        dim3 grid((unsigned)((B + p->eval_threads[m] - 1) / p->eval_threads[m]), (unsigned)c->model_insts[m].size());
        CUDA_TRY(cudaLaunchKernel((const void*)p->k_eval[m], grid, dim3(p->eval_threads[m]), kargs, p->eval_smem[m], p->stream));
        launches++;
    }
    if (noise) {
        struct VaArgsHead { long long B; const double* x; const double* alpha; const int* list; const double* cache; double* out; };
        for (size_t m = 0; m < c->models.size(); m++) {
            const ModelH& M = c->models[m];
            if (c->model_insts[m].empty() || M.noise_pos.empty()) continue;
            char args[256];
            fill_va_args(p, m, opt, args);
            VaArgsHead* h = (VaArgsHead*)args;
            h->cache = p->d_cachen + (size_t)p->cachen_off[m] * p->Bpad;
            ((VaArgsH*)args)->uni = p->d_uni[2][m];
            h->out = p->d_noise_out + (size_t)p->noise_off[m] * B;
            void* kargs[] = {args};
            if (!p->setupn_valid) {
                dim3 g((unsigned)((B + 127) / 128), (unsigned)c->model_insts[m].size());
                CUDA_TRY(cudaLaunchKernel((const void*)p->k_setupn[m], g, dim3(128), kargs, 0, p->stream));
                launches++;
            }
            dim3 grid((unsigned)((B + p->evaln_threads[m] - 1) / p->evaln_threads[m]), (unsigned)c->model_insts[m].size());
            CUDA_TRY(cudaLaunchKernel((const void*)p->k_evaln[m], grid, dim3(p->evaln_threads[m]), kargs, p->evaln_smem[m], p->stream));
            launches++;
        }
        p->setupn_valid = true;
    }
    double* d_freqs = nullptr;
    double* d_out = nullptr;
    const size_t out_count = (size_t)a.O * (size_t)F * (size_t)B * (noise ? 1 : 2);
    CUDA_TRY(cudaMalloc((void**)&d_freqs, (size_t)F * sizeof(double)));
    if (cudaMalloc((void**)&d_out, std::max<size_t>(1, out_count) * sizeof(double)) != cudaSuccess) {
        cudaFree(d_freqs);
        return fail(CB_ERR_CUDA, "cudaMalloc of the small-signal result failed");
    }
    cudaMemcpyAsync(d_freqs, freqs, (size_t)F * sizeof(double), cudaMemcpyHostToDevice, p->stream);
    AArgs aa = p->aa;
    aa.n = a; aa.freqs = d_freqs; aa.F = (int)F; aa.out = d_out;
    aa.temp_val = opt->temp.value; aa.temp_col = opt->temp.col;
    const dim3 grid((unsigned)((B + AC_PTS - 1) / AC_PTS), (unsigned)F);
    if (noise) k_ac<true><<<grid, AC_PTS * AC_W, p->ac_smem, p->stream>>>(aa);
    else k_ac<false><<<grid, AC_PTS * AC_W, p->ac_smem, p->stream>>>(aa);
    launches++;
    cudaError_t err = cudaGetLastError();
    cudaEventRecord(e1, p->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(p->stream);
    const double t = now_s();
    if (err == cudaSuccess && out_count)
        err = rows_to_host(out, d_out, (size_t)a.O * (size_t)F, B, pitch, sizeof(double) * (noise ? 1 : 2));
    if (err == cudaSuccess && status)
        err = cudaMemcpy(status, a.ist + (size_t)IS_STATUS * B, B * sizeof(int), cudaMemcpyDeviceToHost);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_freqs); cudaFree(d_out);
    if (err != cudaSuccess) return fail(CB_ERR_CUDA, std::string("small-signal solve: ") + cudaGetErrorString(err));
    if (stats) {
        stats->d2h_seconds = now_s() - t;
        stats->newton_seconds += ms * 1e-3;   // linearisation + complex LU of all (point, frequency) systems
        stats->solve_seconds += ms * 1e-3;
        stats->kernel_launches += launches;
        stats->lu_factors += (int64_t)B * F;
    }
    return CB_OK;
}

extern "C" int cb_plan_set_timing(cb_plan* p, int enable) {
    if (!p) return fail(CB_ERR_INVALID, "null plan");
    p->timing = enable != 0;
    for (cb_plan* l : p->lanes) l->timing = enable != 0;
    return CB_OK;
}

extern "C" int cb_measure_fp64_peak(int device_id, double* tflops) {
    if (!tflops) return fail(CB_ERR_INVALID, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(CB_ERR_NO_DEVICE, "no CUDA device");
    CUDA_TRY(cudaSetDevice(device_id));
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_id));
    const int blocks = sms * 8, threads = 256, iters = 8192;
    double* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        k_fp64_peak<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = (double)blocks * threads * iters * 8.0 * 2.0 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *tflops = best;
    return CB_OK;
}

// ------------------------------------------------------------------------------------------
// Lanes: the public entry points fan out over the lanes of a plan, one host thread per lane.
static int default_lanes(int64_t n_inst) {
    if (const char* e = std::getenv("CB_LANES")) return std::max(1, std::atoi(e));
    // every lane keeps >= 2048 points (its kernels still fill the 148 SMs); 4 lanes measured best on B200
    return (int)std::max<int64_t>(1, std::min<int64_t>(4, n_inst / 2048));
}

// points [off, off + n) of the parent on device `device_id`, split into n_lanes lanes
static int add_lanes(cb_plan* parent, cb_circuit* c, long long off, long long n, int device_id, int n_lanes) {
    if (n_lanes <= 0) n_lanes = default_lanes(n);
    n_lanes = (int)std::min<long long>(n_lanes, n);
    // lane sizes: multiples of 128 points (cache blocks of the device kernels), the last lane takes the remainder
    const long long per = ((n + n_lanes - 1) / n_lanes + 127) / 128 * 128, end = off + n;
    while (off < end) {
        const long long nb = std::min<long long>(per, end - off);
        cb_plan* l = nullptr;
        int rc = plan_create1(c, nb, device_id, &l);
        if (rc != CB_OK) return rc;
        l->lane_off = off;
        l->device_points = n;
        parent->lanes.push_back(l);
        off += nb;
    }
    return CB_OK;
}

extern "C" int cb_plan_create_lanes(cb_circuit* c, int64_t n_inst, int device_id, int n_lanes, cb_plan** out) {
    if (!c || !out || n_inst <= 0) return fail(CB_ERR_INVALID, "bad argument");
    if (n_lanes <= 0) n_lanes = default_lanes(n_inst);
    n_lanes = (int)std::min<int64_t>(n_lanes, n_inst);
    if (n_lanes == 1) return plan_create1(c, n_inst, device_id, out);
    if (!c->compiled) return fail(CB_ERR_STATE, "circuit not compiled");
    auto p = std::make_unique<cb_plan>();
    p->c = c; p->B = n_inst; p->device = device_id;
    int rc = add_lanes(p.get(), c, 0, n_inst, device_id, n_lanes);
    if (rc != CB_OK) { cb_plan_destroy(p.release()); return rc; }
    p->na.O = (int)c->outputs.size(); p->na.N = c->N;
    *out = p.release();
    return CB_OK;
}

// One plan over several GPUs of this process: device g of G owns the contiguous block
// [g ceil(B / G), min(B, (g + 1) ceil(B / G))) of the sweep points (blocks rounded up to 128 points), split into lanes
// like a single-device plan.  Every lane has its own host thread, stream and device context; there is no collective:
// each device copies its slice of the results straight into the caller's host arrays.
extern "C" int cb_plan_create_multi(cb_circuit* c, int64_t n_inst, const int* device_ids, int n_devices, int lanes_per_device,
                                    cb_plan** out) {
    if (!c || !out || n_inst <= 0 || !device_ids || n_devices <= 0) return fail(CB_ERR_INVALID, "bad argument");
    if (!c->compiled) return fail(CB_ERR_STATE, "circuit not compiled");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(CB_ERR_NO_DEVICE, "no CUDA device available; this engine has no CPU fallback");
    for (int g = 0; g < n_devices; g++)
        if (device_ids[g] < 0 || device_ids[g] >= ndev) return fail(CB_ERR_INVALID, "device id out of range");
    auto p = std::make_unique<cb_plan>();
    p->c = c; p->B = n_inst; p->device = device_ids[0];
    const long long per = ((n_inst + n_devices - 1) / n_devices + 127) / 128 * 128;
    for (int g = 0; g < n_devices; g++) {
        const long long lo = std::min<long long>(n_inst, (long long)g * per), hi = std::min<long long>(n_inst, lo + per);
        if (lo >= hi) break;
        int rc = add_lanes(p.get(), c, lo, hi - lo, device_ids[g], lanes_per_device);
        if (rc != CB_OK) { cb_plan_destroy(p.release()); return rc; }
        if (device_ids[g] != device_ids[0]) p->multi_device = true;
    }
    p->na.O = (int)c->outputs.size(); p->na.N = c->N;
    *out = p.release();
    return CB_OK;
}

extern "C" int cb_plan_devices(const cb_plan* p) {
    if (!p) return 0;
    std::vector<int> seen;
    for (const cb_plan* l : p->lanes) if (std::find(seen.begin(), seen.end(), l->device) == seen.end()) seen.push_back(l->device);
    return seen.empty() ? 1 : (int)seen.size();
}

extern "C" int cb_plan_create(cb_circuit* c, int64_t n_inst, int device_id, cb_plan** out) {
    return cb_plan_create_lanes(c, n_inst, device_id, 0, out);
}

extern "C" int cb_plan_lanes(const cb_plan* p) { return p ? std::max<int>(1, (int)p->lanes.size()) : 0; }

// runs fn(lane) on one host thread per lane; returns the first failing code (its message becomes the caller's)
template <class Fn>
static int for_lanes(cb_plan* p, Fn fn) {
    const size_t K = p->lanes.size();
    std::vector<int> rcs(K, CB_OK);
    std::vector<std::string> msgs(K);
    std::vector<std::thread> th;
    for (size_t k = 0; k < K; k++)
        th.emplace_back([&, k] {
            rcs[k] = fn(p->lanes[k], k);
            if (rcs[k] != CB_OK) msgs[k] = g_err;
        });
    for (auto& t : th) t.join();
    for (size_t k = 0; k < K; k++)
        if (rcs[k] != CB_OK) return fail(rcs[k], msgs[k]);
    return CB_OK;
}

static void merge_stats(cb_stats* dst, const std::vector<cb_stats>& st) {
    if (!dst) return;
    std::memset(dst, 0, sizeof(*dst));
    for (const cb_stats& s : st) {
        dst->newton_iters += s.newton_iters; dst->lu_factors += s.lu_factors;
        dst->steps_accepted += s.steps_accepted; dst->steps_rejected += s.steps_rejected;
        dst->rounds += s.rounds; dst->kernel_launches += s.kernel_launches;
        dst->value_rounds += s.value_rounds; dst->full_iters += s.full_iters;
        dst->pivot_fallbacks += s.pivot_fallbacks; dst->dc_source_stepped += s.dc_source_stepped;
        // lanes run concurrently: wall-like times are the maximum over lanes, per-kernel times add up
        dst->solve_seconds = std::max(dst->solve_seconds, s.solve_seconds);
        dst->h2d_seconds = std::max(dst->h2d_seconds, s.h2d_seconds);
        dst->d2h_seconds = std::max(dst->d2h_seconds, s.d2h_seconds);
        dst->eval_seconds += s.eval_seconds; dst->newton_seconds += s.newton_seconds;
        dst->evalv_seconds += s.evalv_seconds; dst->newtonv_seconds += s.newtonv_seconds;
    }
}

// parameters written by the caller into the parent's [P][B] device buffer -> the lanes' own buffers
static int scatter_params(cb_plan* p) {
    if (!p->params_all_dirty) return CB_OK;
    for (cb_plan* l : p->lanes) {
        int rc = set_params1(l, p->d_params_all + l->lane_off, p->B, cudaMemcpyDeviceToDevice);
        if (rc != CB_OK) return rc;
    }
    p->params_all_dirty = false;
    return CB_OK;
}

extern "C" int cb_plan_set_params(cb_plan* p, const double* params) {
    if (!p) return fail(CB_ERR_INVALID, "null plan");
    if (p->lanes.empty()) return set_params1(p, params, p->B, cudaMemcpyHostToDevice);
    for (cb_plan* l : p->lanes) {
        int rc = set_params1(l, params ? params + l->lane_off : nullptr, p->B, cudaMemcpyHostToDevice);
        if (rc != CB_OK) return rc;
    }
    p->params_set = true;
    p->params_all_dirty = false;
    return CB_OK;
}

extern "C" int cb_plan_set_x0(cb_plan* p, const double* x0, int per_point) {
    if (!p) return fail(CB_ERR_INVALID, "null plan");
    if (p->lanes.empty()) return set_x01(p, x0, per_point, p->B);
    for (cb_plan* l : p->lanes) {
        int rc = set_x01(l, x0 ? x0 + (per_point ? l->lane_off : 0) : nullptr, per_point, p->B);
        if (rc != CB_OK) return rc;
    }
    return CB_OK;
}

extern "C" int cb_dc(cb_plan* p, const cb_options* opt, double* x_out, double* x_full, int32_t* status, cb_stats* stats) {
    if (!p || !opt) return fail(CB_ERR_INVALID, "null argument");
    { const int rco = check_options(opt); if (rco != CB_OK) return rco; }
    if (p->lanes.empty()) return dc1(p, opt, x_out, x_full, status, stats, p->B);
    int rc = scatter_params(p);
    if (rc != CB_OK) return rc;
    std::vector<cb_stats> st(p->lanes.size());
    rc = for_lanes(p, [&](cb_plan* l, size_t k) {
        return dc1(l, opt, x_out ? x_out + l->lane_off : nullptr, x_full ? x_full + l->lane_off : nullptr,
                   status ? status + l->lane_off : nullptr, &st[k], p->B);
    });
    merge_stats(stats, st);
    return rc;
}

extern "C" int cb_tran(cb_plan* p, double t0, double t1, const double* saveat, int64_t n_save, const cb_options* opt,
                       double* y_out, int32_t* status, cb_stats* stats) {
    if (!p || !opt) return fail(CB_ERR_INVALID, "null argument");
    { const int rco = check_options(opt); if (rco != CB_OK) return rco; }
    if (p->lanes.empty()) return tran1(p, t0, t1, saveat, n_save, opt, y_out, status, stats, p->B);
    int rc = scatter_params(p);
    if (rc != CB_OK) return rc;
    std::vector<cb_stats> st(p->lanes.size());
    rc = for_lanes(p, [&](cb_plan* l, size_t k) {
        return tran1(l, t0, t1, saveat, n_save, opt, y_out ? y_out + l->lane_off : nullptr,
                     status ? status + l->lane_off : nullptr, &st[k], p->B);
    });
    merge_stats(stats, st);
    return rc;
}

extern "C" int cb_sens_dc(cb_plan* p, const cb_options* opt, int64_t n_dir, const double* params_plus, const double* params_minus,
                          const double* step, double* x_out, double* sens_out, int32_t* status, cb_stats* stats) {
    if (!p || !opt || n_dir < 0 || (n_dir > 0 && (!step || !sens_out || (p->c->P > 0 && (!params_plus || !params_minus)))))
        return fail(CB_ERR_INVALID, "null argument");
    { const int rco = check_options(opt); if (rco != CB_OK) return rco; }
    if (p->lanes.empty()) return sens_dc1(p, opt, n_dir, params_plus, params_minus, step, x_out, sens_out, status, stats, p->B);
    int rc = scatter_params(p);
    if (rc != CB_OK) return rc;
    std::vector<cb_stats> st(p->lanes.size());
    rc = for_lanes(p, [&](cb_plan* l, size_t k) {
        return sens_dc1(l, opt, n_dir, params_plus ? params_plus + l->lane_off : nullptr, params_minus ? params_minus + l->lane_off : nullptr,
                        step ? step + l->lane_off : nullptr, x_out ? x_out + l->lane_off : nullptr,
                        sens_out ? sens_out + l->lane_off : nullptr, status ? status + l->lane_off : nullptr, &st[k], p->B);
    });
    merge_stats(stats, st);
    return rc;
}

// gathers the lanes' [rows][B_lane] results into the parent's contiguous [rows][B] device buffers
static int gather_device(cb_plan* p, size_t rows, bool tran, double** d_out, int32_t** d_status) {
    CUDA_TRY(cudaSetDevice(p->device));
    const size_t count = std::max<size_t>(1, rows * (size_t)p->B);
    double** dst = tran ? &p->d_y_all : &p->d_xout_all;
    if (tran ? count > p->y_all_capacity : !p->d_xout_all) {
        if (*dst) cudaFree(*dst);
        *dst = nullptr;
        CUDA_TRY(cudaMalloc((void**)dst, count * sizeof(double)));
        if (tran) p->y_all_capacity = count;
    }
    if (!p->d_status_all) CUDA_TRY(cudaMalloc((void**)&p->d_status_all, (size_t)p->B * sizeof(int)));
    for (cb_plan* l : p->lanes) {
        if (rows)
            CUDA_TRY(cudaMemcpy2DAsync(*dst + l->lane_off, (size_t)p->B * sizeof(double), tran ? l->d_y : l->d_xout,
                                       (size_t)l->B * sizeof(double), (size_t)l->B * sizeof(double), rows,
                                       cudaMemcpyDeviceToDevice, l->stream));
        CUDA_TRY(cudaMemcpyAsync(p->d_status_all + l->lane_off, l->na.ist + (size_t)IS_STATUS * l->B, (size_t)l->B * sizeof(int),
                                 cudaMemcpyDeviceToDevice, l->stream));
    }
    for (cb_plan* l : p->lanes) CUDA_TRY(cudaStreamSynchronize(l->stream));
    if (d_out) *d_out = *dst;
    if (d_status) *d_status = p->d_status_all;
    return CB_OK;
}

extern "C" int cb_dc_device(cb_plan* p, const cb_options* opt, double** d_x_out, int32_t** d_status, cb_stats* stats) {
    if (!p || !opt) return fail(CB_ERR_INVALID, "null argument");
    if (p->multi_device) return fail(CB_ERR_INVALID, "device-resident entry points need a single-device plan");
    { const int rco = check_options(opt); if (rco != CB_OK) return rco; }
    if (p->lanes.empty()) return dc_device1(p, opt, d_x_out, d_status, stats);
    int rc = scatter_params(p);
    if (rc != CB_OK) return rc;
    std::vector<cb_stats> st(p->lanes.size());
    rc = for_lanes(p, [&](cb_plan* l, size_t k) { return dc_device1(l, opt, nullptr, nullptr, &st[k]); });
    merge_stats(stats, st);
    if (rc != CB_OK) return rc;
    return gather_device(p, (size_t)p->na.O, false, d_x_out, d_status);
}

extern "C" int cb_tran_device(cb_plan* p, double t0, double t1, const double* saveat, int64_t n_save,
                              const cb_options* opt, double** d_y_out, int32_t** d_status, cb_stats* stats) {
    if (!p || !opt) return fail(CB_ERR_INVALID, "null argument");
    if (p->multi_device) return fail(CB_ERR_INVALID, "device-resident entry points need a single-device plan");
    { const int rco = check_options(opt); if (rco != CB_OK) return rco; }
    if (p->lanes.empty()) return tran_device1(p, t0, t1, saveat, n_save, opt, d_y_out, d_status, stats);
    int rc = scatter_params(p);
    if (rc != CB_OK) return rc;
    std::vector<cb_stats> st(p->lanes.size());
    rc = for_lanes(p, [&](cb_plan* l, size_t k) { return tran_device1(l, t0, t1, saveat, n_save, opt, nullptr, nullptr, &st[k]); });
    merge_stats(stats, st);
    if (rc != CB_OK) return rc;
    return gather_device(p, (size_t)p->na.O * (size_t)n_save, true, d_y_out, d_status);
}

static int small_signal_lanes(cb_plan* p, bool noise, const double* freqs, int64_t F, const cb_options* opt, double* out,
                              int32_t* status, cb_stats* stats) {
    if (!p || !opt) return fail(CB_ERR_INVALID, "null argument");
    { const int rco = check_options(opt); if (rco != CB_OK) return rco; }
    if (p->lanes.empty()) return small_signal(p, noise, freqs, F, opt, out, status, stats, p->B);
    int rc = scatter_params(p);
    if (rc != CB_OK) return rc;
    std::vector<cb_stats> st(p->lanes.size());
    rc = for_lanes(p, [&](cb_plan* l, size_t k) {
        return small_signal(l, noise, freqs, F, opt, out ? out + l->lane_off * (noise ? 1 : 2) : nullptr,
                            status ? status + l->lane_off : nullptr, &st[k], p->B);
    });
    merge_stats(stats, st);
    return rc;
}

extern "C" int cb_ac(cb_plan* p, const double* freqs, int64_t n_freq, const cb_options* opt, double* y_out, int32_t* status,
                     cb_stats* stats) {
    return small_signal_lanes(p, false, freqs, n_freq, opt, y_out, status, stats);
}

extern "C" int cb_noise(cb_plan* p, const double* freqs, int64_t n_freq, const cb_options* opt, double* psd, int32_t* status,
                        cb_stats* stats) {
    return small_signal_lanes(p, true, freqs, n_freq, opt, psd, status, stats);
}

extern "C" void cb_plan_destroy(cb_plan* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (cb_plan* l : p->lanes) cb_plan_destroy(l);
    p->lanes.clear();
    if (p->d_params_all) cudaFree(p->d_params_all);
    if (p->d_y_all) cudaFree(p->d_y_all);
    if (p->d_xout_all) cudaFree(p->d_xout_all);
    if (p->d_status_all) cudaFree(p->d_status_all);
    if (p->stream) cudaStreamSynchronize(p->stream);
    for (void* q : p->allocs) cudaFree(q);
    if (p->d_y) cudaFree(p->d_y);
    if (p->d_saveat) cudaFree(p->d_saveat);
    if (p->d_bp) cudaFree(p->d_bp);
    if (p->h_done) cudaFreeHost(p->h_done);
    if (p->lib) cudaLibraryUnload(p->lib);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    for (int m = 0; m < 8; m++) {
        if (p->ev_join[m]) cudaEventDestroy(p->ev_join[m]);
        if (p->mstream[m]) cudaStreamDestroy(p->mstream[m]);
        if (p->uev[m]) cudaEventDestroy(p->uev[m]);
        if (p->ustream[m]) cudaStreamDestroy(p->ustream[m]);
    }
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

extern "C" void cb_circuit_destroy(cb_circuit* c) { delete c; }
