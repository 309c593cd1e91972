/* abi_smoke.c -- proves that include/cedarb200.h compiles as plain C (gcc -std=c99 -Wall -Werror -pedantic) and reports
 * the layout of every struct of the boundary, so that the tests can compare it with the mirrors of the bindings
 * (ctypes in cedarsim.jl_b200/flat.py, Julia in ext/CedarSimB200Ext.jl).
 *
 *   abi_smoke layout           print {"struct.field": [offset, size], ..., "struct": [0, sizeof]} as JSON
 *   abi_smoke run <lib.so>     dlopen the engine, build the two-resistor divider of the reference's test/sweep.jl:326-340
 *                              through the C ABI alone and solve its DC sweep on the GPU (needs a GPU)
 *   abi_smoke netlist <lib.so> [flatten]
 *                              the same sweep from DECK TEXT: cb_netlist_flatten (the library's own SPICE reader) ->
 *                              cb_netlist_circuit -> cb_circuit_compile -> cb_plan_create -> cb_dc; with `flatten` it stops
 *                              after the host-side steps (no GPU needed) and prints what the reader built
 */
#include <dlfcn.h>
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cedarb200.h"

#define F(S, f) printf("%s\"" #S "." #f "\": [%zu, %zu]", (first ? "" : ", "), offsetof(S, f), sizeof(((S *)0)->f)), first = 0
#define S_(S) printf("%s\"" #S "\": [0, %zu]", (first ? "" : ", "), sizeof(S)), first = 0

static int layout(void) {
    int first = 1;
    printf("{");
    S_(cb_pref); F(cb_pref, value); F(cb_pref, col);
    S_(cb_device); F(cb_device, kind); F(cb_device, n); F(cb_device, branch); F(cb_device, wave); F(cb_device, value); F(cb_device, mult);
    S_(cb_wave); F(cb_wave, kind); F(cb_wave, has_dc); F(cb_wave, dc); F(cb_wave, npts); F(cb_wave, t); F(cb_wave, y); F(cb_wave, v);
    F(cb_wave, ac_mag);
    S_(cb_va_model); F(cb_va_model, name); F(cb_va_model, nterm); F(cb_va_model, nparam); F(cb_va_model, ncache); F(cb_va_model, nj);
    F(cb_va_model, jrow); F(cb_va_model, jcol); F(cb_va_model, host_setup); F(cb_va_model, host_eval); F(cb_va_model, n_noise);
    F(cb_va_model, ncache_n); F(cb_va_model, noise_pos); F(cb_va_model, noise_neg); F(cb_va_model, host_setupn);
    F(cb_va_model, host_noise); F(cb_va_model, linear); F(cb_va_model, ncache_v); F(cb_va_model, host_setupv); F(cb_va_model, host_evalv);
    S_(cb_va_inst); F(cb_va_inst, model); F(cb_va_inst, term); F(cb_va_inst, par); F(cb_va_inst, given); F(cb_va_inst, mult);
    S_(cb_flat_circuit); F(cb_flat_circuit, n_unknowns); F(cb_flat_circuit, n_nodes); F(cb_flat_circuit, n_params);
    F(cb_flat_circuit, n_devices); F(cb_flat_circuit, devices); F(cb_flat_circuit, n_waves); F(cb_flat_circuit, n_va_models);
    F(cb_flat_circuit, waves); F(cb_flat_circuit, va_models); F(cb_flat_circuit, n_va_insts); F(cb_flat_circuit, n_outputs);
    F(cb_flat_circuit, va_insts); F(cb_flat_circuit, outputs);
    S_(cb_options); F(cb_options, struct_size); F(cb_options, abi_version); F(cb_options, temp); F(cb_options, gmin);
    F(cb_options, reltol); F(cb_options, vabstol); F(cb_options, iabstol); F(cb_options, nr_reltol); F(cb_options, nr_vabstol);
    F(cb_options, nr_iabstol); F(cb_options, dc_abstol); F(cb_options, dv_max); F(cb_options, max_newton_dc);
    F(cb_options, max_newton_tran); F(cb_options, method); F(cb_options, fixed_step); F(cb_options, dt); F(cb_options, dt_min);
    F(cb_options, dt_max); F(cb_options, gmin_steps); F(cb_options, skip_dc); F(cb_options, nr_rate_test);
    F(cb_options, value_rounds); F(cb_options, mixed_rounds); F(cb_options, source_steps); F(cb_options, t0_reinit);
    F(cb_options, pivot_repair); F(cb_options, pivot_growth_max);
    S_(cb_stats); F(cb_stats, newton_iters); F(cb_stats, lu_factors); F(cb_stats, steps_accepted); F(cb_stats, steps_rejected);
    F(cb_stats, rounds); F(cb_stats, kernel_launches); F(cb_stats, solve_seconds); F(cb_stats, h2d_seconds); F(cb_stats, d2h_seconds);
    F(cb_stats, eval_seconds); F(cb_stats, newton_seconds); F(cb_stats, value_rounds); F(cb_stats, full_iters);
    F(cb_stats, evalv_seconds); F(cb_stats, newtonv_seconds); F(cb_stats, pivot_fallbacks); F(cb_stats, dc_source_stepped);
    printf("}\n");
    return 0;
}

#define SYM(name) \
    *(void **)(&p_##name) = dlsym(h, #name); \
    if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

static int run(const char *path) {
    void *h = dlopen(path, RTLD_NOW);
    int (*p_cb_version)(void);
    const char *(*p_cb_last_error)(void);
    size_t (*p_cb_options_size)(void);
    int (*p_cb_options_init)(cb_options *, size_t);
    int (*p_cb_circuit_create)(const cb_flat_circuit *, cb_circuit **);
    int (*p_cb_circuit_compile)(cb_circuit *, const char *, double *);
    int (*p_cb_plan_create)(cb_circuit *, int64_t, int, cb_plan **);
    int (*p_cb_plan_set_params)(cb_plan *, const double *);
    int (*p_cb_dc)(cb_plan *, const cb_options *, double *, double *, int32_t *, cb_stats *);
    void (*p_cb_plan_destroy)(cb_plan *);
    void (*p_cb_circuit_destroy)(cb_circuit *);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    SYM(cb_version) SYM(cb_last_error) SYM(cb_options_size) SYM(cb_options_init) SYM(cb_circuit_create) SYM(cb_circuit_compile)
    SYM(cb_plan_create) SYM(cb_plan_set_params) SYM(cb_dc) SYM(cb_plan_destroy) SYM(cb_circuit_destroy)
    if (p_cb_version() != CB_ABI_VERSION || p_cb_options_size() != sizeof(cb_options)) {
        fprintf(stderr, "header / library mismatch\n");
        return 3;
    }
    {
        /* V 1 V between vcc and ground, R1 vcc-out, R2 out-ground; unknowns: vcc, out | I(V).  Swept: R1, R2. */
        enum { B = 400 };
        cb_wave wave;
        cb_device dev[3];
        cb_flat_circuit fc;
        cb_options opt, bad;
        cb_circuit *c = NULL;
        cb_plan *p = NULL;
        cb_stats st;
        int32_t outs[1] = {2};
        static double params[2 * B], x[B], xf[3 * B];
        static int32_t status[B];
        int i, j, rc;
        double err = 0.0;
        memset(&wave, 0, sizeof wave); memset(dev, 0, sizeof dev); memset(&fc, 0, sizeof fc);
        wave.kind = CB_W_DC; wave.has_dc = 1; wave.dc.value = 1.0; wave.dc.col = -1;
        for (i = 0; i < 7; i++) wave.v[i].col = -1;
        for (i = 0; i < 3; i++) { dev[i].mult = 1.0; dev[i].branch = -1; dev[i].wave = -1; dev[i].n[2] = dev[i].n[3] = -1; dev[i].value.col = -1; }
        dev[0].kind = CB_DEV_VSRC; dev[0].n[0] = 0; dev[0].n[1] = -1; dev[0].branch = 2; dev[0].wave = 0;
        dev[1].kind = CB_DEV_R; dev[1].n[0] = 0; dev[1].n[1] = 1; dev[1].value.col = 0;
        dev[2].kind = CB_DEV_R; dev[2].n[0] = 1; dev[2].n[1] = -1; dev[2].value.col = 1;
        fc.n_unknowns = 3; fc.n_nodes = 2; fc.n_params = 2; fc.n_devices = 3; fc.devices = dev; fc.n_waves = 1; fc.waves = &wave;
        fc.n_outputs = 1; fc.outputs = outs;
        for (i = 0; i < 20; i++)
            for (j = 0; j < 20; j++) { params[j * 20 + i] = 100.0 * (i + 1); params[B + j * 20 + i] = 100.0 * (j + 1); }
        /* a binding with a stale (shorter) mirror of cb_options is refused, not overrun */
        if (p_cb_options_init(&bad, sizeof(cb_options) - 8) != CB_ERR_INVALID) { fprintf(stderr, "stale options accepted\n"); return 4; }
        if (p_cb_options_init(&opt, sizeof opt) != CB_OK) { fprintf(stderr, "%s\n", p_cb_last_error()); return 4; }
        rc = p_cb_circuit_create(&fc, &c);
        if (rc == CB_OK) rc = p_cb_circuit_compile(c, NULL, NULL);
        if (rc == CB_OK) rc = p_cb_plan_create(c, B, 0, &p);
        if (rc == CB_OK) rc = p_cb_plan_set_params(p, params);
        memset(&bad, 0, sizeof bad);   /* options that were never initialised are refused by the solve calls */
        if (rc == CB_OK && p_cb_dc(p, &bad, x, xf, status, &st) != CB_ERR_INVALID) { fprintf(stderr, "uninitialised options accepted\n"); return 4; }
        if (rc == CB_OK) rc = p_cb_dc(p, &opt, x, xf, status, &st);
        if (rc != CB_OK) { fprintf(stderr, "cb error %d: %s\n", rc, p_cb_last_error()); return 5; }
        for (i = 0; i < B; i++) {
            const double want = -1.0 / (params[i] + params[B + i]);   /* reference test/sweep.jl:338 */
            if (status[i] != CB_ST_SUCCESS) { fprintf(stderr, "point %d status %d\n", i, status[i]); return 6; }
            if (fabs(x[i] - want) > err) err = fabs(x[i] - want);
        }
        printf("{\"points\": %d, \"max_abs_err\": %.3e, \"newton_iters\": %lld}\n", (int)B, err, (long long)st.newton_iters);
        p_cb_plan_destroy(p);
        p_cb_circuit_destroy(c);
        return err < 1e-12 ? 0 : 7;
    }
}

static int run_netlist(const char *path, int flatten_only) {
    void *h = dlopen(path, RTLD_NOW);
    const char *(*p_cb_last_error)(void);
    int (*p_cb_options_init)(cb_options *, size_t);
    int (*p_cb_netlist_flatten)(const char *, const char *, const char *const *, int, const double *, int64_t, const char *const *, int,
                                cb_netlist **);
    int (*p_cb_netlist_circuit)(cb_netlist *, cb_circuit **);
    int32_t (*p_cb_netlist_n_unknowns)(const cb_netlist *);
    int32_t (*p_cb_netlist_n_params)(const cb_netlist *);
    const double *(*p_cb_netlist_params)(const cb_netlist *);
    const char *(*p_cb_netlist_param_name)(const cb_netlist *, int32_t);
    int32_t (*p_cb_netlist_unknown)(const cb_netlist *, const char *);
    void (*p_cb_netlist_destroy)(cb_netlist *);
    int (*p_cb_circuit_compile)(cb_circuit *, const char *, double *);
    int (*p_cb_plan_create)(cb_circuit *, int64_t, int, cb_plan **);
    int (*p_cb_plan_set_params)(cb_plan *, const double *);
    int (*p_cb_dc)(cb_plan *, const cb_options *, double *, double *, int32_t *, cb_stats *);
    void (*p_cb_plan_destroy)(cb_plan *);
    void (*p_cb_circuit_destroy)(cb_circuit *);
    /* the deck of the reference's test/sweep.jl:342-371: a subcircuit parameter swept through the instance */
    static const char deck[] =
        "* Parameter scoping test\n"
        ".subckt subcircuit1 vss gnd l=11\n"
        ".param r_load=1\n"
        "r1 vss gnd 'r_load'\n"
        ".ends\n"
        ".param v_in=1\n"
        "x1 vcc 0 subcircuit1 r_load=10\n"
        "v1 vcc 0 DC 'v_in'\n";
    enum { B = 256 };
    const char *names[2] = {"v_in", "x1.r_load"};
    const char *outs[1] = {"v1.i"};
    static double vals[2 * B], x[B];
    static int32_t status[B];
    cb_netlist *nl = NULL;
    cb_circuit *c = NULL;
    cb_plan *p = NULL;
    cb_options opt;
    cb_stats st;
    int i, rc;
    double err = 0.0;
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    SYM(cb_last_error) SYM(cb_options_init) SYM(cb_netlist_flatten) SYM(cb_netlist_circuit) SYM(cb_netlist_n_unknowns)
    SYM(cb_netlist_n_params) SYM(cb_netlist_params) SYM(cb_netlist_param_name) SYM(cb_netlist_unknown) SYM(cb_netlist_destroy)
    SYM(cb_circuit_compile) SYM(cb_plan_create) SYM(cb_plan_set_params) SYM(cb_dc) SYM(cb_plan_destroy) SYM(cb_circuit_destroy)
    for (i = 0; i < B; i++) { vals[i] = 0.5 + (i % 16) * 0.1; vals[B + i] = 5.0 + (i / 16); }
    rc = p_cb_netlist_flatten(deck, NULL, names, 2, vals, B, outs, 1, &nl);
    if (rc == CB_OK) rc = p_cb_netlist_circuit(nl, &c);
    if (rc == CB_OK) rc = p_cb_circuit_compile(c, NULL, NULL);
    if (rc != CB_OK) { fprintf(stderr, "cb error %d: %s\n", rc, p_cb_last_error()); return 5; }
    if (p_cb_netlist_n_unknowns(nl) != 2 || p_cb_netlist_n_params(nl) != 2 || p_cb_netlist_unknown(nl, "v1.i") != 1 ||
        p_cb_netlist_unknown(nl, "x1.node_vss") != 0) { fprintf(stderr, "unexpected flat circuit\n"); return 6; }
    if (flatten_only) {
        printf("{\"unknowns\": %d, \"params\": [\"%s\", \"%s\"], \"first_row\": [%.17g, %.17g]}\n", (int)p_cb_netlist_n_unknowns(nl),
               p_cb_netlist_param_name(nl, 0), p_cb_netlist_param_name(nl, 1), p_cb_netlist_params(nl)[0], p_cb_netlist_params(nl)[1]);
        p_cb_circuit_destroy(c);
        p_cb_netlist_destroy(nl);
        return 0;
    }
    if (p_cb_options_init(&opt, sizeof opt) != CB_OK) { fprintf(stderr, "%s\n", p_cb_last_error()); return 4; }
    rc = p_cb_plan_create(c, B, 0, &p);
    if (rc == CB_OK) rc = p_cb_plan_set_params(p, p_cb_netlist_params(nl));
    if (rc == CB_OK) rc = p_cb_dc(p, &opt, x, NULL, status, &st);
    if (rc != CB_OK) { fprintf(stderr, "cb error %d: %s\n", rc, p_cb_last_error()); return 5; }
    for (i = 0; i < B; i++) {
        const double want = -vals[i] / vals[B + i];   /* reference test/sweep.jl:369: I = v_in / r_load through the source */
        if (status[i] != CB_ST_SUCCESS) { fprintf(stderr, "point %d status %d\n", i, status[i]); return 6; }
        if (fabs(x[i] - want) > err) err = fabs(x[i] - want);
    }
    printf("{\"points\": %d, \"max_abs_err\": %.3e}\n", (int)B, err);
    p_cb_plan_destroy(p);
    p_cb_circuit_destroy(c);
    p_cb_netlist_destroy(nl);
    return err < 1e-12 ? 0 : 7;
}

int main(int argc, char **argv) {
    if (argc >= 2 && strcmp(argv[1], "layout") == 0) return layout();
    if (argc >= 3 && strcmp(argv[1], "run") == 0) return run(argv[2]);
    if (argc >= 3 && strcmp(argv[1], "netlist") == 0) return run_netlist(argv[2], argc >= 4 && strcmp(argv[3], "flatten") == 0);
    fprintf(stderr, "usage: abi_smoke layout | run <libcedarb200.so> | netlist <libcedarb200.so> [flatten]\n");
    return 1;
}
