# cedarsim_cpu_sweep.jl -- the reference's own CPU path on the benchmark workload, for a Julia-equipped box.
#
# STATUS: UNTESTED here (no Julia toolchain in this image; CedarSim also needs its private registry and the BSIM4 /
# GF180 packages, which are not vendored in the reference tree).  BASELINE.md "B-cedarsim" stays "not measured" until
# this has been run; bench.py's `cpu_baseline` / `--impl reference` time the C++ restatement instead.
#
#   JULIA_NUM_THREADS=$(nproc) julia --project=<CedarSim checkout> bench/cedarsim_cpu_sweep.jl [points] [deck]
#
# Measures what SURVEY.md 8(d) asks for:
#   (a) compile latency  = @elapsed CircuitSweep(...) + first solve
#   (b) steady-state sweep points/s of (i) the reference's own serial loop (src/sweeps.jl:473,490) and
#       (ii) a Threads.@threads loop over remake(prob, p = sim) -- a courtesy upper bound, the reference has no
#       threaded sweep.
# Workload = bench.py's: DFF Monte-Carlo, TandemSweep of pre-drawn (w, l) per FET, tspan (0, 6e-7), reltol 1e-4,
# abstol 1e-6, IDA.  The draws are read from a CSV written by `python scripts/dump_mc_draws.py` so that both sides
# integrate identical instances.
using CedarSim, Sundials, SciMLBase, DelimitedFiles, Printf

points = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 256
deck = length(ARGS) >= 2 ? ARGS[2] : joinpath(dirname(pathof(CedarSim)), "..", "test", "DFF", "DFF_cap_all.cir")
draws_csv = get(ENV, "CB_MC_DRAWS", "mc_draws.csv")            # header: swept names; one row per instance

hdr, vals = let raw = readdlm(draws_csv, ','; header = true); (vec(String.(raw[2])), raw[1]); end
points = min(points, size(vals, 1))
sweep = TandemSweep((Sweep(Symbol(hdr[k]), vals[1:points, k]) for k in eachindex(hdr))...)

ast = CedarSim.SpectreNetlistParser.SPICENetlistParser.SPICENetlistCSTParser.parsefile(deck)
circuit = CedarSim.make_spectre_circuit(ast)                     # as test/gf180_dff.jl:11-20 does

t_compile = @elapsed begin
    cs = CircuitSweep(circuit, sweep)
    sim1 = first(cs)
    prob = DAEProblem(cs.sys, nothing, nothing, (0.0, 6e-7), sim1; initializealg = CedarSim.CedarDCOp())
    sol1 = solve(prob, IDA(); abstol = 1e-6, reltol = 1e-4)
end
@printf("compile + first solve: %.2f s, retcode %s\n", t_compile, sol1.retcode)

sims = collect(cs)
function serial(sims)
    for sim in sims
        solve(remake(prob, p = sim), IDA(); abstol = 1e-6, reltol = 1e-4)
    end
end
function threaded(sims)
    Threads.@threads for i in eachindex(sims)
        solve(remake(prob, p = sims[i]), IDA(); abstol = 1e-6, reltol = 1e-4)
    end
end
serial(sims[1:min(2, end)]); threaded(sims[1:min(2, end)])       # warm up
ns = min(points, 16)
t_serial = @elapsed serial(sims[1:ns])
t_thr = @elapsed threaded(sims)
@printf("{\"impl\": \"cedarsim\", \"points\": %d, \"threads\": %d, \"compile_seconds\": %.3f, \"serial_points_per_s\": %.4f, \"threaded_points_per_s\": %.4f}\n",
        points, Threads.nthreads(), t_compile, ns / t_serial, points / t_thr)
