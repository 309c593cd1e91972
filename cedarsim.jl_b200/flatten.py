"""Command-line front end for bindings that have no netlist front end of their own (ext/CedarSimB200Ext.jl):

    python -m cedarsim.jl_b200.flatten deck.cir --sweep sweep.csv --outputs q,d --out /tmp/run1 [--lang spectre]

reads a SPICE (or Spectre) deck and the sweep points (CSV: header = swept names as CircuitSweep spells them, e.g.
`x1.r_load`, `temp`; one row per sweep point in `collect(cs)` order) and writes

    <out>.flatckt      flat circuit + generated CUDA C of its Verilog-A models   -> cb_circuit_load
    <out>.params.f64   the per-point parameter matrix, float64 [P][B] row-major  -> cb_plan_set_params
    <out>.json         {"B", "P", "param_names", "outputs", "unknowns", "options": {"temp": ..}, "tran": [t0, t1] | null}

This is the compile-once step of CircuitSweep(circuit, iterator) (reference src/sweeps.jl:414-417) for callers that
hold a netlist rather than packed structs.
"""
import argparse
import csv
import json

import numpy as np

from . import netlist
from .flat import save_flatckt


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("deck")
    ap.add_argument("--sweep", help="CSV of sweep points (header row = names)")
    ap.add_argument("--outputs", default="", help="comma-separated unknown names (default: every unknown)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--lang", default="spice", choices=["spice", "spectre"])
    args = ap.parse_args(argv)
    text = open(args.deck).read()
    if args.lang == "spectre":
        from . import spectre
        nl = spectre.parse_spectre(text, path=args.deck)
    else:
        nl = netlist.parse_netlist(text, path=args.deck)
    sweep = {}
    if args.sweep:
        with open(args.sweep, newline="") as fh:
            rows = list(csv.reader(fh))
        names = [n.strip() for n in rows[0]]
        vals = np.array([[float(x) for x in r] for r in rows[1:] if r], dtype=np.float64)
        sweep = {n: np.ascontiguousarray(vals[:, k]) for k, n in enumerate(names)}
    outputs = [o.strip() for o in args.outputs.split(",") if o.strip()] or None
    fl = netlist.flatten(nl, sweep, outputs=outputs)
    fc = fl.fc
    save_flatckt(fc, fl.models, args.out + ".flatckt")
    P = np.ascontiguousarray(fl.params, dtype=np.float64)
    P.tofile(args.out + ".params.f64")
    unknowns = list(fc.node_names) + list(fc.branch_names)
    opts = {}
    for k, v in fl.options.items():
        if np.ndim(v) == 0:
            opts[k] = float(v)
        else:   # swept SimSpec value: a column of the parameter matrix
            opts[k] = {"col": fc.param_names.index(k)} if k in fc.param_names else None
    meta = {"B": int(P.shape[1]) if P.ndim == 2 and P.size else (len(next(iter(sweep.values()))) if sweep else 1),
            "P": len(fc.param_names), "param_names": list(fc.param_names), "outputs": [unknowns[o] for o in fc.outputs],
            "unknowns": unknowns, "n_nodes": fc.n_nodes, "options": opts, "tran": list(fl.tran) if fl.tran else None}
    with open(args.out + ".json", "w") as fh:
        json.dump(meta, fh)
    print(json.dumps({"flatckt": args.out + ".flatckt", "B": meta["B"], "P": meta["P"], "unknowns": len(unknowns)}))


if __name__ == "__main__":
    main()
