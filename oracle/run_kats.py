#!/usr/bin/env python
"""Slow known-answer runs of the oracle (128^3 shapes, minutes each): the published timed-iteration counts of
benchmark-results/bench_diffusion_scaling_{gpu,cpu}.csv and the 128^3 point value. TEST INFRASTRUCTURE ONLY.

Writes oracle/KAT_RESULTS.json (committed) so the pin of the oracle at the published shapes is on record even
though the default CPU test-suite only re-runs the fast shapes.  Usage: python oracle/run_kats.py [filter]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_lib as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
kats = json.load(open(os.path.join(HERE, "..", "tests", "golden", "part1_kats.json")))
out_path = os.path.join(HERE, "KAT_RESULTS.json")
results = json.load(open(out_path)) if os.path.exists(out_path) else {}

cases = []
for rec in kats["scaling_counts"]:
    name = "ranks%d_%s_dims%s" % (rec["ranks"], "strong" if rec["strong_scaling"] else "weak",
                                   "x".join(map(str, rec["dims"])))
    cases.append((name, tuple(rec["local"]), tuple(rec["dims"]), not rec["strong_scaling"], O.HALO_REFERENCE_LAG2,
                  rec["timed_iters"]))
# z-slab layouts used by the B200 path (by x<->z symmetry they must equal the published 2x1x1 rows)
cases.append(("zslab2_strong_dims1x1x2", (128, 128, 64), (1, 1, 2), False, O.HALO_REFERENCE_LAG2, 13074))
cases.append(("zslab2_weak_dims1x1x2", (128, 128, 128), (1, 1, 2), True, O.HALO_REFERENCE_LAG2, 12499))
cases.append(("ranks2_strong_dims2x1x1_consistent", (64, 128, 128), (2, 1, 1), False, O.HALO_CONSISTENT, None))
cases.append(("zslab4_strong_dims1x1x4", (128, 128, 32), (1, 1, 4), False, O.HALO_REFERENCE_LAG2, None))

flt = sys.argv[1] if len(sys.argv) > 1 else ""
for name, local, dims, scale, halo, expect in cases:
    if flt not in name or name in results:
        continue
    t0 = time.time()
    d = O.Diffusion3D(*local, dims=dims, halo_mode=halo, scale_physical_size=scale)
    its = d.run(ttot=2.0, tol=1e-6)
    rec = {"local": local, "dims": dims, "scale_physical_size": scale, "halo_mode": halo, "iters_per_step": its,
           "total_iters": sum(its), "timed_iters": sum(its[3:]), "published_timed_iters": expect,
           "match": (sum(its[3:]) == expect) if expect is not None else None, "seconds": round(time.time() - t0, 1),
           "threads": O.num_threads()}
    if dims == (1, 1, 1):
        H = d.gather()
        ix = int(round(4.5 / d.dx + 1)) - 1
        rec["H_probe"] = repr(float(H[ix, ix, ix]))
        rec["H_probe_published"] = "0.07998698561461763"
        rec["H_probe_match"] = float(H[ix, ix, ix]) == 0.07998698561461763
    results[name] = rec
    print(name, rec, flush=True)
    json.dump(results, open(out_path, "w"), indent=1)
