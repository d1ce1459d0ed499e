#!/usr/bin/env python3
"""One NON-extrapolated run of the CPU baseline: the oracle (SEAL-algorithm restatement, oracle/) evaluates the WHOLE
16M-4096 query — ComputePowers for all four bundle indices and eval_patstock of all 28 BinBundles of the 2^24-item
synthetic DB — with -t 1 and with all host threads, as BASELINE.md §3 promises.  bench.py's per-run `cpu_baseline`
extrapolates from a bounded sample; this record (profiles/cpu_full_query_r02.json) is what the extrapolation is checked
against.  Needs about 7 GB of host memory (the DB) and, single-threaded, a few minutes.

    python tools/cpu_full_query.py [--threads 1,all] > profiles/cpu_full_query_r02.json
"""
import argparse
import json
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="16M-4096")
    ap.add_argument("--db-log2", type=int, default=24)
    ap.add_argument("--threads", default="all,1")
    args = ap.parse_args()
    O.build()
    pj = bench.load_params_json(args.workload)
    p = O.Params(pj, args.workload + ".json")
    degrees = bench.simulate_bundle_degrees(pj, args.db_log2, bench.SEEDS["db"])
    n_bundles = sum(len(r) for r in degrees)
    bic = p.bundle_idx_count
    cts, relin, masks = bench.synth_query(p.primes, p.t, p.N, p.first_L, p.K, len(p.query_powers), bic, max(len(r) for r in degrees) * bic,
                                          bench.SEEDS["query"])
    ctx = O.Context.from_params(p)
    db = O.ReceiverDB(ctx, p)
    t0 = time.perf_counter()
    for b in range(bic):
        for c, d in enumerate(degrees[b]):
            assert db.add_bundle_synthetic(b, d + 1, bench.SEEDS["db"] * 1000 + b * 64 + c) == c
    fill_s = time.perf_counter() - t0
    runs = []
    for tok in args.threads.split(","):
        t = (os.cpu_count() or 1) if tok == "all" else int(tok)
        ses = db.run_query(p.query_powers, cts, relin, masks, threads=t)
        runs.append({"threads": t, "compute_powers_ms": ses.powers_ms, "eval_ms": ses.eval_ms, "query_eval_ms": ses.powers_ms + ses.eval_ms,
                     "binbundles_per_s": n_bundles / ((ses.powers_ms + ses.eval_ms) / 1e3)})
        digest = bench.results_digest(*_as_arrays(ses.results(), p.N))
        runs[-1]["results_sha256"] = digest
        del ses
    rec = ROOT / "profiles" / f"results_sha256_{args.workload}_2p{args.db_log2}.json"
    gpu_digest = json.loads(rec.read_text())["results_sha256"] if rec.exists() else None
    cpu_model = next((l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "")
    print(json.dumps({
        "what": "oracle (SEAL-algorithm restatement, not SEAL), whole query, no extrapolation", "workload": args.workload, "db_log2": args.db_log2,
        "bin_bundles": n_bundles, "host_cores": os.cpu_count(), "cpu_model": cpu_model, "db_fill_s": fill_s, "runs": runs,
        "gpu_results_sha256": gpu_digest,
        "all_28_results_equal_gpu_digest": bool(gpu_digest and all(r["results_sha256"] == gpu_digest for r in runs)),
    }, indent=1))


def _as_arrays(results, N):
    n = len(results)
    out = np.zeros((n, 2, N), dtype=np.uint64)
    b = np.zeros(n, dtype=np.uint32)
    c = np.zeros(n, dtype=np.uint32)
    for k, (bb, cc, ct) in enumerate(results):
        out[k], b[k], c[k] = ct, bb, cc
    return out, b, c, n


if __name__ == "__main__":
    main()
