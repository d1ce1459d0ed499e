#!/usr/bin/env python3
"""One profiled receiver query on the bench workload (for ncu):

  python tools/profile_query.py [--workload 16M-4096] [--warmup 1] [--bundles-per-idx K]
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_query.py

The profiled region (cudaProfilerStart/Stop) is exactly one compute_powers + eval_all.
"""
import argparse
import ctypes as C
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.WORKLOAD)
    ap.add_argument("--db-log2", type=int, default=bench.DB_LOG2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--bundles-per-idx", type=int, default=0, help="truncate every bundle index to this many BinBundles")
    ap.add_argument("--only-idx", type=int, default=-1, help="populate this bundle index only (what one rank of an 8-GPU run holds)")
    args = ap.parse_args()
    import torch
    import apsu_b200
    from apsu_b200 import capi

    pj = bench.load_params_json(args.workload)
    degrees = bench.simulate_bundle_degrees(pj, args.db_log2, bench.SEEDS["db"])
    if args.bundles_per_idx:
        degrees = [row[:args.bundles_per_idx] for row in degrees]
    if args.only_idx >= 0:
        degrees = [row if b == args.only_idx else [] for b, row in enumerate(degrees)]
    params = apsu_b200.PSUParams.Load(json.dumps(pj))
    db = apsu_b200.ReceiverDB(params, 0)
    rx = apsu_b200.Receiver(db)
    for b, row in enumerate(degrees):
        for c, d in enumerate(row):
            db.add_bin_bundle_synthetic(b, d + 1, bench.SEEDS["db"] * 1000 + b * 64 + c)
    N, t, primes = params.poly_modulus_degree(), params.plain_modulus(), params.coeff_modulus()
    bic = params.bundle_idx_count()
    cts, relin, masks = bench.synth_query(primes, t, N, db.level(0), len(primes), len(params.query_powers()), bic,
                                          max(len(r) for r in degrees) * bic, bench.SEEDS["query"])
    rx.load_query(apsu_b200.Query(params.query_powers(), cts, relin))
    rx.set_masks(masks)
    lib, h = capi.lib(), db._h
    for _ in range(args.warmup):
        capi.check(lib.apsu_b200_compute_powers(h))
        capi.check(lib.apsu_b200_eval_all(h))
    capi.check(lib.apsu_b200_ctx_synchronize(h))
    torch.cuda.profiler.start()
    capi.check(lib.apsu_b200_compute_powers(h))
    capi.check(lib.apsu_b200_eval_all(h))
    capi.check(lib.apsu_b200_ctx_synchronize(h))
    torch.cuda.profiler.stop()
    print(json.dumps(rx.timings()))
    db.close()


if __name__ == "__main__":
    main()
