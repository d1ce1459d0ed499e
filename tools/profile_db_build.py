#!/usr/bin/env python3
"""One profiled device build of a full BinBundle from raw bins (row f1), for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/db_build.csv python tools/profile_db_build.py"""
import ctypes as C
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    import torch
    import apsu_b200
    from apsu_b200 import capi
    pj = bench.load_params_json(bench.WORKLOAD)
    params = apsu_b200.PSUParams.Load(json.dumps(pj))
    rng = np.random.default_rng(7)
    nb, cap, t = params.bins_per_bundle(), params.table_params()["max_items_per_bin"] - 1, params.plain_modulus()
    loads = rng.integers(cap // 2, cap + 1, size=nb)
    loads[0] = cap
    sizes = np.ascontiguousarray(loads, dtype=np.uint32)
    roots = np.ascontiguousarray(rng.integers(0, t, size=int(loads.sum()), dtype=np.uint64))
    db = apsu_b200.ReceiverDB(params, 0)
    ci = C.c_uint32()
    capi.check(capi.lib().apsu_b200_db_add_binbundle_from_bins(db._h, 0, sizes, roots, C.byref(ci)))
    capi.check(capi.lib().apsu_b200_db_clear(db._h))
    torch.cuda.profiler.start()
    t0 = time.perf_counter()
    capi.check(capi.lib().apsu_b200_db_add_binbundle_from_bins(db._h, 0, sizes, roots, C.byref(ci)))
    ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.profiler.stop()
    print(json.dumps({"items": int(loads.sum()), "ms": ms}))
    db.close()


if __name__ == "__main__":
    main()
