#!/usr/bin/env python3
"""NTT micro-benchmark: ns per polynomial and effective GB/s (16*N bytes per transform).
usage: bench_ntt.py [workload] [count[,count...]]   (APSU_B200_NTT_SPLIT=0|2|4 forces the launch shape, context.cu)"""
import os
import ctypes as C, json, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import apsu_b200
from apsu_b200 import capi
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "16M-4096"
counts = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2960]
params = apsu_b200.PSUParams.Load(json.dumps(bench.load_params_json(name)))
db = apsu_b200.ReceiverDB(params, 0)
N = params.poly_modulus_degree()
for count in counts:
    out = {}
    for inv in (0, 1):
        ms = C.c_float()
        capi.check(capi.lib().apsu_b200_bench_ntt(db._h, count, 50, inv, C.byref(ms)))
        out["inverse" if inv else "forward"] = dict(us=round(ms.value * 1e3, 2), ns_per_poly=round(ms.value * 1e6 / count, 1), gbs=round(count * 16 * N / ms.value / 1e6))
    print(json.dumps(dict(workload=name, N=N, polys=count, split=os.environ.get("APSU_B200_NTT_SPLIT", "auto"), **out)), flush=True)
db.close()
