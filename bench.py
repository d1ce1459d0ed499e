#!/usr/bin/env python3
"""bench.py — receiver query evaluation (ComputePowers + eval/eval_patstock over every BinBundle).

Metric (BASELINE.json): BinBundles/s and query-eval ms for parameters/16M-4096.json on synthetic sets
of the named size (receiver set 2^24), at N = 1/2/4/8 B200 (BinBundles sharded, strong scaling).

  python bench.py --gpus 1 --steps K --warmup W          # this framework, one JSON line
  python bench.py --impl reference ...                   # the CPU restatement of the reference path
  torchrun ... bench.py --gpus N ...                     # one rank per GPU

A "step" is one query: powers of the query ciphertexts for every bundle index, then the polynomial
evaluation of every BinBundle down to the result ciphertexts.  `value` times that with query, keys and
masks resident in HBM; `e2e` times the C-ABI call apsu_b200_run_query with pinned HOST buffers (H2D of the
query/keys/masks and D2H of the results inside the timed region, plus the NCCL broadcast/gather for N>1).
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = "16M-4096"
DB_LOG2 = 24
SEEDS = dict(db=0xD8, query=0x51, key=0x4B, mask=0x4D)


# ------------------------------------------------------------------------------------------------
# workload: bundle shapes of a 2^24-item receiver DB under parameters/16M-4096.json
# ------------------------------------------------------------------------------------------------
def load_params_json(name: str) -> dict:
    table = json.loads((ROOT / "tests" / "golden" / "parameters.json").read_text())
    return table[name + ".json"]


def simulate_bundle_degrees(pj: dict, db_log2: int, seed: int):
    """Degrees (max bin load) of every BinBundle after inserting 2^db_log2 items with all hash functions
    (receiver/apsu/receiver_db.cpp:70-79, 280-288) and first-fit over bundles with capacity
    max_items_per_bin-1 per bin (receiver_db.cpp:370-433).  Items land uniformly on table slots; an item's
    felts occupy the felts_per_item bins of its slot, so the bins of one slot always carry the same load."""
    tp, ip, sp = pj["table_params"], pj["item_params"], pj["seal_params"]
    N = sp["poly_modulus_degree"]
    items_per_bundle = N // ip["felts_per_item"]
    bic = tp["table_size"] // items_per_bundle
    cap = tp["max_items_per_bin"] - 1
    rng = np.random.default_rng(seed)
    loads = rng.multinomial((1 << db_log2) * tp["hash_func_count"], np.full(tp["table_size"], 1.0 / tp["table_size"]))
    degrees = []
    for b in range(bic):
        slot_loads = loads[b * items_per_bundle:(b + 1) * items_per_bundle]
        n_bundles = int(-(-slot_loads.max() // cap))
        degrees.append([int(np.clip(slot_loads - c * cap, 0, cap).max()) for c in range(n_bundles)])
    return degrees


def bundle_bytes(pj: dict, ncoeffs: int, low_L: int) -> int:
    ps = pj["query_params"]["ps_low_degree"]
    N = pj["seal_params"]["poly_modulus_degree"]
    n_plain = (ncoeffs + ps) // (ps + 1) if ps else 1
    return ((ncoeffs - n_plain) * low_L + n_plain) * N * 8


def shard(degrees, world: int):
    from apsu_b200.sharding import shard_bundles
    return shard_bundles(degrees, world)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every ~10 ms through NVML while the timed region runs."""

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.stop_flag, self.th, self.err = gpu_index, [], False, None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.idx).uuid)
            except Exception:
                pass
            self.h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        self.h = nv.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        continue
            if self.h is None:
                self.h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(0.01)

    def stop(self) -> dict:
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=1)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + str(self.err)]}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(n for n, bit in names.items() if any(r & bit for _, r in self.rows))
        return {"sm_mhz": float(np.median([s for s, _ in self.rows])), "sm_max_mhz": float(self.max_sm), "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# synthetic query material (uniform residues: timing is data independent; the same arrays feed the
# oracle for the CPU baseline and the bit-exact spot check)
# ------------------------------------------------------------------------------------------------
def synth_query(primes, t, N, first_L, K, nsrc, bic, npack, seed):
    rng = np.random.default_rng(seed)
    cts = np.zeros((nsrc, bic, 2, first_L, N), dtype=np.uint64)
    for j in range(first_L):
        cts[:, :, :, j, :] = rng.integers(0, primes[j], size=(nsrc, bic, 2, N), dtype=np.uint64)
    relin = np.zeros((K - 1, 2, K, N), dtype=np.uint64)
    for j in range(K):
        relin[:, :, j, :] = rng.integers(0, primes[j], size=(K - 1, 2, N), dtype=np.uint64)
    masks = rng.integers(0, t, size=(npack, N), dtype=np.uint64)
    return cts, relin, masks


def db_build_measure(pj, name, params, device, with_cpu):
    """Row f1 (BinBundle::regen_cache, bin_bundle.cpp:934-1041) on the device: one FULL BinBundle (every bin
    between half and completely full) built from its raw bins — polyn_with_roots, encode, NTT, device store —
    timed through the C ABI with host inputs; beside it the oracle (one thread, like one regen_cache task) on a
    bundle with 1/16 of the bins (polyn_with_roots is per bin, so its cost is linear in the bins)."""
    import apsu_b200
    rng = np.random.default_rng(SEEDS["db"] + 7)
    nb, cap, t = params.bins_per_bundle(), params.table_params()["max_items_per_bin"] - 1, params.plain_modulus()
    loads = rng.integers(cap // 2, cap + 1, size=nb)
    loads[0] = cap
    bins = [rng.integers(0, t, size=int(k), dtype=np.uint64) for k in loads]
    sizes = np.ascontiguousarray(loads, dtype=np.uint32)
    roots = np.ascontiguousarray(np.concatenate(bins))
    import ctypes as C
    from apsu_b200 import capi
    db = apsu_b200.ReceiverDB(params, device)
    try:
        ci = C.c_uint32()
        ms = []
        for _ in range(3):  # successive BinBundles of one DB build (the build scratch is reused after the first)
            t0 = time.perf_counter()
            capi.check(capi.lib().apsu_b200_db_add_binbundle_from_bins(db._h, 0, sizes, roots, C.byref(ci)))
            ms.append((time.perf_counter() - t0) * 1e3)
    finally:
        db.close()
    out = {"what": "one full BinBundle from raw bins on the device (apsu_b200_db_add_binbundle_from_bins, host inputs)",
           "items": int(loads.sum()), "bins": int(nb), "ms": float(min(ms)), "items_per_s": float(loads.sum() / (min(ms) / 1e3))}
    if with_cpu:
        from oracle import oracle as O
        p = O.Params(pj, name + ".json")
        odb = O.ReceiverDB(O.Context.from_params(p), p)
        sub = max(1, nb // 16)
        small = [b.tolist() if i < sub else [] for i, b in enumerate(bins)]
        small[sub - 1] = bins[0].tolist()  # keep the full degree so that every plaintext is built
        t0 = time.perf_counter()
        odb.add_bundle_from_bins(0, small)
        cms = (time.perf_counter() - t0) * 1e3
        n_items = sum(len(b) for b in small)
        out["cpu_port"] = {"what": f"oracle, 1 thread, same bundle with {sub} of {nb} bins populated", "items": int(n_items), "ms": float(cms),
                           "items_per_s": float(n_items / (cms / 1e3))}
    return out


def full_db_build_measure(pj, params, device, db_log2):
    """Row f1, whole DB: ReceiverDB::set_data (first-fit insertion, receiver_db.cpp:330-438, + every BinBundle cache) on
    the device from the algebraised items: 2^db_log2 items x hash_func_count table locations in the reference's own order
    (location-major, preprocess_unlabeled_data :292-301), uniform felts.  The items are generated on the device (torch)
    and handed over resident (apsu_b200_db_set_data_device): the time is the build, not the upload of 2.4 GB."""
    import ctypes as C
    import torch
    import apsu_b200
    from apsu_b200 import capi
    tp, F = pj["table_params"], pj["item_params"]["felts_per_item"]
    n = (1 << db_log2) * tp["hash_func_count"]
    g = torch.Generator(device="cuda")
    g.manual_seed(SEEDS["db"] + 11)
    locs = torch.sort(torch.randint(0, tp["table_size"], (n,), device="cuda", generator=g, dtype=torch.int64)).values
    cidx = (locs * F).contiguous()
    felts = torch.randint(0, params.plain_modulus(), (n, F), device="cuda", generator=g, dtype=torch.int64).contiguous()
    torch.cuda.synchronize()
    db = apsu_b200.ReceiverDB(params, device)
    try:
        counts = np.zeros(params.bundle_idx_count(), dtype=np.uint32)
        ms = []
        for _ in range(2):  # the second build first frees the 28 BinBundles of the first (cudaFree: 20-150 ms), so the minimum is reported
            t0 = time.perf_counter()
            capi.check(capi.lib().apsu_b200_db_set_data_device(db._h, C.c_void_p(felts.data_ptr()), C.c_void_p(cidx.data_ptr()), n, capi.ptr(counts)))
            ms.append((time.perf_counter() - t0) * 1e3)
        nb = int(counts.sum())
    finally:
        db.close()
    del felts, cidx, locs
    torch.cuda.empty_cache()
    return {"what": "whole DB from the algebraised items on the device (apsu_b200_db_set_data_device: sort by slot, first-fit windows, "
                    "polyn_with_roots + encode + NTT of every BinBundle), items resident",
            "items": int(n), "bin_bundles": nb, "per_bundle_index": [int(x) for x in counts], "ms": float(min(ms)), "ms_runs": [float(v) for v in ms],
            "items_per_s": float(n / (min(ms) / 1e3))}


def oracle_results(pj, name, degrees, cts, relin, masks, pairs, threads):
    """Result ciphertexts of the BinBundles `pairs` = [(bundle_idx, cache_idx), ...] of the synthetic DB, computed by
    the CPU oracle (same seeds as the GPU DB): the checker of `parity_sample`.  -> {pair: ndarray [2][N]}"""
    from oracle import oracle as O
    p = O.Params(pj, name + ".json")
    ctx = O.Context.from_params(p)
    bic = p.bundle_idx_count
    db = O.ReceiverDB(ctx, p)
    local = {}
    for (b, c) in pairs:
        local[(b, c)] = db.add_bundle_synthetic(b, degrees[b][c] + 1, SEEDS["db"] * 1000 + b * 64 + c)
    alpha = max(db.bundle_count(b) for b in range(bic))
    m = np.zeros((alpha * bic, p.N), dtype=np.uint64)
    for (b, c), lc in local.items():
        m[b + lc * bic] = masks[b + c * bic]
    ses = db.run_query(p.query_powers, cts, relin, m, threads=threads)
    inv = {(b, lc): (b, c) for (b, c), lc in local.items()}
    return {inv[(bb, lc)]: ct for (bb, lc, ct) in ses.results()}


def cpu_baseline(pj, name, degrees, cts, relin, masks, threads):
    """Times the oracle (CPU restatement of the reference's SEAL path, oracle/) on a bounded sample of the workload,
    once with -t 1 and once with -t `threads` (all host threads), like the reference's thread pool: ComputePowers for
    ONE bundle index (the reference runs bundle indices serially with -t workers over the DAG,
    receiver_ddh.cpp:325-333) + the evaluation of t full BinBundles in parallel, one per worker as in
    receiver_ddh.cpp:340-364.  The whole query is extrapolated linearly in bundle indices and plaintext count
    (a non-extrapolated full query is recorded once in profiles/, tools/cpu_full_query.py)."""
    from oracle import oracle as O
    p = O.Params(pj, name + ".json")
    ctx = O.Context.from_params(p)
    bic = p.bundle_idx_count
    full = p.max_items_per_bin
    b0 = max(range(bic), key=lambda b: len(degrees[b]))
    total_coeffs = sum(d + 1 for row in degrees for d in row)
    n_active = sum(1 for row in degrees if row)
    n_bundles = sum(len(r) for r in degrees)
    runs = {}
    for t in sorted({1, max(threads, 1)}):
        db = O.ReceiverDB(ctx, p)
        timed = [db.add_bundle_synthetic(b0, full, 999000 + k) for k in range(t)]
        ses = db.run_query(p.query_powers, cts, relin, None, threads=t, powers_only=True)
        m = np.zeros((len(timed) * bic, p.N), dtype=np.uint64)
        m[:] = masks[b0]
        eval_ms = ses.eval_subset([(b0, lc) for lc in timed], relin, m, threads=t)
        est_ms = ses.powers_ms * n_active + eval_ms * total_coeffs / (len(timed) * full)
        runs[t] = {"threads": t, "compute_powers_one_index_ms": ses.powers_ms, "eval_ms": eval_ms, "bundles_evaluated": len(timed),
                   "query_eval_ms_est": est_ms, "value": n_bundles / (est_ms / 1e3)}
        del ses, db
    best = runs[max(runs)]
    cpu_model = ""
    try:
        cpu_model = next(l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name"))
    except Exception:  # noqa: BLE001
        pass
    return {
        "value": best["value"], "unit": "BinBundles/s", "cores": best["threads"], "kind": "port",
        "sample": f"oracle (SEAL-algorithm restatement, not SEAL; built on this host with -O3 -march=native), -t {best['threads']}: ComputePowers "
                  f"for 1 of {n_active} bundle indices ({best['compute_powers_one_index_ms']:.0f} ms) + eval_patstock of {best['bundles_evaluated']} full "
                  f"BinBundles in parallel ({best['eval_ms']:.0f} ms); extrapolated linearly to the whole query = {best['query_eval_ms_est']:.0f} ms",
        "query_eval_ms_est": best["query_eval_ms_est"], "host_cores": os.cpu_count(), "cpu_model": cpu_model,
        "t1": runs[1], "t_all": best,
    }


def results_digest(out, bidx, cidx, n):
    """sha256 over every result ciphertext in (bundle_idx, cache_idx) order: the same synthetic DB, query and masks give
    the same digest at every GPU count (profiles/results_sha256_*.json holds the N=1 record)."""
    import hashlib
    order = sorted(range(n), key=lambda k: (int(bidx[k]), int(cidx[k])))
    h = hashlib.sha256()
    for k in order:
        h.update(np.array([bidx[k], cidx[k]], dtype=np.uint32).tobytes())
        h.update(np.ascontiguousarray(out[k]).tobytes())
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="apsu_b200", choices=["apsu_b200", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--db-log2", type=int, default=DB_LOG2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the bit-exact comparison of sampled BinBundles with the oracle")
    ap.add_argument("--no-db-build", action="store_true", help="skip the device DB-build measurement (row f1)")
    ap.add_argument("--no-dag-split", action="store_true", help="ranks sharing a bundle index recompute its powers instead of splitting the PowersDag")
    ap.add_argument("--dag-split", action="store_true", help="split the PowersDag whenever ranks share a bundle index (default: only for large DAGs)")
    ap.add_argument("--chunk", type=int, default=None, help="BinBundles per evaluation chunk (APSU_B200_CHUNK)")
    ap.add_argument("--write-digest", action="store_true", help="record the result digest of this run under profiles/")
    args = ap.parse_args()
    if args.chunk:
        os.environ["APSU_B200_CHUNK"] = str(args.chunk)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = args.workload
    pj = load_params_json(name)
    degrees = simulate_bundle_degrees(pj, args.db_log2, SEEDS["db"])
    n_bundles = sum(len(r) for r in degrees)
    config = {
        "workload": f"parameters/{name}.json, receiver set 2^{args.db_log2} (synthetic), {n_bundles} BinBundles "
                    f"{[len(r) for r in degrees]} per bundle index, degrees {degrees}",
        "l2": "DB plaintext stream per query is larger than L2 (no flush needed)",
        "seeds": SEEDS, "parallelism": f"BinBundles sharded over {world} GPU(s); query powers recomputed per owning GPU",
    }

    # ---------------- reference arm: the CPU restatement, all host threads ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import oracle as O
        O.build()  # compiled on THIS host (-march=native)
        p = O.Params(pj, name + ".json")
        threads = os.cpu_count() or 1
        cts, relin, masks = synth_query(p.primes, p.t, p.N, p.first_L, p.K, len(p.query_powers), p.bundle_idx_count,
                                        max(len(r) for r in degrees) * p.bundle_idx_count, SEEDS["query"])
        vals = []
        for _ in range(args.warmup + args.steps):
            vals.append(cpu_baseline(pj, name, degrees, cts, relin, masks, threads))
        vals = vals[args.warmup:] or vals
        ms = float(np.mean([v["query_eval_ms_est"] for v in vals]))
        info = vals[-1]
        info["value"] = n_bundles / (ms / 1e3)
        print(json.dumps({
            "impl": "reference", "metric": "receiver_query_eval_binbundles_per_s", "value": info["value"], "unit": "BinBundles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "cpu_baseline": info,
            "e2e": {"value": info["value"], "unit": "BinBundles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ---------------- this framework ----------------
    import torch
    import ctypes as C
    import apsu_b200
    from apsu_b200 import capi, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: apsu_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    params = apsu_b200.PSUParams.Load(json.dumps(pj))
    N, t, primes = params.poly_modulus_degree(), params.plain_modulus(), params.coeff_modulus()
    K = len(primes)
    bic = params.bundle_idx_count()
    db = apsu_b200.ReceiverDB(params, local_rank)
    rx = apsu_b200.Receiver(db)
    first_L, low_L = db.level(0), db.level(1)
    lib = capi.lib()
    h = db._h

    # this rank's shard; local cache indices are dense per bundle index
    parts = shard(degrees, world)
    mine = parts[rank]
    local_of = {}
    for (b, c, d) in mine:
        local_of[(b, c)] = db.add_bin_bundle_synthetic(b, d + 1, SEEDS["db"] * 1000 + b * 64 + c)
    my_bytes = db.stream_bytes()
    total_bytes = sum(bundle_bytes(pj, d + 1, low_L) for row in degrees for d in row)

    # multi-GPU: the C++ path (csrc/mgpu.cu) over its own NCCL communicator; torch.distributed only carries the
    # communicator id, the barriers and the max-over-ranks of the timings
    mg, mg_info = None, {"dag_group_size": 1}
    if world > 1:
        box = [sharding.MultiGpu.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        mg = sharding.MultiGpu(db, box[0], rank, world)
        mg_info = mg.commit([c for (_, c, _) in mine], 0 if args.no_dag_split else (1 if args.dag_split else -1))
        if mg_info["dag_group_size"] > 1:
            config["parallelism"] = (f"BinBundles sharded over {world} GPU(s); PowersDag of a bundle index split over the "
                                     f"{mg_info['dag_group_size']} rank(s) that share it, levels exchanged by {mg_info['dag_exchange']}")
        config["multi_gpu"] = f"C++ host path apsu_b200_mgpu_* over NCCL {mg_info['nccl_version']}: query scattered by bundle index, results gathered unpadded"

    nsrc = len(params.query_powers())
    src_powers = np.array(params.query_powers(), dtype=np.uint32)
    npack_global = max(len(r) for r in degrees) * bic
    cts, relin, masks = synth_query(primes, t, N, first_L, K, nsrc, bic, npack_global, SEEDS["query"])
    alpha_local = (max(local_of.values()) + 1) if local_of else 1
    masks_local = np.zeros((alpha_local * bic, N), dtype=np.uint64)
    for (b, c), lc in local_of.items():
        masks_local[b + lc * bic] = masks[b + c * bic]

    # pinned host staging for the e2e path
    def pinned(a):
        tns = torch.from_numpy(a.view(np.int64)).pin_memory()
        return tns, tns.numpy().view(np.uint64)
    cts_t, cts_p = pinned(cts)
    relin_t, relin_p = pinned(relin)
    masks_t, masks_p = pinned(masks_local)
    n_out = n_bundles if (rank == 0 or world == 1) else 1
    out_t = torch.empty((max(n_out, 1), 2, N), dtype=torch.int64).pin_memory()
    out_p = out_t.numpy().view(np.uint64)
    bidx = np.zeros(max(n_out, 1), dtype=np.uint32)
    cidx = np.zeros(max(n_out, 1), dtype=np.uint32)

    stream_ptr = rx.stream()
    stream = torch.cuda.ExternalStream(stream_ptr)  # the context's own stream: timing events are recorded on it

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident path: inputs in HBM before the timed region ----
    capi.check(lib.apsu_b200_query_begin(h, src_powers, nsrc, cts_p.reshape(-1)))
    capi.check(lib.apsu_b200_set_relin_keys(h, capi.ptr(relin_p)))
    capi.check(lib.apsu_b200_set_masks(h, masks_p.reshape(-1), masks_p.shape[0]))
    rx.set_profiling(True)

    def step_resident():
        if mg is not None:
            mg.compute_powers()
        else:
            capi.check(lib.apsu_b200_compute_powers(h))
        capi.check(lib.apsu_b200_eval_all(h))

    if mine:
        for _ in range(args.warmup):
            step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # the barrier comes AFTER the sampler start (NVML initialisation takes a different time on every rank): ranks that
    # split a PowersDag wait for each other inside the step, so a start skew would be billed to whoever started first
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if mine:
        for _ in range(args.steps):
            step_resident()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_step_local = e0.elapsed_time(e1) / args.steps
    tm = rx.timings()  # last step: scopes + DB-stream kernel launches timed live with CUDA events

    # ---- e2e path: host buffers through the C ABI; N>1: the C++ multi-GPU path (scatter / gather over NCCL inside) ----
    def step_e2e():
        if mg is None:
            capi.check(lib.apsu_b200_run_query(h, src_powers, nsrc, capi.ptr(cts_p), capi.ptr(relin_p), capi.ptr(masks_p),
                                               masks_p.shape[0], capi.ptr(out_p), capi.ptr(bidx), capi.ptr(cidx)))
        else:
            mg.run_query(src_powers, cts_p if rank == 0 else None, relin_p if rank == 0 else None, masks_p, out_p, bidx, cidx)

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()  # returns with the results on the (root's) host
    e2e_ms_local = (time.perf_counter() - w0) * 1e3 / args.steps
    barrier()

    # ---- e2e at N>1 when every rank can read the query in host memory (one shared segment / threads of one process; here
    # every process holds an identical pinned copy): each rank uploads its own part over its own PCIe link, no scatter ----
    e2e_shared_local = 0.0
    if mg is not None:
        def step_shared():
            mg.run_query(src_powers, cts_p, relin_p, masks_p, out_p, bidx, cidx, shared=True)
        for _ in range(args.warmup):
            step_shared()
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step_shared()
        e2e_shared_local = (time.perf_counter() - w0) * 1e3 / args.steps
        barrier()

    # ---- e2e at N>1 as the C++ host runs it (apsu::receiver::MultiGpuReceiver): shared query AND every rank delivers the
    # results of its own BinBundles to its own pinned host buffer (no gather; the D2H copies of all GPUs run in parallel) ----
    e2e_local_results = 0.0
    if mg is not None:
        n_loc = mg.local_count()
        loc_t = torch.empty((max(n_loc, 1), 2, N), dtype=torch.int64).pin_memory()
        loc_p = loc_t.numpy().view(np.uint64)
        lb, lc = np.zeros(max(n_loc, 1), dtype=np.uint32), np.zeros(max(n_loc, 1), dtype=np.uint32)

        def step_local():
            mg.run_query_local(src_powers, cts_p, relin_p, masks_p, loc_p, lb, lc)
        for _ in range(args.warmup):
            step_local()
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step_local()
        e2e_local_results = (time.perf_counter() - w0) * 1e3 / args.steps
        barrier()
        # the locally delivered ciphertexts must be the ones the gathered call returned (checked below against the oracle)
        step_shared()
        import hashlib
        loc_digests = [None] * world
        dist.all_gather_object(loc_digests, {(int(lb[k]), int(lc[k])): hashlib.sha256(loc_p[k].tobytes()).hexdigest() for k in range(n_loc)})
        local_ok = True
        if rank == 0:
            merged = {}
            for d_ in loc_digests:
                merged.update(d_)
            local_ok = len(merged) == n_bundles and all(
                merged.get((int(bidx[k]), int(cidx[k]))) == hashlib.sha256(out_p[k].tobytes()).hexdigest() for k in range(n_bundles))

    # ---- e2e, fed as the reference's RunQuery is (N=1): the query as on the wire (seeded ciphertexts and keys: c1 is a
    # 64-byte seed expanded on the device, row f2) and the masks drawn on the device inside the call (row f3) ----
    e2e_seeded = None
    if mg is None:
        rng = np.random.default_rng(SEEDS["query"] + 1)
        c0_t, c0_p = pinned(np.ascontiguousarray(cts[:, :, 0]))
        sd_t = torch.from_numpy(rng.integers(0, 256, size=(nsrc, bic, 64), dtype=np.uint8)).pin_memory()
        rc0_t, rc0_p = pinned(np.ascontiguousarray(relin[:, 0]))
        rsd_t = torch.from_numpy(rng.integers(0, 256, size=(K - 1, 64), dtype=np.uint8)).pin_memory()
        mask_seed = np.frombuffer(bytes(range(64)), dtype=np.uint8).copy()
        counts_b = [len(r) for r in degrees]
        alpha = max(counts_b)
        padded = np.ascontiguousarray([1 if c >= counts_b[b] else 0 for c in range(alpha) for b in range(bic)], dtype=np.uint8)
        rm_t = torch.empty((alpha * bic, params.items_per_bundle(), 2), dtype=torch.int64).pin_memory()

        def step_seeded():
            capi.check(lib.apsu_b200_run_query_seeded(
                h, src_powers, nsrc, capi.ptr(c0_p), C.c_void_p(sd_t.data_ptr()), capi.ptr(rc0_p), C.c_void_p(rsd_t.data_ptr()), capi.ptr(mask_seed),
                capi.ptr(padded), alpha * bic, C.c_void_p(rm_t.data_ptr()), capi.ptr(out_p), capi.ptr(bidx), capi.ptr(cidx)))
        saved = out_p[:n_bundles].copy(), bidx.copy(), cidx.copy()  # results of the expanded query: the parity sample below
        for _ in range(args.warmup):
            step_seeded()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step_seeded()
        ms_seeded = (time.perf_counter() - w0) * 1e3 / args.steps
        out_p[:n_bundles], bidx[:], cidx[:] = saved
        e2e_seeded = {"value": n_bundles / (ms_seeded / 1e3), "unit": "BinBundles/s", "ms_per_step": ms_seeded,
                      "h2d_bytes_per_step": int(c0_p.nbytes + sd_t.numel() + rc0_p.nbytes + rsd_t.numel() + 64 + padded.nbytes),
                      "d2h_bytes_per_step": int(n_bundles * 2 * N * 8 + rm_t.numel() * 8),
                      "what": "apsu_b200_run_query_seeded: seeded query ciphertexts and relinearisation keys as on the wire (c1 expanded on the device), "
                              "masks and PEQT blocks generated on the device (blake2xb), results + random_matrix back"}

    # per-rank view (diagnostics): every rank's own loop time and scopes
    per_rank = None
    if dist is not None:
        mine_info = {"rank": rank, "bin_bundles": len(mine), "bundle_indices": sorted({b for (b, _, _) in mine}), "ms_per_step": ms_step_local,
                     "e2e_ms": e2e_ms_local, "compute_powers_ms": tm["compute_powers_ms"], "eval_ms": tm["eval_ms"]}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine_info)

    # max over ranks
    if dist is not None:
        v = torch.tensor([ms_step_local, e2e_ms_local, e2e_shared_local, e2e_local_results], device="cuda", dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms, e2e_shared_ms, e2e_local_ms = float(v[0]), float(v[1]), float(v[2]), float(v[3])
    else:
        ms_step, e2e_ms, e2e_shared_ms, e2e_local_ms = ms_step_local, e2e_ms_local, 0.0, 0.0

    if rank == 0:
        # N=1: local cache indices are the global ones (one rank holds everything, in order)
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        mac_gbs = (tm["db_stream_bytes"] / 1e9) / (tm["db_stream_ms"] / 1e3) if tm["db_stream_ms"] else 0.0
        # DRAM traffic of the same launch from the committed ncu --set full capture (profiles/), per launch
        traffic = None
        ts = ROOT / "profiles" / "ncu_r01_k_db_mac_kt_summary.json"
        if ts.exists() and world == 1 and name == WORKLOAD and args.db_log2 == DB_LOG2:
            tj = json.loads(ts.read_text())
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        launches = max(tm["db_stream_launches"], 1)
        out = {
            "metric": "receiver_query_eval_binbundles_per_s", "value": n_bundles / (ms_step / 1e3), "unit": "BinBundles/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "query_eval_ms": ms_step,
            "scopes_ms_rank0_last_step": {"Receiver::ComputePowers": tm["compute_powers_ms"], "Receiver::ProcessBinBundleCache(all)": tm["eval_ms"]},
            "db_stream": {"bytes_per_query": total_bytes, "bytes_rank0": my_bytes, "effective_gbs_whole_eval": (my_bytes / 1e9) / (tm["eval_ms"] / 1e3) if tm["eval_ms"] else None},
            "roofline": {
                "kernel": "k_db_mac_kt (DB-stream multiply-accumulate, K1)", "bound": "hbm", "achieved": mac_gbs, "peak": hbm_peak,
                "unit": "GB/s", "frac": mac_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "launches_timed": tm["db_stream_launches"], "avg_launch_ms": tm["db_stream_ms"] / launches,
                "algorithmic_bytes_per_launch": tm["db_stream_bytes"] / launches,
                "share_of_step": tm["db_stream_ms"] / (tm["compute_powers_ms"] + tm["eval_ms"]) if tm["eval_ms"] else None,
            },
            "clocks": clocks,
            "e2e": {"value": n_bundles / (e2e_ms / 1e3), "unit": "BinBundles/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(cts.nbytes + relin.nbytes + masks.nbytes),
                    "d2h_bytes_per_step": int(n_bundles * 2 * N * 8),
                    "what": "apsu_b200_run_query"},
            "gpu_launches": int(tm["kernel_launches"]) * args.steps,
            "launch_mode": "eager" if os.environ.get("APSU_B200_NO_GRAPH", "0") not in ("", "0") else
                           f"CUDA graphs replayed per query ({int(tm['kernel_launches'])} kernel nodes of this repo's kernels per query on rank 0)",
        }
        if world == 1:
            # the integer-pipe side of the path: the transforms (49 % of the query) against the multiplier bound of their
            # butterfly (436 quarter-rate + 494 half-rate multiplier instructions per 68 butterflies
            # = 57 ns per N = 8192 polynomial on 148 SMs at 1965 MHz)
            ntt = {}
            for inv, key in ((0, "forward"), (1, "inverse")):
                ms = C.c_float()
                capi.check(lib.apsu_b200_bench_ntt(h, 2960, 20, inv, C.byref(ms)))
                ntt[key] = ms.value * 1e6 / 2960
            logn = N.bit_length() - 1
            bound_ns = (N // 2 * logn) / 32 * (2732 / 68) / 4 / 148 / 1.965  # warp-butterflies x pipe cycles / sub-partitions / SMs / GHz
            out["roofline_int"] = {
                "kernel": "ntt_kernel (negacyclic NTT / iNTT, one RNS prime)", "bound": "int32 multiplier pipe (IMAD.WIDE quarter rate)",
                "unit": "ns per polynomial", "achieved_forward": ntt["forward"], "achieved_inverse": ntt["inverse"], "bound_ns": bound_ns,
                "frac": bound_ns / max(ntt["forward"], ntt["inverse"]),
                "effective_gbs": 16 * N / min(ntt["forward"], ntt["inverse"]), "batch": "2960 polynomials per launch, 20 launches",
                "source": "instruction mix from cuobjdump -sass (DESIGN.md §4); ncu: fmaheavy 65-71 % busy (profiles/ncu_r01_ntt_final_raw.csv); "
                          "k_ks_mac / k_ks_moddown: profiles/ncu_r02_keyswitch_summary.json"}
        if per_rank is not None:
            out["per_rank"] = per_rank
        if e2e_seeded is not None:
            out["e2e_seeded"] = e2e_seeded
        if mg is not None:
            # N>1: the headline end-to-end number is the call a multi-GPU C++ host makes (apsu::receiver::MultiGpuReceiver:
            # one process, one thread per GPU, every GPU reads the query from the same host buffer): under torchrun every
            # rank process holds the query in its own pinned memory, which times the same thing.  The variant where ONE
            # process received the query and the library scatters it over NCCL is reported beside it.
            nb_keys = int(relin.nbytes)
            out["e2e_root_scatter"] = dict(out["e2e"], what="apsu_b200_mgpu_run_query: only rank 0 holds the query; it uploads it chunk by chunk (one PCIe "
                                           "link) and the library sends every rank the ciphertexts of its bundle indices (ncclSend/Recv) and the keys (ncclBroadcast); "
                                           "results gathered on rank 0")
            out["e2e_gathered"] = {
                "value": n_bundles / (e2e_shared_ms / 1e3), "unit": "BinBundles/s", "ms_per_step": e2e_shared_ms,
                "h2d_bytes_per_step": int(cts.nbytes + world * nb_keys + masks.nbytes), "d2h_bytes_per_step": int(n_bundles * 2 * N * 8),
                "what": "apsu_b200_mgpu_run_query_shared: every rank reads the query from host memory and uploads the ciphertexts of its own bundle "
                        "indices + the keys over its own PCIe link in parallel (pinned host buffers, inside the timed region); results gathered on rank 0 "
                        "over NCCL and copied to its host buffer"}
            out["e2e"] = {
                "value": n_bundles / (e2e_local_ms / 1e3), "unit": "BinBundles/s", "ms_per_step": e2e_local_ms,
                "h2d_bytes_per_step": int(cts.nbytes + world * nb_keys + masks.nbytes), "d2h_bytes_per_step": int(n_bundles * 2 * N * 8),
                "what": "apsu_b200_mgpu_run_query_local (what apsu::receiver::MultiGpuReceiver calls): every rank reads the query from host memory and "
                        "uploads the ciphertexts of its own bundle indices + the keys over its own PCIe link, and copies the result ciphertexts of its own "
                        "BinBundles to its own pinned host buffer (H2D and D2H inside the timed region, max over ranks; nothing gathered: each host thread "
                        "/ process forwards its ResultPackages as the reference's workers do)",
                "local_results_equal_gathered": bool(local_ok)}
            if not local_ok:
                out["INVALID"] = "locally delivered results differ from the gathered ones"
        # ---- parity at every N: the gathered results of the last e2e step against the oracle (one BinBundle per rank,
        # the fullest and the smallest at N=1) and their digest against the N=1 record ----
        got = {(int(bidx[k]), int(cidx[k])): out_p[k] for k in range(n_bundles)}
        digest = results_digest(out_p, bidx, cidx, n_bundles)
        out["results_sha256"] = digest
        rec = ROOT / "profiles" / f"results_sha256_{name}_2p{args.db_log2}.json"
        if args.write_digest and world == 1:
            rec.write_text(json.dumps({"workload": config["workload"], "seeds": SEEDS, "n_gpus": 1, "results_sha256": digest}, indent=1) + "\n")
        if rec.exists():
            want = json.loads(rec.read_text())["results_sha256"]
            out["results_sha256_matches_n1_record"] = bool(want == digest)
            if want != digest:
                out["INVALID"] = "result digest differs from the recorded N=1 run (profiles/" + rec.name + ")"
        if not args.no_parity:
            sample = []
            for part in parts:
                if part:
                    sample.append((part[0][0], part[0][1]))
            if world == 1:
                b0 = mine[0][0]
                small = [(b, c) for (b, c, d) in mine if b == b0][-1:]
                sample += [s_ for s_ in small if s_ not in sample]
            ref = oracle_results(pj, name, degrees, cts, relin, masks, sample, os.cpu_count() or 1)
            ok = all(np.array_equal(got[k].reshape(2, -1), ref[k]) for k in sample)
            out["parity_sample"] = {"bundles": sample, "bit_exact_vs_oracle": bool(ok),
                                    "what": "result ciphertexts gathered on rank 0 by the last end-to-end step vs the CPU oracle, one BinBundle per rank"}
            if not ok:
                out["INVALID"] = "GPU results differ from the oracle on the sampled BinBundles"
        if not args.no_cpu_baseline and world == 1:
            from oracle import oracle as O
            O.build()  # on this host
            out["cpu_baseline"] = cpu_baseline(pj, name, degrees, cts, relin, masks, os.cpu_count() or 1)
        if not args.no_db_build and world == 1:
            out["db_build"] = db_build_measure(pj, name, params, local_rank, not args.no_cpu_baseline)
            db.close()  # free the query DB: the full build below needs the memory of a second one
            try:
                out["db_build_full"] = full_db_build_measure(pj, params, local_rank, args.db_log2)
            except Exception as e:  # noqa: BLE001
                out["db_build_full"] = {"error": repr(e)}
        print(json.dumps(out))
    if mg is not None:
        mg.close()
    db.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
