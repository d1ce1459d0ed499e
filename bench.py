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


def cpu_baseline(pj, name, degrees, cts, relin, masks, sample_pairs, threads):
    """Times the oracle (CPU restatement of the reference's SEAL path, oracle/) with all host threads busy on a
    bounded sample of the workload: ComputePowers for ONE bundle index (the reference runs bundle indices
    serially with -t workers over the DAG, receiver_ddh.cpp:325-333) + the evaluation of `threads` full
    BinBundles in parallel, one per worker as in receiver_ddh.cpp:340-364.  The whole query is extrapolated
    linearly in bundle indices and plaintext count.  `sample_pairs` (same seeds as the GPU DB) are also
    evaluated so their result ciphertexts can be compared bit for bit.  Returns (dict, {pair: ndarray})."""
    from oracle import oracle as O
    p = O.Params(pj, name + ".json")
    ctx = O.Context.from_params(p)
    bic = p.bundle_idx_count
    db = O.ReceiverDB(ctx, p)
    b0 = sample_pairs[0][0]
    # the oracle DB holds bundles at index b0 only (other indices empty => ComputePowers skipped there)
    local = {}
    for (b, c) in sample_pairs:
        assert b == b0
        local[(b, c)] = db.add_bundle_synthetic(b, degrees[b][c] + 1, SEEDS["db"] * 1000 + b * 64 + c)
    full = p.max_items_per_bin
    timed = [lc for (b, c), lc in local.items() if degrees[b][c] + 1 == full][:threads]
    while len(timed) < max(threads, 1):  # timing-only full-size bundles so that every worker has one
        timed.append(db.add_bundle_synthetic(b0, full, 999000 + len(timed)))
    ses = db.run_query(p.query_powers, cts, relin, None, threads=threads, powers_only=True)
    powers_ms_one = ses.powers_ms
    alpha = db.bundle_count(b0)
    m = np.zeros((alpha * bic, p.N), dtype=np.uint64)
    m[:] = masks[b0]
    for (b, c), lc in local.items():
        m[b + lc * bic] = masks[b + c * bic]
    eval_ms = ses.eval_subset([(b0, lc) for lc in timed], relin, m, threads=threads)
    rest = [(b0, lc) for lc in local.values() if lc not in timed]
    if rest:
        ses.eval_subset(rest, relin, m, threads=threads)
    res = {}
    for (bb, lc, ct) in ses.results():
        for (b, c), l2 in local.items():
            if (bb, lc) == (b, l2):
                res[(b, c)] = ct
    timed_coeffs = len(timed) * full
    total_coeffs = sum(d + 1 for row in degrees for d in row)
    n_active = sum(1 for row in degrees if row)
    est_ms = powers_ms_one * n_active + eval_ms * total_coeffs / timed_coeffs
    n_bundles = sum(len(r) for r in degrees)
    info = {
        "value": n_bundles / (est_ms / 1e3), "unit": "BinBundles/s", "cores": threads, "kind": "port",
        "sample": f"oracle (SEAL-algorithm restatement, not SEAL), -t {threads}: ComputePowers for 1 of {n_active} bundle indices "
                  f"({powers_ms_one:.0f} ms) + eval_patstock of {len(timed)} full BinBundles in parallel = {timed_coeffs} of {total_coeffs} "
                  f"plaintexts ({eval_ms:.0f} ms); extrapolated linearly to the whole query = {est_ms:.0f} ms",
        "query_eval_ms_est": est_ms, "host_cores": os.cpu_count(),
    }
    return info, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="apsu_b200", choices=["apsu_b200", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--db-log2", type=int, default=DB_LOG2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-db-build", action="store_true", help="skip the device DB-build measurement (row f1)")
    ap.add_argument("--no-dag-split", action="store_true", help="ranks sharing a bundle index recompute its powers instead of splitting the PowersDag")
    ap.add_argument("--dag-split", action="store_true", help="split the PowersDag whenever ranks share a bundle index (default: only for large DAGs)")
    ap.add_argument("--chunk", type=int, default=None, help="BinBundles per evaluation chunk (APSU_B200_CHUNK)")
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
        p = O.Params(pj, name + ".json")
        threads = os.cpu_count() or 1
        cts, relin, masks = synth_query(p.primes, p.t, p.N, p.first_L, p.K, len(p.query_powers), p.bundle_idx_count,
                                        max(len(r) for r in degrees) * p.bundle_idx_count, SEEDS["query"])
        b0 = 0
        full = [(b0, c) for c, d in enumerate(degrees[b0]) if d + 1 == p.max_items_per_bin][:1]
        small = [(b0, len(degrees[b0]) - 1)]
        sample = full + [s for s in small if s not in full]
        vals = []
        for _ in range(args.warmup + args.steps):
            info, _ = cpu_baseline(pj, name, degrees, cts, relin, masks, sample, threads)
            vals.append(info)
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
    from apsu_b200 import capi

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
    stream = torch.cuda.Stream()
    capi.check(capi.lib().apsu_b200_ctx_set_stream(db._h, C.c_void_p(stream.cuda_stream)))

    # this rank's shard; local cache indices are dense per bundle index
    parts = shard(degrees, world)
    mine = parts[rank]
    # more ranks than bundle indices: the ranks sharing an index split its PowersDag and all-gather the powers
    # level by level (collective C2) instead of each recomputing them
    part_group, part_index, part_pg = [rank], 0, None
    if world > 1:
        from apsu_b200 import sharding
        part_group, part_index, all_groups = sharding.powers_partition(parts, rank)
        dagp = apsu_b200.PowersDag(params)
        n_products = len(dagp.nodes) - dagp.source_count()
        if args.no_dag_split or not (args.dag_split or sharding.worth_splitting(n_products, len(part_group))):
            part_group, part_index, all_groups = [rank], 0, []
        if any(len(g) > 1 for g in all_groups):
            for g in all_groups:  # every rank creates every group
                pg = dist.new_group(g)
                if g == part_group:
                    part_pg = pg
        if len(part_group) > 1:
            rx.set_powers_partition(part_index, len(part_group))
        if len(part_group) > 1:
            config["parallelism"] = (f"BinBundles sharded over {world} GPU(s); PowersDag of a bundle index split over the "
                                     f"{len(part_group)} rank(s) that share it, powers all-gathered per DAG level (NCCL)")
    local_of = {}
    for (b, c, d) in mine:
        local_of[(b, c)] = db.add_bin_bundle_synthetic(b, d + 1, SEEDS["db"] * 1000 + b * 64 + c)
    my_bytes = db.stream_bytes()
    total_bytes = sum(bundle_bytes(pj, d + 1, low_L) for row in degrees for d in row)

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
    n_local = len(mine)
    out_t = torch.empty((max(n_local, 1), 2, N), dtype=torch.int64).pin_memory()
    out_p = out_t.numpy().view(np.uint64)

    lib = capi.lib()
    h = db._h

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident path: inputs in HBM before the timed region ----
    with torch.cuda.stream(stream):
        capi.check(lib.apsu_b200_query_begin(h, src_powers, nsrc, cts_p.reshape(-1)))
        capi.check(lib.apsu_b200_set_relin_keys(h, capi.ptr(relin_p)))
        capi.check(lib.apsu_b200_set_masks(h, masks_p.reshape(-1), masks_p.shape[0]))
        rx.set_profiling(True)

        def compute_powers():
            if len(part_group) > 1:
                rx.ComputePowers(exchange=lambda level, regs: sharding.exchange_powers(regs, part_index, len(part_group), part_pg))
            else:
                capi.check(lib.apsu_b200_compute_powers(h))

        def step_resident():
            compute_powers()
            capi.check(lib.apsu_b200_eval_all(h))

        for _ in range(args.warmup):
            step_resident()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per_step = []
        e0.record(stream)
        for _ in range(args.steps):
            step_resident()
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        ms_total = e0.elapsed_time(e1)
        tm = rx.timings()  # last step: scopes + DB-stream kernel launches timed live with CUDA events
        ms_step_local = ms_total / args.steps

        # ---- e2e path: host buffers through the C ABI (+ NCCL broadcast / gather for N>1) ----
        bidx = np.zeros(max(n_local, 1), dtype=np.uint32)
        cidx = np.zeros(max(n_local, 1), dtype=np.uint32)
        if dist is not None:
            # one pinned host block and one device block for the whole query (ciphertexts + relinearisation keys):
            # one H2D copy on the rank that received it, one NCCL broadcast
            q_host = torch.empty(cts_t.numel() + relin_t.numel(), dtype=torch.int64).pin_memory()
            q_host[:cts_t.numel()] = cts_t.reshape(-1)
            q_host[cts_t.numel():] = relin_t.reshape(-1)
            d_query = torch.empty(q_host.shape, dtype=torch.int64, device="cuda")
            d_cts, d_relin = d_query[:cts_t.numel()], d_query[cts_t.numel():]
            counts = [len(x) for x in parts]
            gatherer = sharding.ResultGatherer(counts, N, torch.device("cuda", local_rank), dst=0)

        def step_e2e():
            if dist is None:
                capi.check(lib.apsu_b200_run_query(h, src_powers, nsrc, capi.ptr(cts_p), capi.ptr(relin_p), capi.ptr(masks_p),
                                                   masks_p.shape[0], capi.ptr(out_p), capi.ptr(bidx), capi.ptr(cidx)))
            else:
                # rank 0 holds the query on the host: H2D once, NCCL broadcast over NVLink, evaluate, gather, D2H
                if rank == 0:
                    d_query.copy_(q_host, non_blocking=True)
                sharding.broadcast_query([d_query], src=0)
                capi.check(lib.apsu_b200_query_begin_device(h, src_powers, nsrc, C.c_void_p(d_cts.data_ptr())))
                capi.check(lib.apsu_b200_set_relin_keys_device(h, C.c_void_p(d_relin.data_ptr())))
                capi.check(lib.apsu_b200_set_masks(h, masks_p.reshape(-1), masks_p.shape[0]))
                compute_powers()
                capi.check(lib.apsu_b200_eval_all(h))
                if n_local:
                    capi.check(lib.apsu_b200_copy_results_device(h, C.c_void_p(gatherer.local_buffer().data_ptr())))
                gatherer.gather()
                if rank == 0:
                    stream.synchronize()  # the results are on the host

        for _ in range(args.warmup):
            step_e2e()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        f1.record(stream)
        barrier()
        e2e_ms_local = max(f0.elapsed_time(f1), (time.perf_counter() - w0) * 1e3) / args.steps

    # max over ranks
    if dist is not None:
        v = torch.tensor([ms_step_local, e2e_ms_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms = float(v[0]), float(v[1])
    else:
        ms_step, e2e_ms = ms_step_local, e2e_ms_local

    if rank == 0:
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
                    "h2d_bytes_per_step": int(cts.nbytes + relin.nbytes + masks_local.nbytes),
                    "d2h_bytes_per_step": int(n_bundles * 2 * N * 8)},
            "gpu_launches": int(tm["kernel_launches"]) * args.steps,
            "launch_mode": "eager" if os.environ.get("APSU_B200_NO_GRAPH", "0") not in ("", "0") else
                           f"CUDA graphs replayed per query ({int(tm['kernel_launches'])} kernel nodes of this repo's kernels per query)",
        }
        if not args.no_cpu_baseline and world == 1:
            # bounded CPU sample of the same workload + bit-exact spot check of those BinBundles
            b0 = mine[0][0]
            full = [(b, c) for (b, c, d) in mine if b == b0 and d + 1 == pj["table_params"]["max_items_per_bin"]][:1]
            small = [(b, c) for (b, c, d) in mine if b == b0][-1:]
            sample = full + [s for s in small if s not in full]
            info, ref = cpu_baseline(pj, name, degrees, cts, relin, masks, sample, os.cpu_count() or 1)
            out["cpu_baseline"] = info
            got = {(r.bundle_idx, r.cache_idx): r.psu_result.reshape(2, -1) for r in rx.results()}
            ok = all(np.array_equal(got[(b, local_of[(b, c)])], ref[(b, c)]) for (b, c) in sample)
            out["parity_sample"] = {"bundles": sample, "bit_exact_vs_oracle": bool(ok)}
            if not ok:
                out["INVALID"] = "GPU results differ from the oracle on the sampled BinBundles"
        if not args.no_db_build and world == 1:
            out["db_build"] = db_build_measure(pj, name, params, local_rank, not args.no_cpu_baseline)
        print(json.dumps(out))
    db.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
