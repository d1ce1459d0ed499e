"""Host-side mirror of the reference's receiver surface for the query-evaluation path, over the C ABI.

Names follow the reference (common/apsu/psu_params.h, common/apsu/powers.h,
receiver/apsu/receiver_db.h, receiver/apsu/receiver_ddh.h, common/apsu/network/result_package.h);
the C++ facade with the same names lives in apsu_b200/host/apsu_b200.hpp.  This module is glue for
tests, bench.py and torch.distributed sharding — all arithmetic happens in libapsu_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import json
from dataclasses import dataclass

import numpy as np

from . import capi


class PSUParams:
    """common/apsu/psu_params.h:31-222 — parameters + derived bundle geometry."""

    def __init__(self, c: capi.CParams):
        self._c = c

    @staticmethod
    def Load(json_text: str) -> "PSUParams":
        """PSUParams::Load(const std::string&), common/apsu/psu_params.cpp:290-374"""
        c = capi.CParams()
        capi.check(capi.lib().apsu_b200_params_load_json(json_text.encode(), C.byref(c)))
        return PSUParams(c)

    @staticmethod
    def from_fields(N, plain_modulus, coeff_modulus, hash_func_count, table_size, max_items_per_bin, felts_per_item,
                    ps_low_degree, query_powers) -> "PSUParams":
        c = capi.CParams()
        c.poly_modulus_degree, c.plain_modulus = N, plain_modulus
        c.coeff_modulus_count = len(coeff_modulus)
        for i, q in enumerate(coeff_modulus):
            c.coeff_modulus[i] = q
        c.hash_func_count, c.table_size, c.max_items_per_bin = hash_func_count, table_size, max_items_per_bin
        c.felts_per_item, c.ps_low_degree = felts_per_item, ps_low_degree
        qp = sorted(set([1] + list(query_powers)))
        c.query_power_count = len(qp)
        for i, p in enumerate(qp):
            c.query_powers[i] = p
        capi.check(capi.lib().apsu_b200_params_validate(C.byref(c)))
        return PSUParams(c)

    # accessors named as in the reference
    def poly_modulus_degree(self): return self._c.poly_modulus_degree
    def plain_modulus(self): return self._c.plain_modulus
    def coeff_modulus(self): return [self._c.coeff_modulus[i] for i in range(self._c.coeff_modulus_count)]
    def table_params(self): return dict(hash_func_count=self._c.hash_func_count, table_size=self._c.table_size, max_items_per_bin=self._c.max_items_per_bin)
    def item_params(self): return dict(felts_per_item=self._c.felts_per_item)
    def query_params(self): return dict(ps_low_degree=self._c.ps_low_degree, query_powers=self.query_powers())
    def query_powers(self): return [self._c.query_powers[i] for i in range(self._c.query_power_count)]
    def item_bit_count(self): return self._c.item_bit_count
    def item_bit_count_per_felt(self): return self._c.item_bit_count_per_felt
    def items_per_bundle(self): return self._c.items_per_bundle
    def bins_per_bundle(self): return self._c.bins_per_bundle
    def bundle_idx_count(self): return self._c.bundle_idx_count


@dataclass
class PowersNode:
    power: int
    depth: int
    parents: tuple

    def is_source(self):
        return self.parents == (0, 0)


class PowersDag:
    """common/apsu/powers.h:41-293 (configure / depth / target_powers / source_nodes)."""

    def __init__(self, params: PSUParams):
        cap = params.table_params()["max_items_per_bin"] + 1
        a = [np.zeros(cap, dtype=np.uint32) for _ in range(4)]
        n, d = C.c_uint32(), C.c_uint32()
        capi.check(capi.lib().apsu_b200_powers_dag(C.byref(params._c), cap, *a, C.byref(n), C.byref(d)))
        self.nodes = {int(a[0][i]): PowersNode(int(a[0][i]), int(a[1][i]), (int(a[2][i]), int(a[3][i]))) for i in range(n.value)}
        self._depth = d.value

    def depth(self): return self._depth
    def target_powers(self): return sorted(self.nodes)
    def source_nodes(self): return [n for n in self.nodes.values() if n.is_source()]
    def source_count(self): return len(self.source_nodes())


@dataclass
class ResultPackage:
    """common/apsu/network/result_package.h:44-62 — one per BinBundle."""
    bundle_idx: int
    cache_idx: int
    psu_result: np.ndarray  # ciphertext [2][1][N] at the last level, coefficient form


class ReceiverDB:
    """Device-resident receiver DB: only the query-time surface of receiver/apsu/receiver_db.h:227-292
    plus the upload of BinBundle caches (batched_coeffs, receiver/apsu/bin_bundle.h:57)."""

    def __init__(self, params: PSUParams, device: int = 0):
        self.params = params
        h = C.c_void_p()
        capi.check(capi.lib().apsu_b200_ctx_create(C.byref(params._c), device, C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().apsu_b200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def get_params(self): return self.params

    def level(self, which: int) -> int:
        v = C.c_uint32()
        capi.check(capi.lib().apsu_b200_ctx_level(self._h, which, C.byref(v)))
        return v.value

    def add_bin_bundle(self, bundle_idx: int, batched_coeffs) -> int:
        """batched_coeffs: list of ndarrays, [low_L][N] for NTT-form degrees, [N] for coefficient-form ones."""
        arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in batched_coeffs]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        ci = C.c_uint32()
        capi.check(capi.lib().apsu_b200_db_add_binbundle(self._h, bundle_idx, ptrs, len(arrs), C.byref(ci)))
        return ci.value

    def add_bin_bundle_synthetic(self, bundle_idx: int, ncoeffs: int, seed: int) -> int:
        ci = C.c_uint32()
        capi.check(capi.lib().apsu_b200_db_add_binbundle_synthetic(self._h, bundle_idx, ncoeffs, seed, C.byref(ci)))
        return ci.value

    def add_bin_bundle_from_bins(self, bundle_idx: int, bins) -> int:
        """BinBundle::regen_cache on the device (receiver/apsu/bin_bundle.cpp:934-1041): `bins` is one list of
        field elements (the items hashed to that bin) per bin of the bundle; polyn_with_roots, BatchEncoder::encode
        and the NTTs run on the GPU and the cache stays device-resident."""
        sizes = np.ascontiguousarray([len(b) for b in bins], dtype=np.uint32)
        if len(sizes) != self.params.bins_per_bundle():
            raise ValueError("bins must hold one entry per bin of the bundle")
        roots = np.ascontiguousarray([x for b in bins for x in b] or [0], dtype=np.uint64)
        ci = C.c_uint32()
        capi.check(capi.lib().apsu_b200_db_add_binbundle_from_bins(self._h, bundle_idx, sizes, roots, C.byref(ci)))
        return ci.value

    def set_data(self, felts, cuckoo_idx):
        """ReceiverDB::set_data (receiver/apsu/receiver_db.cpp:966 -> insert_or_assign_worker :330-438 -> generate_caches
        :808) on the device: felts [n][felts_per_item] and cuckoo_idx [n] (= location * felts_per_item) are the
        reference's data_with_indices in its order; first-fit insertion into BinBundles and every cache are built on the
        GPU.  Returns the number of BinBundles per bundle index."""
        f = np.ascontiguousarray(felts, dtype=np.uint64)
        c = np.ascontiguousarray(cuckoo_idx, dtype=np.uint64)
        if f.ndim != 2 or f.shape[1] != self.params.item_params()["felts_per_item"] or f.shape[0] != c.shape[0]:
            raise ValueError("felts must be [n][felts_per_item] and cuckoo_idx [n]")
        counts = np.zeros(self.params.bundle_idx_count(), dtype=np.uint32)
        capi.check(capi.lib().apsu_b200_db_set_data(self._h, capi.ptr(f), capi.ptr(c), f.shape[0], capi.ptr(counts)))
        return [int(x) for x in counts]

    def get_bin_bundle_count(self, bundle_idx: int | None = None) -> int:
        v = C.c_uint32()
        if bundle_idx is None:
            capi.check(capi.lib().apsu_b200_db_total_bin_bundle_count(self._h, C.byref(v)))
        else:
            capi.check(capi.lib().apsu_b200_db_bin_bundle_count(self._h, bundle_idx, C.byref(v)))
        return v.value

    def bin_bundle_coeff(self, bundle_idx: int, cache_idx: int, k: int) -> np.ndarray:
        L = C.c_uint32()
        capi.check(capi.lib().apsu_b200_db_binbundle_coeff(self._h, bundle_idx, cache_idx, k, None, C.byref(L)))
        N = self.params.poly_modulus_degree()
        out = np.zeros((max(L.value, 1), N), dtype=np.uint64)
        capi.check(capi.lib().apsu_b200_db_binbundle_coeff(self._h, bundle_idx, cache_idx, k, capi.ptr(out), C.byref(L)))
        return out if L.value else out[0]

    def stream_bytes(self) -> int:
        v = C.c_uint64()
        capi.check(capi.lib().apsu_b200_db_stream_bytes(self._h, C.byref(v)))
        return v.value

    def clear(self):
        capi.check(capi.lib().apsu_b200_db_clear(self._h))


class Query:
    """receiver/apsu/query.h:26-89 — the already-deserialised query: source powers -> ciphertexts per
    bundle index (host arrays), plus relinearisation keys."""

    def __init__(self, src_powers, cts: np.ndarray, relin_keys: np.ndarray | None):
        self.src_powers = np.ascontiguousarray(list(src_powers), dtype=np.uint32)
        self.cts = np.ascontiguousarray(cts, dtype=np.uint64)  # [nsrc][bundle_idx_count][2][L][N]
        self.relin_keys = None if relin_keys is None else np.ascontiguousarray(relin_keys, dtype=np.uint64)


class Receiver:
    """receiver/apsu/receiver_ddh.h:184-243 — the HE part of RunQuery."""

    def __init__(self, db: ReceiverDB):
        self.db = db
        self._L = capi.lib()
        self._part_rank, self._part_size = 0, 1
        self._stage_regions = None

    # ---- staged API (ComputePowers / ProcessBinBundleCache) ----
    def load_query(self, query: Query):
        h = self.db._h
        capi.check(self._L.apsu_b200_query_begin(h, query.src_powers, len(query.src_powers), query.cts))
        if query.relin_keys is not None:
            capi.check(self._L.apsu_b200_set_relin_keys(h, capi.ptr(query.relin_keys)))

    def load_query_seeded(self, src_powers, c0: np.ndarray, seeds: np.ndarray, relin_c0: np.ndarray | None = None, relin_seeds: np.ndarray | None = None):
        """Row f2: the query as on the wire (seal::Serializable): c0 [nsrc][bundle_idx_count][L][N] + one 64-byte seed per
        ciphertext [nsrc][bundle_idx_count][64]; relinearisation keys c0 [K-1][K][N] + seeds [K-1][64]; the second
        polynomials are sampled on the device (sample_poly_uniform)."""
        sp = np.ascontiguousarray(list(src_powers), dtype=np.uint32)
        c0 = np.ascontiguousarray(c0, dtype=np.uint64)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint8)
        capi.check(self._L.apsu_b200_query_begin_seeded(self.db._h, sp, len(sp), capi.ptr(c0), capi.ptr(seeds)))
        if relin_c0 is not None:
            rc0 = np.ascontiguousarray(relin_c0, dtype=np.uint64)
            rs = np.ascontiguousarray(relin_seeds, dtype=np.uint8)
            capi.check(self._L.apsu_b200_set_relin_keys_seeded(self.db._h, capi.ptr(rc0), capi.ptr(rs)))

    def set_masks(self, masks: np.ndarray):
        masks = np.ascontiguousarray(masks, dtype=np.uint64)
        capi.check(self._L.apsu_b200_set_masks(self.db._h, masks, masks.shape[0]))

    def generate_masks(self, seed: bytes | None = None, cache_counts=None, want_values: bool = False):
        """receiver_ddh.cpp:218-283 on the device: draws the masks of every (cache_idx, bundle_idx) pair from SEAL's
        blake2xb generator keyed with the 64-byte `seed` (None: 64 bytes of OS randomness, as the reference does),
        keeps their encodings resident for the next evaluation and returns random_matrix
        [npack][items_per_bundle][2] (low, high words of the PEQT blocks; all ones for padded pairs).
        cache_counts[bundle_idx] defaults to the DB's."""
        bic = self.db.params.bundle_idx_count()
        if cache_counts is None:
            cache_counts = [self.db.get_bin_bundle_count(b) for b in range(bic)]
        alpha = max(max(cache_counts), 1)
        padded = np.ascontiguousarray([1 if c >= cache_counts[b] else 0 for c in range(alpha) for b in range(bic)], dtype=np.uint8)
        npack = alpha * bic
        blocks = np.zeros((npack, self.db.params.items_per_bundle(), 2), dtype=np.uint64)
        values = np.zeros((npack, self.db.params.poly_modulus_degree()), dtype=np.uint64) if want_values else None
        sd = None
        if seed is not None:
            if len(seed) != 64:
                raise ValueError("the mask seed is 64 bytes (seal::prng_seed_type)")
            sd = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
        capi.check(self._L.apsu_b200_generate_masks(self.db._h, capi.ptr(sd), capi.ptr(padded), npack, capi.ptr(blocks), capi.ptr(values)))
        return (blocks, values) if want_values else blocks

    def decrypt_results(self, secret_key_ntt_q0: np.ndarray, cts: np.ndarray, want_blocks: bool = True):
        """sender side (ResultPackage::extract, result_package.cpp:175-213 + sender_ddh.cpp:588-594) on the device:
        cts [n][2][N] -> (slot values [n][N], blocks [n][items_per_bundle][2] or None, noise budgets [n])."""
        cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, 2, self.db.params.poly_modulus_degree())
        sk = np.ascontiguousarray(secret_key_ntt_q0, dtype=np.uint64)
        n, N = cts.shape[0], cts.shape[2]
        values = np.zeros((n, N), dtype=np.uint64)
        blocks = np.zeros((n, self.db.params.items_per_bundle(), 2), dtype=np.uint64) if want_blocks else None
        budget = np.zeros(n, dtype=np.int32)
        capi.check(self._L.apsu_b200_decrypt_results(self.db._h, capi.ptr(sk), capi.ptr(cts), n, capi.ptr(values), capi.ptr(blocks), capi.ptr(budget)))
        return values, blocks, budget

    def encode_masks(self, slot_values: np.ndarray) -> np.ndarray:
        v = np.ascontiguousarray(slot_values, dtype=np.uint64)
        out = np.zeros_like(v)
        capi.check(self._L.apsu_b200_encode_masks(self.db._h, v, v.shape[0], out))
        return out

    def ComputePowers(self, exchange=None):
        """Receiver::ComputePowers for every bundle index.  With a split PowersDag (set_powers_partition) `exchange`
        is called after every DAG level as exchange(level, regions) with regions = [(device_ptr, chunk_bytes), ...]
        and must all-gather them between the ranks of the partition (apsu_b200/sharding.py::exchange_powers, which
        needs engine_stream=self.stream() to order its collective with the context's stream)."""
        if self._part_size == 1:
            capi.check(self._L.apsu_b200_compute_powers(self.db._h))
            return
        # the regions are fixed per (parameters, DB, partition); one probe per query notices a rebuilt plan
        if self._stage_regions is not None and self._stage_regions[0] != self.powers_exchange_regions(1):
            self._stage_regions = None
        if self._stage_regions is None:
            n = C.c_uint32()
            capi.check(self._L.apsu_b200_powers_stage_count(self.db._h, C.byref(n)))
            self._stage_regions = [self.powers_exchange_regions(s + 1) if s + 1 < n.value else None for s in range(n.value)]
        for s, regions in enumerate(self._stage_regions):
            capi.check(self._L.apsu_b200_compute_powers_stage(self.db._h, s))
            if regions is not None:
                exchange(s + 1, regions)

    def stream(self) -> int:
        """cudaStream_t the context runs on (apsu_b200_ctx_get_stream), as an integer"""
        v = C.c_void_p()
        capi.check(self._L.apsu_b200_ctx_get_stream(self.db._h, C.byref(v)))
        return int(v.value or 0)

    def set_powers_partition(self, rank: int, size: int):
        """collective C2: this receiver computes chunk `rank` of `size` of every PowersDag level."""
        capi.check(self._L.apsu_b200_set_powers_partition(self.db._h, rank, size))
        self._part_rank, self._part_size = rank, size
        self._stage_regions = None

    def powers_exchange_regions(self, level: int):
        cap = self.db.params.bundle_idx_count()
        ptrs, sizes, n = (C.c_void_p * cap)(), (C.c_uint64 * cap)(), C.c_uint32()
        capi.check(self._L.apsu_b200_powers_exchange_regions(self.db._h, level, ptrs, sizes, cap, C.byref(n)))
        return [(int(ptrs[i]), int(sizes[i])) for i in range(n.value)]

    def ProcessBinBundleCaches(self):
        capi.check(self._L.apsu_b200_eval_all(self.db._h))

    def set_eval_chunk(self, bin_bundles: int):
        """BinBundles per Paterson-Stockmeyer chunk = per delivery of ProcessBinBundleCachesStreamed (rebuilds the plan:
        call it before loading a query)."""
        capi.check(self._L.apsu_b200_ctx_set_eval_chunk(self.db._h, bin_bundles))

    def ProcessBinBundleCachesStreamed(self, on_result=None):
        """ProcessBinBundleCache for every BinBundle with per-BinBundle delivery (send_rp_fun, receiver_ddh.cpp:527-534):
        on_result(ResultPackage) is called as each BinBundle's ciphertext reaches the host, while later chunks are still
        evaluated.  Returns the ResultPackages in delivery order."""
        n = self.db.get_bin_bundle_count()
        N = self.db.params.poly_modulus_degree()
        out = np.zeros((max(n, 1), 2, N), dtype=np.uint64)
        delivered = []
        FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64))

        def cb(_user, b, c, ct):
            arr = np.ctypeslib.as_array(ct, shape=(2, 1, N)).copy()
            rp = ResultPackage(int(b), int(c), arr)
            delivered.append(rp)
            if on_result is not None:
                on_result(rp)

        fn = FN(cb)
        capi.check(self._L.apsu_b200_eval_all_stream(self.db._h, capi.ptr(out), C.cast(fn, C.c_void_p), None))
        return delivered

    def get_power(self, bundle_idx: int, power: int):
        L, ntt = C.c_uint32(), C.c_int()
        capi.check(self._L.apsu_b200_get_power(self.db._h, bundle_idx, power, None, C.byref(L), C.byref(ntt)))
        out = np.zeros((2, L.value, self.db.params.poly_modulus_degree()), dtype=np.uint64)
        capi.check(self._L.apsu_b200_get_power(self.db._h, bundle_idx, power, capi.ptr(out), C.byref(L), C.byref(ntt)))
        return L.value, bool(ntt.value), out

    def results(self):
        n = self.db.get_bin_bundle_count()
        N = self.db.params.poly_modulus_degree()
        out = np.zeros((n, 2, N), dtype=np.uint64)
        b = np.zeros(n, dtype=np.uint32)
        c = np.zeros(n, dtype=np.uint32)
        capi.check(self._L.apsu_b200_fetch_results(self.db._h, capi.ptr(out), capi.ptr(b), capi.ptr(c)))
        return [ResultPackage(int(b[k]), int(c[k]), out[k].reshape(2, 1, N)) for k in range(n)]

    # ---- one-call API: host buffers in, host buffers out ----
    def RunQuery(self, query: Query, masks: np.ndarray):
        """HE part of Receiver::RunQuery (receiver_ddh.cpp:295-369): returns one ResultPackage per BinBundle."""
        n = self.db.get_bin_bundle_count()
        N = self.db.params.poly_modulus_degree()
        masks = np.ascontiguousarray(masks, dtype=np.uint64)
        out = np.zeros((n, 2, N), dtype=np.uint64)
        b = np.zeros(n, dtype=np.uint32)
        c = np.zeros(n, dtype=np.uint32)
        capi.check(self._L.apsu_b200_run_query(
            self.db._h, query.src_powers, len(query.src_powers), capi.ptr(query.cts), capi.ptr(query.relin_keys),
            capi.ptr(masks), masks.shape[0], capi.ptr(out), capi.ptr(b), capi.ptr(c)))
        return [ResultPackage(int(b[k]), int(c[k]), out[k].reshape(2, 1, N)) for k in range(n)]

    def RunQuerySeeded(self, src_powers, c0, seeds, relin_c0, relin_seeds, mask_seed: bytes | None, cache_counts=None):
        """The HE part of RunQuery fed as the reference's is (apsu_b200_run_query_seeded): seeded ciphertexts / keys as on
        the wire, masks drawn inside the call.  -> (ResultPackages, random_matrix [npack][items_per_bundle][2])"""
        p = self.db.params
        bic, N = p.bundle_idx_count(), p.poly_modulus_degree()
        if cache_counts is None:
            cache_counts = [self.db.get_bin_bundle_count(b) for b in range(bic)]
        alpha = max(max(cache_counts), 1)
        padded = np.ascontiguousarray([1 if c >= cache_counts[b] else 0 for c in range(alpha) for b in range(bic)], dtype=np.uint8)
        n = self.db.get_bin_bundle_count()
        sp = np.ascontiguousarray(list(src_powers), dtype=np.uint32)
        c0 = np.ascontiguousarray(c0, dtype=np.uint64)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint8)
        rc0 = None if relin_c0 is None else np.ascontiguousarray(relin_c0, dtype=np.uint64)
        rs = None if relin_seeds is None else np.ascontiguousarray(relin_seeds, dtype=np.uint8)
        ms = None if mask_seed is None else np.frombuffer(bytes(mask_seed), dtype=np.uint8).copy()
        blocks = np.zeros((alpha * bic, p.items_per_bundle(), 2), dtype=np.uint64)
        out = np.zeros((max(n, 1), 2, N), dtype=np.uint64)
        b = np.zeros(max(n, 1), dtype=np.uint32)
        c = np.zeros(max(n, 1), dtype=np.uint32)
        capi.check(self._L.apsu_b200_run_query_seeded(
            self.db._h, sp, len(sp), capi.ptr(c0), capi.ptr(seeds), capi.ptr(rc0), capi.ptr(rs), capi.ptr(ms), capi.ptr(padded), alpha * bic,
            capi.ptr(blocks), capi.ptr(out), capi.ptr(b), capi.ptr(c)))
        return [ResultPackage(int(b[k]), int(c[k]), out[k].reshape(2, 1, N)) for k in range(n)], blocks

    def timings(self) -> dict:
        t = capi.CTimings()
        capi.check(self._L.apsu_b200_last_timings(self.db._h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in capi.CTimings._fields_}

    def set_profiling(self, on: bool):
        capi.check(self._L.apsu_b200_set_profiling(self.db._h, int(on)))

    # ---- stand-alone evaluator ops ----
    def modulus_index(self, kind: int, i: int = 0) -> int:
        v = C.c_uint32()
        capi.check(self._L.apsu_b200_ctx_modulus_index(self.db._h, kind, i, C.byref(v)))
        return v.value

    def op_ntt(self, polys: np.ndarray, pattern, inverse=False) -> np.ndarray:
        N = self.db.params.poly_modulus_degree()
        out = np.ascontiguousarray(polys, dtype=np.uint64).copy()
        pat = np.ascontiguousarray(list(pattern), dtype=np.uint32)
        capi.check(self._L.apsu_b200_op_ntt(self.db._h, out.reshape(-1), out.size // N, pat, len(pat), int(inverse)))
        return out

    def op_prng_stream(self, seed: bytes, first_refill: int, n_words: int) -> np.ndarray:
        sd = np.frombuffer(bytes(seed), dtype=np.uint8).copy()
        out = np.zeros(n_words, dtype=np.uint64)
        capi.check(self._L.apsu_b200_op_prng_stream(self.db._h, capi.ptr(sd), first_refill, capi.ptr(out), n_words))
        return out

    def op_expand_seeds(self, num_primes: int, seeds: np.ndarray) -> np.ndarray:
        sd = np.ascontiguousarray(seeds, dtype=np.uint8).reshape(-1, 64)
        out = np.zeros((sd.shape[0], num_primes, self.db.params.poly_modulus_degree()), dtype=np.uint64)
        capi.check(self._L.apsu_b200_op_expand_seeds(self.db._h, num_primes, capi.ptr(sd), sd.shape[0], capi.ptr(out)))
        return out

    def op_multiply(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        n_ops, _, L, N = a.shape
        out = np.zeros((n_ops, 3, L, N), dtype=np.uint64)
        capi.check(self._L.apsu_b200_op_multiply(self.db._h, L, np.ascontiguousarray(a).reshape(-1), np.ascontiguousarray(b).reshape(-1), out.reshape(-1), n_ops))
        return out

    def op_relinearize(self, c3: np.ndarray) -> np.ndarray:
        n_ops, _, L, N = c3.shape
        out = np.zeros((n_ops, 2, L, N), dtype=np.uint64)
        capi.check(self._L.apsu_b200_op_relinearize(self.db._h, L, np.ascontiguousarray(c3).reshape(-1), out.reshape(-1), n_ops))
        return out

    def op_mod_switch_next(self, polys: np.ndarray) -> np.ndarray:
        n, L, N = polys.shape
        out = np.zeros((n, L - 1, N), dtype=np.uint64)
        capi.check(self._L.apsu_b200_op_mod_switch_next(self.db._h, L, np.ascontiguousarray(polys).reshape(-1), out.reshape(-1), n))
        return out
