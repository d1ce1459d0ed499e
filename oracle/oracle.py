"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes binding of oracle/liborc.so).

CPU restatement of APSU's receiver-side query evaluation and of the SEAL 3.7 algorithms under it
(see oracle/seal_restate.hpp for provenance; PARITY UNPINNED — no ciphertext-level golden vectors
exist in the reference tree).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs import this module; the product package apsu_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import json
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def _host_tag() -> str:
    """CPU model + ISA flags of this host: liborc.so is compiled with -march=native, so a binary built elsewhere (the
    build container) is rebuilt on the box that runs it (GPU box host cores for the CPU baseline)."""
    import hashlib
    model, flags = "", ""
    try:
        for line in pathlib.Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name") and not model:
                model = line.split(":", 1)[1].strip()
            if line.startswith("flags") and not flags:
                flags = line.split(":", 1)[1].strip()
    except OSError:
        pass
    return model + " " + hashlib.sha256(flags.encode()).hexdigest()[:16]


def build(force: bool = False) -> pathlib.Path:
    so, tag_file = _HERE / "liborc.so", _HERE / "liborc.so.host"
    srcs = [_HERE / n for n in ("oracle_capi.cpp", "apsu_restate.hpp", "seal_restate.hpp", "prng_restate.hpp")]
    tag = _host_tag()
    stale = not so.exists() or any(s.exists() and s.stat().st_mtime > so.stat().st_mtime for s in srcs)
    other_host = not tag_file.exists() or tag_file.read_text() != tag
    if force or stale or other_host:
        subprocess.check_call(["make", "-C", str(_HERE), "-s", "-B"])
        tag_file.write_text(tag)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        L.orc_last_error.restype = C.c_char_p
        L.orc_plain_modulus_batching.restype = C.c_uint64
        L.orc_plain_modulus_batching.argtypes = [C.c_size_t, C.c_int]
        L.orc_minimal_primitive_root.restype = C.c_uint64
        L.orc_minimal_primitive_root.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_get_primes.argtypes = [C.c_uint64, C.c_int, C.c_size_t, u64p]
        L.orc_coeff_modulus_create.argtypes = [C.c_size_t, C.POINTER(C.c_int), C.c_size_t, u64p]
        L.orc_ntt_mod.argtypes = [C.c_size_t, C.c_uint64, u64p, C.c_int]
        L.orc_ctx_create.restype = C.c_void_p
        L.orc_ctx_create.argtypes = [C.c_size_t, C.c_uint64, u64p, C.c_size_t]
        L.orc_ctx_destroy.argtypes = [C.c_void_p]
        for f in ("orc_ctx_first_L",):
            getattr(L, f).restype = C.c_size_t
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_ctx_level_for_chain_idx.restype = C.c_size_t
        L.orc_ctx_level_for_chain_idx.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_ctx_root.restype = C.c_uint64
        L.orc_ctx_root.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_ctx_aux_base.restype = C.c_size_t
        L.orc_ctx_aux_base.argtypes = [C.c_void_p, C.c_size_t, u64p]
        L.orc_ntt.argtypes = [C.c_void_p, C.c_size_t, u64p, C.c_int]
        L.orc_encode.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p]
        L.orc_decode.argtypes = [C.c_void_p, u64p, u64p]
        L.orc_plain_to_ntt.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p]
        L.orc_multiply.argtypes = [C.c_void_p, C.c_size_t, u64p, C.c_size_t, u64p, C.c_size_t, u64p]
        L.orc_relinearize.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p, u64p]
        L.orc_mod_switch_next.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, u64p, u64p]
        L.orc_add_plain.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, u64p, u64p]
        L.orc_multiply_plain_normal.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, u64p, u64p, u64p]
        L.orc_keygen.restype = C.c_void_p
        L.orc_keygen.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_km_destroy.argtypes = [C.c_void_p]
        L.orc_km_relin_words.restype = C.c_size_t
        L.orc_km_relin_words.argtypes = [C.c_void_p]
        L.orc_km_relin.argtypes = [C.c_void_p, u64p]
        L.orc_km_secret.argtypes = [C.c_void_p, np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")]
        L.orc_encrypt.argtypes = [C.c_void_p, C.c_void_p, u64p, C.c_uint64, u64p]
        L.orc_decrypt_last.argtypes = [C.c_void_p, C.c_void_p, u64p, C.c_size_t, u64p]
        L.orc_powers_dag.argtypes = [C.c_uint32, C.c_uint32, u32p, C.c_size_t, u32p, u32p, u32p, u32p]
        L.orc_db_create.restype = C.c_void_p
        L.orc_db_create.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p, C.c_size_t]
        L.orc_db_destroy.argtypes = [C.c_void_p]
        L.orc_db_bundle_idx_count.restype = C.c_uint32
        L.orc_db_bundle_idx_count.argtypes = [C.c_void_p]
        L.orc_db_bins_per_bundle.restype = C.c_uint32
        L.orc_db_bins_per_bundle.argtypes = [C.c_void_p]
        L.orc_db_bundle_count.restype = C.c_size_t
        L.orc_db_bundle_count.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_db_add_bundle_from_bins.argtypes = [C.c_void_p, C.c_uint32, u32p, u64p]
        L.orc_db_add_bundle_synthetic.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_db_bundle_ncoeffs.restype = C.c_size_t
        L.orc_db_bundle_ncoeffs.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_db_bundle_coeff.restype = C.c_size_t
        L.orc_db_bundle_coeff.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_run_query.restype = C.c_void_p
        L.orc_run_query.argtypes = [C.c_void_p, u32p, C.c_size_t, u64p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_session_destroy.argtypes = [C.c_void_p]
        L.orc_session_powers_ms.restype = C.c_double
        L.orc_session_powers_ms.argtypes = [C.c_void_p]
        L.orc_session_eval_ms.restype = C.c_double
        L.orc_session_eval_ms.argtypes = [C.c_void_p]
        L.orc_session_result_count.restype = C.c_size_t
        L.orc_session_result_count.argtypes = [C.c_void_p]
        L.orc_session_result.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), u64p]
        L.orc_session_power.restype = C.c_size_t
        L.orc_session_power.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_int), C.c_void_p]
        L.orc_session_eval_subset.restype = C.c_double
        L.orc_session_eval_subset.argtypes = [C.c_void_p, C.c_void_p, u32p, u32p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
        for f in ("orc_blake2b", "orc_blake2xb"):
            getattr(L, f).argtypes = [u8p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_prng_bytes.argtypes = [u8p, C.c_size_t, u8p]
        L.orc_sample_poly_uniform.argtypes = [u8p, u64p, C.c_size_t, C.c_size_t, u64p]
        L.orc_mask_values.argtypes = [u8p, u8p, C.c_size_t, C.c_size_t, C.c_uint64, u64p]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().orc_last_error().decode())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------
# number-theory helpers
# ---------------------------------------------------------------------------------------------
def get_primes(factor: int, bits: int, count: int):
    out = np.zeros(count, dtype=np.uint64)
    _check(lib().orc_get_primes(factor, bits, count, out))
    return [int(x) for x in out]


def coeff_modulus_create(N: int, bit_sizes):
    out = np.zeros(len(bit_sizes), dtype=np.uint64)
    arr = (C.c_int * len(bit_sizes))(*bit_sizes)
    _check(lib().orc_coeff_modulus_create(N, arr, len(bit_sizes), out))
    return [int(x) for x in out]


def plain_modulus_batching(N: int, bits: int) -> int:
    return int(lib().orc_plain_modulus_batching(N, bits))


def minimal_primitive_root(degree: int, modulus: int) -> int:
    return int(lib().orc_minimal_primitive_root(degree, modulus))


def ntt_mod(N: int, modulus: int, data: np.ndarray, inverse: bool = False) -> np.ndarray:
    out = np.ascontiguousarray(data, dtype=np.uint64).copy()
    _check(lib().orc_ntt_mod(N, modulus, out, int(inverse)))
    return out


# ---------------------------------------------------------------------------------------------
# SEAL's default random generator (prng_restate.hpp)
# ---------------------------------------------------------------------------------------------
def _bytes_arg(b):
    a = np.frombuffer(bytes(b), dtype=np.uint8) if len(b) else np.zeros(0, dtype=np.uint8)
    return a, (a.ctypes.data_as(C.c_void_p) if len(b) else None)


def blake2b(data: bytes, outlen: int = 64, key: bytes = b"") -> bytes:
    out = np.zeros(outlen, dtype=np.uint8)
    d, dp = _bytes_arg(data)
    k, kp = _bytes_arg(key)
    _check(lib().orc_blake2b(out, outlen, dp, len(data), kp, len(key)))
    return out.tobytes()


def blake2xb(data: bytes, outlen: int, key: bytes = b"") -> bytes:
    out = np.zeros(outlen, dtype=np.uint8)
    d, dp = _bytes_arg(data)
    k, kp = _bytes_arg(key)
    _check(lib().orc_blake2xb(out, outlen, dp, len(data), kp, len(key)))
    return out.tobytes()


def prng_bytes(seed: bytes, nbytes: int) -> bytes:
    """first nbytes of the Blake2xbPRNG stream keyed with the 64-byte seed"""
    out = np.zeros(nbytes, dtype=np.uint8)
    _check(lib().orc_prng_bytes(np.frombuffer(bytes(seed), dtype=np.uint8).copy(), nbytes, out))
    return out.tobytes()


def sample_poly_uniform(seed: bytes, moduli, N: int) -> np.ndarray:
    """seal::util::sample_poly_uniform with a fresh generator: [L][N]"""
    m = np.array([int(x) for x in moduli], dtype=np.uint64)
    out = np.zeros((len(m), N), dtype=np.uint64)
    _check(lib().orc_sample_poly_uniform(np.frombuffer(bytes(seed), dtype=np.uint8).copy(), m, len(m), N, out))
    return out


def mask_values(seed: bytes, padded, N: int, t: int) -> np.ndarray:
    """receiver_ddh.cpp:241-262: slot values of every pack index (zero rows for padded pairs)"""
    pad = np.ascontiguousarray(padded, dtype=np.uint8)
    out = np.zeros((len(pad), N), dtype=np.uint64)
    _check(lib().orc_mask_values(np.frombuffer(bytes(seed), dtype=np.uint8).copy(), pad, len(pad), N, t, out))
    return out


# ---------------------------------------------------------------------------------------------
# PSUParams (common/apsu/psu_params.cpp:95-180, 290-374) — oracle-side restatement in Python
# ---------------------------------------------------------------------------------------------
class Params:
    def __init__(self, obj: dict, name: str = ""):
        self.name = name
        tp, ip, qp, sp = obj["table_params"], obj["item_params"], obj["query_params"], obj["seal_params"]
        self.hash_func_count = int(tp["hash_func_count"])
        self.table_size = int(tp["table_size"])
        self.max_items_per_bin = int(tp["max_items_per_bin"])
        self.felts_per_item = int(ip["felts_per_item"])
        self.ps_low_degree = int(qp["ps_low_degree"])
        self.query_powers = sorted(set([1] + [int(x) for x in qp["query_powers"]]))
        self.N = int(sp["poly_modulus_degree"])
        if "plain_modulus" in sp and "plain_modulus_bits" in sp:
            raise ValueError("only one of plain_modulus and plain_modulus_bits must be specified")
        if "plain_modulus" in sp:
            self.t = int(sp["plain_modulus"])
        elif "plain_modulus_bits" in sp:
            self.t = plain_modulus_batching(self.N, int(sp["plain_modulus_bits"]))
        else:
            raise ValueError("neither plain_modulus nor plain_modulus_bits was specified")
        self.coeff_modulus_bits = [int(b) for b in sp["coeff_modulus_bits"]]
        self.primes = coeff_modulus_create(self.N, self.coeff_modulus_bits)
        self._validate()

    def _validate(self):
        if not self.table_size:
            raise ValueError("table_size cannot be zero")
        if not self.max_items_per_bin:
            raise ValueError("max_items_per_bin cannot be zero")
        if not 1 <= self.hash_func_count <= 8:
            raise ValueError("hash_func_count is too large or too small")
        if not 2 <= self.felts_per_item <= 32:
            raise ValueError("felts_per_item is too large or too small")
        if self.ps_low_degree > self.max_items_per_bin:
            raise ValueError("ps_low_degree cannot be larger than max_items_per_bin")
        if 0 in self.query_powers or 1 not in self.query_powers:
            raise ValueError("query_powers cannot contain 0 and must contain 1")
        for p in self.query_powers:
            if p > self.max_items_per_bin:
                raise ValueError("query_powers cannot contain values larger than max_items_per_bin")
            if p > self.ps_low_degree and p % (self.ps_low_degree + 1):
                raise ValueError("query_powers above ps_low_degree must be multiples of ps_low_degree + 1")
        if (self.t - 1) % (2 * self.N):
            raise ValueError("plain_modulus must be a prime congruent to 1 modulo 2*poly_modulus_degree")
        self.item_bit_count_per_felt = self.t.bit_length() - 1
        self.item_bit_count = self.item_bit_count_per_felt * self.felts_per_item
        if not 80 <= self.item_bit_count <= 128:
            raise ValueError("parameters result in too large or too small item_bit_count")
        self.items_per_bundle = self.N // self.felts_per_item
        if not self.items_per_bundle:
            raise ValueError("poly_modulus_degree is too small")
        self.bins_per_bundle = self.items_per_bundle * self.felts_per_item
        if self.table_size % self.items_per_bundle:
            raise ValueError("table_size must be a multiple of floor(poly_modulus_degree / felts_per_item)")
        self.bundle_idx_count = self.table_size // self.items_per_bundle
        self.K = len(self.primes)
        self.first_L = self.K - 1 if self.K > 1 else 1

    @staticmethod
    def load(path_or_name: str) -> "Params":
        p = pathlib.Path(path_or_name)
        if p.exists():
            return Params(json.loads(p.read_text()), p.name)
        table = json.loads((_HERE.parent / "tests" / "golden" / "parameters.json").read_text())
        key = path_or_name if path_or_name.endswith(".json") else path_or_name + ".json"
        return Params(table[key], key)

    def to_json(self, obj_cache={}) -> str:
        table = json.loads((_HERE.parent / "tests" / "golden" / "parameters.json").read_text())
        return json.dumps(table[self.name])


def powers_dag(ps_low: int, target_degree: int, sources):
    src = np.array(sorted(sources), dtype=np.uint32)
    n = target_degree + 1
    a = [np.zeros(n, dtype=np.uint32) for _ in range(4)]
    cnt = lib().orc_powers_dag(ps_low, target_degree, src, len(src), *a)
    if cnt < 0:
        raise RuntimeError(lib().orc_last_error().decode())
    return [dict(power=int(a[0][i]), depth=int(a[1][i]), p1=int(a[2][i]), p2=int(a[3][i])) for i in range(cnt)]


# ---------------------------------------------------------------------------------------------
# context / evaluator / harness
# ---------------------------------------------------------------------------------------------
class Context:
    def __init__(self, N: int, t: int, primes):
        self.N, self.t, self.primes = N, t, [int(p) for p in primes]
        self.K = len(self.primes)
        arr = np.array(self.primes, dtype=np.uint64)
        self.h = lib().orc_ctx_create(N, t, arr, len(arr))
        if not self.h:
            raise RuntimeError(lib().orc_last_error().decode())
        self.first_L = int(lib().orc_ctx_first_L(self.h))

    @staticmethod
    def from_params(p: Params) -> "Context":
        return Context(p.N, p.t, p.primes)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ctx_destroy(self.h)
            self.h = None

    def level_for_chain_idx(self, ci: int) -> int:
        return int(lib().orc_ctx_level_for_chain_idx(self.h, ci))

    def root(self, prime_idx: int) -> int:
        return int(lib().orc_ctx_root(self.h, prime_idx))

    def aux_base(self, L: int):
        out = np.zeros(16, dtype=np.uint64)
        nb = lib().orc_ctx_aux_base(self.h, L, out)
        return dict(m_sk=int(out[0]), gamma=int(out[1]), B=[int(x) for x in out[2:2 + nb]])

    def ntt(self, prime_idx: int, data: np.ndarray, inverse: bool = False) -> np.ndarray:
        out = np.ascontiguousarray(data, dtype=np.uint64).copy()
        _check(lib().orc_ntt(self.h, prime_idx, out, int(inverse)))
        return out

    def encode(self, values) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.uint64)
        out = np.zeros(self.N, dtype=np.uint64)
        _check(lib().orc_encode(self.h, v, len(v), out))
        return out

    def decode(self, plain: np.ndarray) -> np.ndarray:
        out = np.zeros(self.N, dtype=np.uint64)
        _check(lib().orc_decode(self.h, np.ascontiguousarray(plain, dtype=np.uint64), out))
        return out

    def plain_to_ntt(self, plain: np.ndarray, L: int) -> np.ndarray:
        out = np.zeros((L, self.N), dtype=np.uint64)
        _check(lib().orc_plain_to_ntt(self.h, L, np.ascontiguousarray(plain, dtype=np.uint64), out))
        return out

    # ciphertexts are ndarrays [size][L][N]
    def multiply(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        L = a.shape[1]
        out = np.zeros((a.shape[0] + b.shape[0] - 1, L, self.N), dtype=np.uint64)
        _check(lib().orc_multiply(self.h, L, np.ascontiguousarray(a), a.shape[0], np.ascontiguousarray(b), b.shape[0], out))
        return out

    def relinearize(self, c3: np.ndarray, keys: np.ndarray) -> np.ndarray:
        L = c3.shape[1]
        out = np.zeros((2, L, self.N), dtype=np.uint64)
        _check(lib().orc_relinearize(self.h, L, np.ascontiguousarray(c3), np.ascontiguousarray(keys), out))
        return out

    def mod_switch_next(self, c: np.ndarray) -> np.ndarray:
        size, L = c.shape[0], c.shape[1]
        out = np.zeros((size, L - 1, self.N), dtype=np.uint64)
        _check(lib().orc_mod_switch_next(self.h, L, size, np.ascontiguousarray(c), out))
        return out

    def add_plain(self, c: np.ndarray, plain: np.ndarray) -> np.ndarray:
        out = np.ascontiguousarray(c).copy()
        _check(lib().orc_add_plain(self.h, c.shape[1], c.shape[0], out, np.ascontiguousarray(plain, dtype=np.uint64)))
        return out

    def multiply_plain_normal(self, c: np.ndarray, plain: np.ndarray) -> np.ndarray:
        out = np.zeros_like(c)
        _check(lib().orc_multiply_plain_normal(self.h, c.shape[1], c.shape[0], np.ascontiguousarray(c),
                                               np.ascontiguousarray(plain, dtype=np.uint64), out))
        return out


class Keys:
    def __init__(self, ctx: Context, seed: int):
        self.ctx = ctx
        self.h = lib().orc_keygen(ctx.h, seed)
        if not self.h:
            raise RuntimeError(lib().orc_last_error().decode())
        n = int(lib().orc_km_relin_words(self.h))
        flat = np.zeros(max(n, 1), dtype=np.uint64)
        if n:
            lib().orc_km_relin(self.h, flat)
            self.relin = flat.reshape(ctx.K - 1, 2, ctx.K, ctx.N)
        else:
            self.relin = None
        self.secret = np.zeros(ctx.N, dtype=np.int8)
        lib().orc_km_secret(self.h, self.secret)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_km_destroy(self.h)
            self.h = None

    def encrypt(self, plain: np.ndarray, seed: int) -> np.ndarray:
        out = np.zeros((2, self.ctx.first_L, self.ctx.N), dtype=np.uint64)
        _check(lib().orc_encrypt(self.ctx.h, self.h, np.ascontiguousarray(plain, dtype=np.uint64), seed, out))
        return out

    def decrypt_last(self, ct: np.ndarray):
        """ct: [size][1][N] at the last level -> (plaintext coeffs, noise budget bits)"""
        out = np.zeros(self.ctx.N, dtype=np.uint64)
        budget = lib().orc_decrypt_last(self.ctx.h, self.h, np.ascontiguousarray(ct), ct.shape[0], out)
        if budget == -1000:
            raise RuntimeError(lib().orc_last_error().decode())
        return out, int(budget)


class ReceiverDB:
    """Oracle-side receiver DB: BinBundles as column-wise plaintexts (bin_bundle.cpp:366-430)."""

    def __init__(self, ctx: Context, p: Params):
        self.ctx, self.p = ctx, p
        qp = np.array(p.query_powers, dtype=np.uint32)
        self.h = lib().orc_db_create(ctx.h, p.felts_per_item, p.table_size, p.max_items_per_bin, p.ps_low_degree, qp, len(qp))
        if not self.h:
            raise RuntimeError(lib().orc_last_error().decode())
        self.bundle_idx_count = int(lib().orc_db_bundle_idx_count(self.h))
        self.bins_per_bundle = int(lib().orc_db_bins_per_bundle(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_db_destroy(self.h)
            self.h = None

    def add_bundle_from_bins(self, bundle_idx: int, bins) -> int:
        """bins: list (len bins_per_bundle) of lists of roots (felts)."""
        assert len(bins) == self.bins_per_bundle
        sizes = np.array([len(b) for b in bins], dtype=np.uint32)
        roots = np.array([x for b in bins for x in b] or [0], dtype=np.uint64)
        idx = lib().orc_db_add_bundle_from_bins(self.h, bundle_idx, sizes, roots)
        if idx < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return idx

    def add_bundle_synthetic(self, bundle_idx: int, ncoeffs: int, seed: int) -> int:
        idx = lib().orc_db_add_bundle_synthetic(self.h, bundle_idx, ncoeffs, seed)
        if idx < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return idx

    def bundle_count(self, bundle_idx: int) -> int:
        return int(lib().orc_db_bundle_count(self.h, bundle_idx))

    def bundle_coeffs(self, bundle_idx: int, cache_idx: int):
        """-> list of (L, ndarray) per degree; L==0 => coefficient form [N], else [L][N]"""
        n = int(lib().orc_db_bundle_ncoeffs(self.h, bundle_idx, cache_idx))
        out = []
        for d in range(n):
            L = int(lib().orc_db_bundle_coeff(self.h, bundle_idx, cache_idx, d, None))
            buf = np.zeros((max(L, 1), self.ctx.N), dtype=np.uint64)
            lib().orc_db_bundle_coeff(self.h, bundle_idx, cache_idx, d, _ptr(buf))
            out.append((L, buf if L else buf[0]))
        return out

    def run_query(self, src_powers, cts: np.ndarray, relin: np.ndarray | None, masks: np.ndarray | None,
                  threads: int = 1, powers_only: bool = False) -> "Session":
        """cts: [nsrc][bundle_idx_count][2][first_L][N]; masks: [alpha_max*bundle_idx_count][N]"""
        sp = np.array(list(src_powers), dtype=np.uint32)
        cts = np.ascontiguousarray(cts, dtype=np.uint64)
        relin_c = None if relin is None else np.ascontiguousarray(relin, dtype=np.uint64)
        masks_c = None if masks is None else np.ascontiguousarray(masks, dtype=np.uint64)
        h = lib().orc_run_query(self.h, sp, len(sp), cts, _ptr(relin_c), _ptr(masks_c), threads, int(powers_only))
        if not h:
            raise RuntimeError(lib().orc_last_error().decode())
        return Session(self, h, relin_c, masks_c)


class Session:
    def __init__(self, db: ReceiverDB, h, relin, masks):
        self.db, self.h, self._relin, self._masks = db, h, relin, masks
        self.powers_ms = float(lib().orc_session_powers_ms(h))
        self.eval_ms = float(lib().orc_session_eval_ms(h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_session_destroy(self.h)
            self.h = None

    def results(self):
        """-> list of (bundle_idx, cache_idx, ndarray [2][N])"""
        n = int(lib().orc_session_result_count(self.h))
        out = []
        for k in range(n):
            b, c = C.c_uint32(), C.c_uint32()
            buf = np.zeros((2, self.db.ctx.N), dtype=np.uint64)
            _check(lib().orc_session_result(self.h, k, C.byref(b), C.byref(c), buf))
            out.append((b.value, c.value, buf))
        return out

    def power(self, bundle_idx: int, power: int):
        """-> (L, is_ntt, ndarray [2][L][N]) or None"""
        ntt = C.c_int(0)
        L = int(lib().orc_session_power(self.h, bundle_idx, power, C.byref(ntt), None))
        if not L:
            return None
        buf = np.zeros((2, L, self.db.ctx.N), dtype=np.uint64)
        lib().orc_session_power(self.h, bundle_idx, power, C.byref(ntt), _ptr(buf))
        return L, bool(ntt.value), buf

    def eval_subset(self, pairs, relin, masks, threads: int = 1) -> float:
        """evaluate (bundle_idx, cache_idx) pairs against the stored powers; returns elapsed ms"""
        b = np.array([p[0] for p in pairs], dtype=np.uint32)
        c = np.array([p[1] for p in pairs], dtype=np.uint32)
        relin_c = None if relin is None else np.ascontiguousarray(relin, dtype=np.uint64)
        masks_c = np.ascontiguousarray(masks, dtype=np.uint64)
        ms = lib().orc_session_eval_subset(self.db.h, self.h, b, c, len(pairs), _ptr(relin_c), _ptr(masks_c), threads)
        if ms < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return float(ms)
