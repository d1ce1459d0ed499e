"""Shared scenario builder for the parity tests, smoke() and the CPU-baseline legs of bench.py.

Plays the *sender* (key generation, query encryption, result decryption — out of scope for the
product, SURVEY.md §2 "Sender") and builds receiver DBs with known contents, all through the CPU
oracle.  TEST INFRASTRUCTURE: imports oracle/, so the product package must never import this.
"""
from __future__ import annotations

import numpy as np

from oracle import oracle as O

# fixed seeds
DB_SEED, QUERY_SEED, KEY_SEED, MASK_SEED = 0xD8, 0x51, 0x4B, 0x4D


def bundle_shape(p: O.Params, ncoeffs: int):
    """(n_ntt, n_coeff_form) plaintext counts of a BinBundle with `ncoeffs` coefficients
    (bin_bundle.cpp:413: NTT unless i==0 / i % (ps_low+1)==0)."""
    if p.ps_low_degree:
        n_coeff = (ncoeffs + p.ps_low_degree) // (p.ps_low_degree + 1)
    else:
        n_coeff = 1
    return ncoeffs - n_coeff, n_coeff


class Scenario:
    """One receiver DB + one encrypted query with known plaintext semantics.

    degrees: list over bundle indices of lists of per-BinBundle polynomial degrees, i.e.
    degrees[b][c] = max bin load of BinBundle (b, c)  (ncoeffs = degree + 1).
    """

    def __init__(self, name: str, degrees=None, planted: int = 16, seed: int = 0, build_db: bool = True):
        self.p = p = O.Params.load(name)
        self.ctx = ctx = O.Context.from_params(p)
        self.keys = O.Keys(ctx, KEY_SEED + seed)
        rng = np.random.default_rng(QUERY_SEED + seed)
        N, t = p.N, p.t
        bic = p.bundle_idx_count
        if degrees is None:
            degrees = [[p.max_items_per_bin - 1] for _ in range(bic)]
        self.degrees = degrees
        self.alpha_max = max(len(d) for d in degrees)
        # query slot values per bundle index (felts of the cuckoo-hashed sender items)
        self.x = rng.integers(0, t, size=(bic, N), dtype=np.uint64)
        self.x[:, p.bins_per_bundle:] = 0
        # DB bins
        self.bins = {}  # (b, c) -> list of lists of roots
        self.planted = {}  # (b, c) -> slots whose query value is a root of that bin's polynomial
        dbrng = np.random.default_rng(DB_SEED + seed)
        self.db = O.ReceiverDB(ctx, p) if build_db else None
        for b in range(bic):
            for c, deg in enumerate(degrees[b]):
                assert deg <= p.max_items_per_bin - 1
                loads = dbrng.integers(0, deg + 1, size=p.bins_per_bundle)
                loads[dbrng.integers(0, p.bins_per_bundle)] = deg  # at least one bin reaches the degree
                slots = dbrng.choice(p.bins_per_bundle, size=min(planted, p.bins_per_bundle), replace=False)
                bins = []
                for i in range(p.bins_per_bundle):
                    r = dbrng.integers(0, t, size=int(loads[i]), dtype=np.uint64).tolist()
                    bins.append(r)
                hit = []
                for s in slots:
                    if bins[s]:
                        bins[s][int(dbrng.integers(0, len(bins[s])))] = int(self.x[b, s])
                        hit.append(int(s))
                self.bins[(b, c)] = bins
                self.planted[(b, c)] = sorted(hit)
                if build_db:
                    assert self.db.add_bundle_from_bins(b, bins) == c
        # masks: r = u32 % t per slot (receiver_ddh.cpp:256-262), dense [alpha_max][bic] table
        mrng = np.random.default_rng(MASK_SEED + seed)
        self.mask_values = (mrng.integers(0, 2**32, size=(self.alpha_max * bic, N), dtype=np.uint64) % np.uint64(t))
        self.masks = np.stack([ctx.encode(v) for v in self.mask_values])
        # encrypted source powers (plaintext_powers.cpp:33-99)
        self.src_powers = list(p.query_powers)
        cts = np.zeros((len(self.src_powers), bic, 2, ctx.first_L, N), dtype=np.uint64)
        for k, e in enumerate(self.src_powers):
            for b in range(bic):
                vals = np.array([pow(int(v), e, t) for v in self.x[b]], dtype=np.uint64)
                cts[k, b] = self.keys.encrypt(ctx.encode(vals), seed=(QUERY_SEED << 16) + k * 64 + b + seed)
        self.cts = cts
        self.relin = self.keys.relin

    def pack_idx(self, b: int, c: int) -> int:
        return b + c * self.p.bundle_idx_count

    def expected_slots(self, b: int, c: int) -> np.ndarray:
        """P_bin(x_slot) + r_slot mod t for every slot of BinBundle (b, c)."""
        p, t = self.p, self.p.t
        out = np.zeros(p.N, dtype=np.uint64)
        bins = self.bins[(b, c)]
        r = self.mask_values[self.pack_idx(b, c)]
        for s in range(p.N):
            v = 1
            if s < p.bins_per_bundle:
                xs = int(self.x[b, s])
                for root in bins[s]:
                    v = v * (xs - int(root)) % t
            else:
                v = 0
            # an empty column (no polynomial at all: slots beyond bins_per_bundle) contributes 0; an
            # empty bin holds the constant polynomial 1 (polyn_with_roots of no roots)
            out[s] = (v + int(r[s])) % t
        return out

    def check_result(self, b: int, c: int, ct: np.ndarray):
        """decrypt + decode a result ciphertext [2][N] and compare with the plaintext semantics."""
        plain, budget = self.keys.decrypt_last(ct.reshape(2, 1, self.p.N))
        got = self.ctx.decode(plain)
        exp = self.expected_slots(b, c)
        return bool(np.array_equal(got, exp)), budget, got, exp


# ---- row f3: mask generation restated (receiver/apsu/receiver_ddh.cpp:70-92, 241-283) ----
def splitmix64_at(seed: int, k: np.ndarray) -> np.ndarray:
    """word k of the counter-based splitmix64 stream (same stream as the synthetic DB fill)."""
    M = (1 << 64) - 1
    z = (np.uint64(seed) + (k.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def ref_vec_to_std_block(vals, felts_per_item: int, plain_modulus: int):
    """vec_to_std_block (receiver_ddh.cpp:70-92) on Python ints -> (low, high) 64-bit words of the block."""
    ln = 1
    while ((1 << ln) - 1) < plain_modulus:
        ln += 1
    mask = (1 << ln) - 1
    mask_lower = (1 << (ln >> 1)) - 1
    mask_higher = mask - mask_lower
    lower = higher = 0
    if felts_per_item & 1:
        lower = vals[felts_per_item - 1] & mask_lower
        higher = (vals[felts_per_item - 1] & mask_higher) >> ((ln >> 1) - 1)
    for pla in range(0, felts_per_item - 1, 2):
        lower = (vals[pla] & mask) | (lower << ln)
        higher = (vals[pla + 1] & mask) | (higher << ln)
    return lower & ((1 << 64) - 1), higher & ((1 << 64) - 1)


def ref_generate_masks(p, seed: bytes, cache_counts):
    """-> (values [npack][N], blocks [npack][items_per_bundle][2], padded [npack]) in pack order
    p = bundle_idx + cache_idx * bundle_idx_count (receiver_ddh.cpp:243-246, 346); the values are the oracle's
    restatement of SEAL's blake2xb generator keyed with the 64-byte seed (oracle/prng_restate.hpp), one 32-bit draw
    per slot of every non-padded pair in (cache_idx, bundle_idx) order (receiver_ddh.cpp:241-262)."""
    bic, N = p.bundle_idx_count, p.N
    alpha = max(max(cache_counts), 1)
    npack = alpha * bic
    padded = np.array([c >= cache_counts[b] for c in range(alpha) for b in range(bic)], dtype=bool)
    values = O.mask_values(seed, padded.astype(np.uint8), N, p.t)
    ipb = p.N // p.felts_per_item
    blocks = np.zeros((npack, ipb, 2), dtype=np.uint64)
    for k in range(npack):
        if padded[k]:
            blocks[k] = np.uint64((1 << 64) - 1)
            continue
        for i in range(ipb):
            lo, hi = ref_vec_to_std_block([int(v) for v in values[k, i * p.felts_per_item:(i + 1) * p.felts_per_item]], p.felts_per_item, p.t)
            blocks[k, i] = (lo, hi)
    return values, blocks, padded
