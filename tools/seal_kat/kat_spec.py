"""Language-neutral definition of the SEAL known-answer test (KAT) operands and of the digest file layout.

The reference tree has no ciphertext-level vectors and SEAL 3.7 (third-party, pinned in
cmake/APSUConfig.cmake.in:44) is not available in the build container, so the CPU oracle (oracle/) is "parity
unpinned".  This KAT closes that gap for whoever has SEAL: tools/seal_kat/seal_kat.cpp computes, with the REAL SEAL 3.7
API, exactly the quantities make_kat_expected.py computes with the oracle, on operands both sides derive from the same
counter-based generator below.  `python tools/seal_kat/compare.py seal_kat_out.json` then pins (or refutes) the oracle.

Operand word k of stream `sid` for modulus q:  mulhi64(splitmix64_at(SEED ^ (sid * 0x9E3779B97F4A7C15 mod 2^64), k), q)
(uniform in [0, q); the same construction as the synthetic DB fill of the product)."""
import hashlib

import numpy as np

SEED = 0x5EA15EA15EA15EA1
CONFIGS = ["256K-512", "1M-1024-cmp", "16M-4096", "256M-4096"]
M64 = (1 << 64) - 1


def splitmix64_at(seed: int, k: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (k.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def stream(sid: int, n: int, q: int) -> np.ndarray:
    """n words of stream `sid`, uniform in [0, q)"""
    seed = SEED ^ ((sid * 0x9E3779B97F4A7C15) & M64)
    w = splitmix64_at(seed, np.arange(n, dtype=np.uint64))
    return np.array([(int(x) * q) >> 64 for x in w], dtype=np.uint64)


def rns_poly(sid: int, primes, N: int) -> np.ndarray:
    """[len(primes)][N]: prime j uses stream sid*16 + j"""
    return np.stack([stream(sid * 16 + j, N, q) for j, q in enumerate(primes)])


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()


PRNG_SEED = bytes((37 * i + 11) & 0xFF for i in range(64))
