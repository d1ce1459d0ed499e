#!/usr/bin/env python3
"""Writes tests/golden/seal_kat_expected.json: what tools/seal_kat/seal_kat.cpp must print when it is run against real
SEAL 3.7, computed here with the CPU oracle (the restatement under test).  See kat_spec.py."""
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))
from oracle import oracle as O  # noqa: E402
import kat_spec as S  # noqa: E402


def config_vectors(name: str) -> dict:
    p = O.Params.load(name)
    ctx = O.Context.from_params(p)
    N, L, K = p.N, ctx.first_L, p.K
    out = {"N": N, "plain_modulus": p.t, "coeff_modulus": [hex(q) for q in p.primes],
           "ntt_roots": [O.minimal_primitive_root(2 * N, q) for q in p.primes],
           "plain_ntt_root": O.minimal_primitive_root(2 * N, p.t)}
    aux = ctx.aux_base(L)
    out["rns_tool_first_level"] = {"m_sk": hex(aux["m_sk"]), "gamma": hex(aux["gamma"]), "base_B": [hex(x) for x in aux["B"]]}
    q = p.primes[:L]
    a = np.stack([S.rns_poly(1, q, N), S.rns_poly(2, q, N)])      # ciphertext a = (a0, a1), coefficient form, level L
    b = np.stack([S.rns_poly(3, q, N), S.rns_poly(4, q, N)])
    out["ntt_forward_q0"] = S.digest(ctx.ntt(0, a[0, 0]))
    out["ntt_inverse_q0"] = S.digest(ctx.ntt(0, a[0, 0], inverse=True))
    values = S.stream(900, N, p.t)
    plain = ctx.encode(values)
    out["batch_encode"] = S.digest(plain)
    out["plain_transform_to_ntt"] = S.digest(ctx.plain_to_ntt(plain, L))
    out["add_plain"] = S.digest(ctx.add_plain(a, plain))
    out["multiply_plain_coeff_form"] = S.digest(ctx.multiply_plain_normal(a, plain))
    prod = ctx.multiply(a, b)
    out["multiply"] = S.digest(prod)
    out["square"] = S.digest(ctx.multiply(a, a))
    if K > 1:
        keys = np.zeros((K - 1, 2, K, N), dtype=np.uint64)
        for J in range(K - 1):
            for c in range(2):
                keys[J, c] = S.rns_poly(100 + 2 * J + c, p.primes, N)
        rel = ctx.relinearize(prod, keys)
        out["relinearize"] = S.digest(rel)
        cur, lvl = rel, L
        while lvl > 1:
            cur = ctx.mod_switch_next(cur)
            lvl -= 1
            out[f"mod_switch_to_{lvl}_primes"] = S.digest(cur)
    # the random generator (rows f2/f3)
    out["blake2xb_prng_first_10000_bytes"] = __import__("hashlib").sha256(O.prng_bytes(S.PRNG_SEED, 10000)).hexdigest()
    out["sample_poly_uniform_first_level"] = S.digest(O.sample_poly_uniform(S.PRNG_SEED, q, N))
    out["sample_poly_uniform_key_level"] = S.digest(O.sample_poly_uniform(S.PRNG_SEED, p.primes, N))
    return out


def main():
    out = {"spec": "tools/seal_kat/kat_spec.py", "seed": hex(S.SEED), "configs": {n: config_vectors(n) for n in S.CONFIGS}}
    path = ROOT / "tests" / "golden" / "seal_kat_expected.json"
    path.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
