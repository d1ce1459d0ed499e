#!/usr/bin/env python3
"""Generate tests/golden/oracle_vectors.json from the CPU oracle (run in the build container; committed).

The reference tree holds no ciphertext-level vectors, so these fixtures pin the *oracle's*
behaviour over time — derived constants for all 36 parameter sets, and SHA-256 digests of evaluator outputs
and of whole-query results on seeded scenarios.  GPU tests compare the CUDA path against the same digests.
"""
import hashlib
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import oracle as O  # noqa: E402
from harness import Scenario  # noqa: E402


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def main():
    table = json.loads((ROOT / "tests" / "golden" / "parameters.json").read_text())
    out = {"params": {}, "ops": {}, "queries": {}}
    for name, obj in sorted(table.items()):
        p = O.Params(obj, name)
        ctx_levels = dict(first_L=p.first_L)
        dag = O.powers_dag(p.ps_low_degree, p.max_items_per_bin, p.query_powers)
        ps = p.ps_low_degree
        first_chain = p.first_L - 1
        low_L = min(p.first_L, min(first_chain, 2 if ps else 1) + 1)
        high_L = min(p.first_L, 2)
        out["params"][name] = dict(
            N=p.N, t=p.t, primes=[hex(q) for q in p.primes], first_L=p.first_L, low_L=low_L, high_L=high_L,
            bundle_idx_count=p.bundle_idx_count, items_per_bundle=p.items_per_bundle, bins_per_bundle=p.bins_per_bundle,
            item_bit_count=p.item_bit_count, targets=len(dag), sources=len(p.query_powers),
            mults=sum(1 for n in dag if n["p1"] or n["p2"]), depth=max(n["depth"] for n in dag),
            roots=[O.minimal_primitive_root(2 * p.N, q) for q in p.primes], root_t=O.minimal_primitive_root(2 * p.N, p.t),
            aux=[hex(x) for x in O.get_primes(2 * p.N, 61, p.first_L + 3)],
        )
    # evaluator ops on seeded random operands (one PS config with 3 levels, one small)
    for name in ("16M-4096", "256K-512"):
        p = O.Params.load(name)
        ctx = O.Context.from_params(p)
        keys = O.Keys(ctx, 77)
        rng = np.random.default_rng(3)
        L = ctx.first_L

        def rnd(*shape):
            a = np.zeros(shape + (L, p.N), dtype=np.uint64)
            for j in range(L):
                a[..., j, :] = rng.integers(0, p.primes[j], size=shape + (p.N,), dtype=np.uint64)
            return a
        a, b = rnd(2), rnd(2)
        prod = ctx.multiply(a, b)
        rel = ctx.relinearize(prod, keys.relin)
        ms = ctx.mod_switch_next(rel)
        out["ops"][name] = dict(multiply=digest(prod), relinearize=digest(rel), mod_switch=digest(ms),
                                ntt0=digest(ctx.ntt(0, a[0, 0])), relin_keys=digest(keys.relin))
    # whole queries (same scenarios the GPU parity tests run)
    from test_gpu_query import CASES
    for name, degrees in CASES.items():
        sc = Scenario(name, degrees, planted=8)
        ses = sc.db.run_query(sc.src_powers, sc.cts, sc.relin, sc.masks, threads=8)
        res = {f"{b},{c}": digest(ct) for b, c, ct in ses.results()}
        for b, c, ct in ses.results():
            ok, budget, _, _ = sc.check_result(b, c, ct)
            assert ok and budget > 0, (name, b, c, budget)
        out["queries"][name] = dict(degrees=degrees, cts=digest(sc.cts), masks=digest(sc.masks), results=res)
        print(name, "ok", len(res), "bundles")
    path = ROOT / "tests" / "golden" / "oracle_vectors.json"
    path.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
