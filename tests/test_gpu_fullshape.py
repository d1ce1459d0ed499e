"""Full-shape parity (slow, GPU): the BASELINE configurations at their REAL degrees, synthetic DB words (the timing DB of
bench.py, same seeds), every result ciphertext compared bit for bit with the CPU oracle.
  * 16M-4096, receiver set 2^24: all 28 BinBundles (1304 coefficients each for the full ones, 45 inner polynomials);
  * 256M-4096: one FULL BinBundle (4000 coefficients, 311-product PowersDag) per bundle index plus a ragged one.
These are the shapes where the DB-stream kernel runs many groups per tile block and the PowersDag is widest; the
smaller tests stop at degree 700."""
import json

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


def _run(name, degrees, threads=None):
    import os
    import apsu_b200
    import bench
    pj = bench.load_params_json(name)
    p = O.Params(pj, name + ".json")
    threads = threads or os.cpu_count() or 1
    bic = p.bundle_idx_count
    npack = max(len(r) for r in degrees) * bic
    cts, relin, masks = bench.synth_query(p.primes, p.t, p.N, p.first_L, p.K, len(p.query_powers), bic, npack, bench.SEEDS["query"])
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(json.dumps(pj)), 0)
    try:
        for b in range(bic):
            for c, d in enumerate(degrees[b]):
                assert db.add_bin_bundle_synthetic(b, d + 1, bench.SEEDS["db"] * 1000 + b * 64 + c) == c
        rx = apsu_b200.Receiver(db)
        got = {(r.bundle_idx, r.cache_idx): r.psu_result.reshape(2, -1)
               for r in rx.RunQuery(apsu_b200.Query(p.query_powers, cts, relin), masks)}
    finally:
        db.close()
    assert len(got) == sum(len(r) for r in degrees)
    # the oracle holds one bundle index at a time (a full 16M-4096 BinBundle is 241 MB of plaintexts)
    checked = 0
    for b in range(bic):
        pairs = [(b, c) for c in range(len(degrees[b]))]
        if not pairs:
            continue
        ref = bench.oracle_results(pj, name, degrees, cts, relin, masks, pairs, threads)
        for key in pairs:
            assert np.array_equal(got[key], ref[key]), (name, key)
            checked += 1
    return checked


def test_16m_4096_every_binbundle_at_full_degree():
    import bench
    degrees = bench.simulate_bundle_degrees(bench.load_params_json("16M-4096"), 24, bench.SEEDS["db"])
    assert sum(len(r) for r in degrees) == 28 and max(max(r) for r in degrees) == 1303
    assert _run("16M-4096", degrees) == 28


def test_256m_4096_full_degree_binbundle_per_bundle_index():
    # 3999 = max_items_per_bin - 1: 12 inner polynomials of 310 terms + the 3999 % 311 remainder
    degrees = [[3999], [3999, 1234], [3999]]
    assert _run("256M-4096", degrees) == 4
