"""GPU parity of the whole query evaluation (ComputePowers + eval / eval_patstock) against the CPU
oracle, bit-exact, through the C ABI, plus the plaintext semantics decode(decrypt(result)) - r == P(x)."""
import numpy as np
import pytest

from harness import Scenario

pytestmark = pytest.mark.gpu

# name -> degrees[bundle_idx][cache_idx]; small degrees keep the oracle fast while still covering
# direct eval, PS with/without remainder, PS bundles below the PS threshold, empty bundle indices
CASES = {
    "256K-512": [[63, 20, 0]],
    "1M-1024-cmp": [[100, 7], [33], []],
    "1M-4096-com": [[97, 30, 12], [20], [25], [9, 8], [18]],
    "16M-4096": [[140, 45], [90], [], [44]],
    "256M-4096": [[700], [311, 5], [622]],
    "100K-1": [[19, 4]],
    "1M-1024-com": [[124, 13], [6, 5]],
}


def _upload(sc, db):
    for b in range(sc.p.bundle_idx_count):
        for c in range(len(sc.degrees[b])):
            coeffs = [arr for (_, arr) in sc.db.bundle_coeffs(b, c)]
            assert db.add_bin_bundle(b, coeffs) == c


@pytest.mark.parametrize("name", list(CASES))
def test_query_parity(name):
    import apsu_b200
    sc = Scenario(name, CASES[name], planted=8)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
    try:
        _upload(sc, db)
        rx = apsu_b200.Receiver(db)
        q = apsu_b200.Query(sc.src_powers, sc.cts, sc.relin)
        ses = sc.db.run_query(sc.src_powers, sc.cts, sc.relin, sc.masks, threads=8)
        # staged path: powers first
        rx.load_query(q)
        rx.set_masks(sc.masks)
        rx.ComputePowers()
        pses = sc.db.run_query(sc.src_powers, sc.cts, sc.relin, None, threads=8, powers_only=True)
        for b in range(sc.p.bundle_idx_count):
            if not sc.degrees[b]:
                continue
            for e in rx_targets(sc.p):
                exp = pses.power(b, e)
                L, ntt, got = rx.get_power(b, e)
                assert exp is not None and (L, ntt) == (exp[0], exp[1]), (b, e)
                assert np.array_equal(got, exp[2]), (b, e)
        rx.ProcessBinBundleCaches()
        got = {(r.bundle_idx, r.cache_idx): r.psu_result for r in rx.results()}
        exp = {(b, c): ct for b, c, ct in ses.results()}
        assert set(got) == set(exp)
        for key in exp:
            assert np.array_equal(got[key].reshape(2, -1), exp[key]), key
            ok, budget, _, _ = sc.check_result(key[0], key[1], got[key].reshape(2, -1))
            assert ok and budget > 0, (key, budget)
        # one-call path with host buffers gives the same bytes
        again = {(r.bundle_idx, r.cache_idx): r.psu_result for r in rx.RunQuery(q, sc.masks)}
        for key in exp:
            assert np.array_equal(again[key].reshape(2, -1), exp[key]), key
    finally:
        db.close()


def rx_targets(p):
    if p.ps_low_degree:
        h = p.ps_low_degree + 1
        return list(range(1, p.ps_low_degree + 1)) + list(range(h, p.max_items_per_bin + 1, h))
    return list(range(1, p.max_items_per_bin + 1))
