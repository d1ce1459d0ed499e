#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): transforms in both directions, one small
query with Paterson-Stockmeyer and one without, a device DB build and device mask generation — every kernel of the
query path at sizes the sanitizer finishes in a minute or two.

  compute-sanitizer --tool racecheck python tools/sanitize_small.py
  APSU_B200_NTT_SPLIT=4 compute-sanitizer --tool memcheck python tools/sanitize_small.py

(On the shared B200 pool this round compute-sanitizer was closed by the operators, so the script itself ran — it checks its
results against the oracle — but no sanitizer report exists; run it on a box where the tool is allowed.)
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import apsu_b200  # noqa: E402
from harness import Scenario  # noqa: E402
from oracle import oracle as O  # noqa: E402

rng = np.random.default_rng(3)
for name in ("256K-512", "1M-4096-com"):
    p = O.Params.load(name)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(p.to_json()), 0)
    rx = apsu_b200.Receiver(db)
    K = len(p.primes)
    x = np.zeros((2, K, p.N), dtype=np.uint64)
    for j, q in enumerate(p.primes):
        x[:, j, :] = rng.integers(0, q, size=(2, p.N), dtype=np.uint64)
    pat = [rx.modulus_index(0, i) for i in range(K)]
    y = rx.op_ntt(x, pat, inverse=False)
    assert np.array_equal(rx.op_ntt(y, pat, inverse=True), x)
    db.close()
for name, degrees in (("256K-512", [[20, 5]]), ("1M-4096-com", [[30], [], [], [12], []])):
    sc = Scenario(name, degrees, planted=2)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
    for b in range(sc.p.bundle_idx_count):
        for c in range(len(sc.degrees[b])):
            db.add_bin_bundle(b, [a for (_, a) in sc.db.bundle_coeffs(b, c)])
    for _ in range(2):  # the second query replays the captured graphs
        got = {(r.bundle_idx, r.cache_idx): r.psu_result.reshape(2, -1)
               for r in apsu_b200.Receiver(db).RunQuery(apsu_b200.Query(sc.src_powers, sc.cts, sc.relin), sc.masks)}
    exp = {(b, c): ct for b, c, ct in sc.db.run_query(sc.src_powers, sc.cts, sc.relin, sc.masks, threads=4).results()}
    for key in exp:
        assert np.array_equal(got[key], exp[key]), key
    db.close()
print("sanitize_small ok")
