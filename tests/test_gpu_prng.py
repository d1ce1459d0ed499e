"""GPU parity of SEAL's default random generator on the device (apsu_b200/csrc/blake2.cuh) — the mask generator of
RunQuery (row f3) and the expansion of seeded query ciphertexts / relinearisation keys (row f2) — against the CPU
oracle's byte-oriented restatement (oracle/prng_restate.hpp), through the C ABI."""
import numpy as np
import pytest

from harness import Scenario
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _db(name):
    import apsu_b200
    p = O.Params.load(name)
    return p, apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(p.to_json()), 0)


def test_prng_stream_matches_oracle():
    import apsu_b200
    p, db = _db("256K-512")
    try:
        rx = apsu_b200.Receiver(db)
        seed = bytes((11 * i + 5) & 0xFF for i in range(64))
        ref = np.frombuffer(O.prng_bytes(seed, 5 * 4096), dtype="<u8")
        assert np.array_equal(rx.op_prng_stream(seed, 0, 5 * 512), ref)
        assert np.array_equal(rx.op_prng_stream(seed, 3, 700), ref[3 * 512:3 * 512 + 700])  # counter-addressable
    finally:
        db.close()


@pytest.mark.parametrize("name", ["16M-4096", "256M-4096", "1M-1024-cmp", "100K-1"])
def test_expand_seeds_matches_oracle(name):
    """sample_poly_uniform at the first data level (query ciphertexts) and at the key level (relinearisation keys)."""
    import apsu_b200
    p, db = _db(name)
    try:
        rx = apsu_b200.Receiver(db)
        rng = np.random.default_rng(5)
        seeds = rng.integers(0, 256, size=(3, 64), dtype=np.uint8)
        for L in sorted({p.first_L, min(p.K, 5)}):
            got = rx.op_expand_seeds(L, seeds)
            for k in range(3):
                assert np.array_equal(got[k], O.sample_poly_uniform(seeds[k].tobytes(), p.primes[:L], p.N)), (L, k)
    finally:
        db.close()


def test_expand_seeds_with_rejections():
    """A hand-made parameter set whose primes sit far below a power of two: about one word in 23 / one in 5 is
    rejected and redrawn from the words after the bulk, in order — the sequential part of sample_poly_uniform."""
    import apsu_b200
    N = 2048
    primes = []
    for start in (0xB50000000000001, 0xCC0000000000001):  # ~0.71 * 2^60, ~0.8 * 2^60; prime = 1 mod 2N
        q = start - (start % (2 * N)) + 1
        while not _is_prime(q):
            q += 2 * N
        primes.append(q)
    params = apsu_b200.PSUParams.from_fields(N, 65537, primes + [0xffffffffc001], 1, 409, 20, 5, 0, range(1, 21))
    db = apsu_b200.ReceiverDB(params, 0)
    try:
        rx = apsu_b200.Receiver(db)
        seeds = np.random.default_rng(6).integers(0, 256, size=(4, 64), dtype=np.uint8)
        got = rx.op_expand_seeds(2, seeds)
        for k in range(4):
            assert np.array_equal(got[k], O.sample_poly_uniform(seeds[k].tobytes(), primes, N)), k
    finally:
        db.close()


def _is_prime(n):
    if n % 2 == 0:
        return False
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


@pytest.mark.parametrize("name,degrees", [("1M-4096-com", [[30, 9], [20], [], [12], [18]]), ("16M-4096", [[50], [], [46, 3], []]), ("256K-512", [[20]])])
def test_seeded_query_equals_expanded_query(name, degrees):
    """Row f2: a query delivered as (c0, seed) pairs + seeded relinearisation keys gives bit-identical result
    ciphertexts to the same query delivered expanded (the oracle evaluates the expanded form; the expansion itself is
    the oracle's sample_poly_uniform)."""
    import apsu_b200
    sc = Scenario(name, degrees, planted=4)
    p = sc.p
    rng = np.random.default_rng(8)
    nsrc, bic = sc.cts.shape[0], sc.cts.shape[1]
    seeds = rng.integers(0, 256, size=(nsrc, bic, 64), dtype=np.uint8)
    cts = sc.cts.copy()
    for k in range(nsrc):
        for b in range(bic):
            cts[k, b, 1] = O.sample_poly_uniform(seeds[k, b].tobytes(), p.primes[:p.first_L], p.N)
    relin, rseeds = None, None
    if sc.relin is not None:
        relin = sc.relin.copy()
        rseeds = rng.integers(0, 256, size=(p.K - 1, 64), dtype=np.uint8)
        for J in range(p.K - 1):
            relin[J, 1] = O.sample_poly_uniform(rseeds[J].tobytes(), p.primes, p.N)
    exp = {(b, c): ct for b, c, ct in sc.db.run_query(sc.src_powers, cts, relin, sc.masks, threads=4).results()}
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(p.to_json()), 0)
    try:
        for b in range(bic):
            for c in range(len(degrees[b])):
                db.add_bin_bundle(b, [a for (_, a) in sc.db.bundle_coeffs(b, c)])
        rx = apsu_b200.Receiver(db)
        rx.load_query_seeded(sc.src_powers, cts[:, :, 0], seeds, None if relin is None else relin[:, 0], rseeds)
        rx.set_masks(sc.masks)
        rx.ComputePowers()
        rx.ProcessBinBundleCaches()
        got = {(r.bundle_idx, r.cache_idx): r.psu_result.reshape(2, -1) for r in rx.results()}
        assert set(got) == set(exp)
        for key in exp:
            assert np.array_equal(got[key], exp[key]), key
    finally:
        db.close()


def test_query_with_out_of_range_residue_is_rejected():
    """seal::is_valid_for (receiver/apsu/query.cpp:54-66): a residue >= its modulus makes the query invalid."""
    import apsu_b200
    sc = Scenario("256K-512", [[10]], planted=2)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
    try:
        db.add_bin_bundle(0, [a for (_, a) in sc.db.bundle_coeffs(0, 0)])
        rx = apsu_b200.Receiver(db)
        bad = sc.cts.copy()
        bad[1, 0, 1, 1, 77] = sc.p.primes[1]
        with pytest.raises(ValueError):
            rx.load_query(apsu_b200.Query(sc.src_powers, bad, sc.relin))
        rx.load_query(apsu_b200.Query(sc.src_powers, sc.cts, sc.relin))  # and the context stays usable
        badk = sc.relin.copy()
        badk[0, 0, 2, 5] = sc.p.primes[2] + 3
        with pytest.raises(ValueError):
            rx.load_query(apsu_b200.Query(sc.src_powers, sc.cts, badk))
    finally:
        db.close()


def test_run_query_seeded_whole_call():
    """apsu_b200_run_query_seeded: seeded query + seeded keys + masks drawn on the device from a 64-byte seed, one call —
    equals the oracle fed with the expanded ciphertexts and the restated masks; random_matrix equals the restatement."""
    import apsu_b200
    from harness import ref_generate_masks
    degrees = [[30, 9], [20], [], [12], [18]]
    sc = Scenario("1M-4096-com", degrees, planted=4)
    p = sc.p
    rng = np.random.default_rng(9)
    nsrc, bic = sc.cts.shape[0], sc.cts.shape[1]
    seeds = rng.integers(0, 256, size=(nsrc, bic, 64), dtype=np.uint8)
    cts = sc.cts.copy()
    for k in range(nsrc):
        for b in range(bic):
            cts[k, b, 1] = O.sample_poly_uniform(seeds[k, b].tobytes(), p.primes[:p.first_L], p.N)
    relin = sc.relin.copy()
    rseeds = rng.integers(0, 256, size=(p.K - 1, 64), dtype=np.uint8)
    for J in range(p.K - 1):
        relin[J, 1] = O.sample_poly_uniform(rseeds[J].tobytes(), p.primes, p.N)
    mask_seed = bytes((5 * i + 2) & 0xFF for i in range(64))
    values, blocks, _ = ref_generate_masks(p, mask_seed, [len(d) for d in degrees])
    masks = np.stack([sc.ctx.encode(v) for v in values])
    exp = {(b, c): ct for b, c, ct in sc.db.run_query(sc.src_powers, cts, relin, masks, threads=4).results()}
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(p.to_json()), 0)
    try:
        for b in range(bic):
            for c in range(len(degrees[b])):
                db.add_bin_bundle(b, [a for (_, a) in sc.db.bundle_coeffs(b, c)])
        res, rm = apsu_b200.Receiver(db).RunQuerySeeded(sc.src_powers, cts[:, :, 0], seeds, relin[:, 0], rseeds, mask_seed)
        assert np.array_equal(rm, blocks)
        got = {(r.bundle_idx, r.cache_idx): r.psu_result.reshape(2, -1) for r in res}
        assert set(got) == set(exp)
        for key in exp:
            assert np.array_equal(got[key], exp[key]), key
    finally:
        db.close()
