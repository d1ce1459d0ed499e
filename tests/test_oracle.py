"""CPU tests of the oracle (oracle/): derived-constant fixtures, an independent big-integer model at toy
ring sizes, algebraic properties, and the committed golden digests."""
import json
import pathlib

import numpy as np
import pytest

import bigint_model as M
from oracle import oracle as O

ROOT = pathlib.Path(__file__).resolve().parent.parent
GOLD = json.loads((ROOT / "tests" / "golden" / "oracle_vectors.json").read_text())
TABLE = json.loads((ROOT / "tests" / "golden" / "parameters.json").read_text())

# SURVEY.md Appendix B (derived there with an independent sympy model): name -> (primes, t, idx, mults, depth)
APPENDIX_B = {
    "16M-4096.json": ([0xfffffffff70001, 0xfffffffff78001, 0xfffffffffb4001, 0x3ffffffffc001], 4079617, 4, 66, 3),
    "256M-4096.json": ([0x3ffffffef4001, 0x3fffffffcc001, 0x3ffffffffc001, 0x3ffffac001, 0x3fff4001], 67043329, 3, 311, 3),
    "1M-1024-cmp.json": ([0xfffffdc001, 0xfffc6001, 0xfffde001], 40961, 3, 85, 1),
    "1M-4096-com.json": ([0xfffffffff78001, 0xfffffffffb4001, 0x3fff4001], 65537, 5, 13, 2),
    "256K-512.json": ([0xffffffffc001, 0xfffde001, 0xffc001], 65537, 1, 52, 1),
    "100K-1.json": ([0xffffffffc001], 65537, 1, 0, 0),
    "256M-1.json": ([0xfffffffff70001, 0xfffffffff78001, 0xfffffffffb4001, 0xfff88001], 4079617, 1, 325, 3),
}
# SURVEY.md Appendix D
ROOTS = {(65537, 4096): 13, (65537, 8192): 15, (40961, 4096): 12, (4079617, 8192): 805, (67043329, 8192): 4721,
         (0xffffffffc001, 4096): 36835623151, (0xfffde001, 4096): 753779, (0xffc001, 4096): 3689,
         (0xfffffffff70001, 8192): 3993828849016, (0x3ffffffffc001, 8192): 11286399139, (0x3ffffac001, 8192): 71485851}
AUX_8192 = [0x1ffffffffffa4001, 0x1ffffffffff74001, 0x1ffffffffff0c001, 0x1fffffffffec4001, 0x1fffffffffe10001,
            0x1fffffffffe00001, 0x1fffffffffdd0001]
AUX_4096 = [0x1ffffffffffde001, 0x1ffffffffffce001, 0x1ffffffffffa4001, 0x1ffffffffff92001, 0x1ffffffffff7a001]


def test_appendix_b_and_d_fixtures():
    for name, (primes, t, idx, mults, depth) in APPENDIX_B.items():
        p = O.Params(TABLE[name], name)
        assert p.primes == primes and p.t == t and p.bundle_idx_count == idx
        dag = O.powers_dag(p.ps_low_degree, p.max_items_per_bin, p.query_powers)
        assert sum(1 for n in dag if n["p1"]) == mults and max(n["depth"] for n in dag) == depth
    for (q, N), r in ROOTS.items():
        assert O.minimal_primitive_root(2 * N, q) == r
    assert O.get_primes(2 * 8192, 61, 7) == AUX_8192
    assert O.get_primes(2 * 4096, 61, 5) == AUX_4096


def test_all_parameter_sets_match_golden():
    assert len(TABLE) == 36
    for name, obj in TABLE.items():
        p = O.Params(obj, name)
        g = GOLD["params"][name]
        assert [hex(q) for q in p.primes] == g["primes"] and p.t == g["t"]
        assert (p.bundle_idx_count, p.items_per_bundle, p.bins_per_bundle, p.item_bit_count) == (
            g["bundle_idx_count"], g["items_per_bundle"], g["bins_per_bundle"], g["item_bit_count"])


def test_ntt_matches_definition_and_convolution():
    N = 64
    q = O.get_primes(2 * N, 40, 1)[0]
    psi = O.minimal_primitive_root(2 * N, q)
    rng = np.random.default_rng(0)
    a = rng.integers(0, q, size=N, dtype=np.uint64)
    b = rng.integers(0, q, size=N, dtype=np.uint64)
    fa = O.ntt_mod(N, q, a)
    assert [int(x) for x in fa] == M.ntt_by_definition([int(x) for x in a], q, psi)
    assert np.array_equal(O.ntt_mod(N, q, fa, inverse=True), a)
    fb = O.ntt_mod(N, q, b)
    prod = np.array([int(x) * int(y) % q for x, y in zip(fa, fb)], dtype=np.uint64)
    conv = O.ntt_mod(N, q, prod, inverse=True)
    assert [int(x) for x in conv] == M.negacyclic_mul([int(x) for x in a], [int(x) for x in b], q)


def _toy_context():
    N = 64
    primes = O.coeff_modulus_create(N, [40, 40, 36, 41])
    return N, 257, primes, O.Context(N, 257, primes)


def test_batch_encoder_roundtrip_and_slotwise_product():
    N, t, primes, ctx = _toy_context()
    rng = np.random.default_rng(1)
    u = rng.integers(0, t, size=N, dtype=np.uint64)
    v = rng.integers(0, t, size=N, dtype=np.uint64)
    pu, pv = ctx.encode(u), ctx.encode(v)
    assert np.array_equal(ctx.decode(pu), u)
    prod = M.negacyclic_mul([int(x) for x in pu], [int(x) for x in pv], t)
    assert np.array_equal(ctx.decode(np.array(prod, dtype=np.uint64)), (u * v) % np.uint64(t))


def _to_lists(ct):
    return [[[int(x) for x in ct[c, j]] for j in range(ct.shape[1])] for c in range(ct.shape[0])]


def test_bfv_ops_match_bigint_model_bit_exact():
    """multiply / relinearize / mod_switch / add_plain: C++ oracle == Python model, and decrypt correctly."""
    N, t, primes, ctx = _toy_context()
    keys = O.Keys(ctx, 5)
    rng = np.random.default_rng(2)
    L = ctx.first_L
    q = primes[:L]
    u = rng.integers(0, t, size=N, dtype=np.uint64)
    v = rng.integers(0, t, size=N, dtype=np.uint64)
    ca, cb = keys.encrypt(ctx.encode(u), 11), keys.encrypt(ctx.encode(v), 12)
    aux = ctx.aux_base(L)
    prod = ctx.multiply(ca, cb)
    model = M.behz_multiply(_to_lists(ca), _to_lists(cb), q, t, aux["m_sk"], aux["B"])
    assert _to_lists(prod) == model
    sec = [int(x) for x in keys.secret]
    plain, budget = M.crt_decrypt(model, sec, q, t)
    assert budget > 0 and np.array_equal(ctx.decode(np.array(plain, dtype=np.uint64)), (u * v) % np.uint64(t))
    # relinearize
    K = len(primes)
    klist = [[[[int(x) for x in keys.relin[J, c, I]] for I in range(K)] for c in range(2)] for J in range(K - 1)]
    fwd = lambda poly, I: [int(x) for x in ctx.ntt(I, np.array(poly, dtype=np.uint64))]
    inv = lambda poly, I: [int(x) for x in ctx.ntt(I, np.array(poly, dtype=np.uint64), inverse=True)]
    rel = ctx.relinearize(prod, keys.relin)
    assert _to_lists(rel) == M.relinearize(model, klist, primes, L, fwd, inv)
    plain, budget = M.crt_decrypt(_to_lists(rel), sec, q, t)
    assert budget > 0 and np.array_equal(ctx.decode(np.array(plain, dtype=np.uint64)), (u * v) % np.uint64(t))
    # the same one level down (what eval_patstock does)
    ms = ctx.mod_switch_next(rel)
    assert _to_lists(ms) == [M.mod_switch_next(p, q) for p in _to_lists(rel)]
    plain, budget = M.crt_decrypt(_to_lists(ms), sec, q[:-1], t)
    assert budget > 0 and np.array_equal(ctx.decode(np.array(plain, dtype=np.uint64)), (u * v) % np.uint64(t))
    prod2 = ctx.multiply(ms, ctx.mod_switch_next(ca))
    aux2 = ctx.aux_base(L - 1)
    assert _to_lists(prod2) == M.behz_multiply(_to_lists(ms), _to_lists(ctx.mod_switch_next(ca)), q[:-1], t, aux2["m_sk"], aux2["B"])
    # add_plain
    w = rng.integers(0, t, size=N, dtype=np.uint64)
    ap = ctx.add_plain(ms, ctx.encode(w))
    assert _to_lists(ap)[0] == M.add_plain(_to_lists(ms)[0], [int(x) for x in ctx.encode(w)], q[:-1], t)
    plain, _ = M.crt_decrypt(_to_lists(ap), sec, q[:-1], t)
    assert np.array_equal(ctx.decode(np.array(plain, dtype=np.uint64)), (u * v + w) % np.uint64(t))


def test_oracle_reproduces_golden_op_digests():
    import hashlib
    for name, g in GOLD["ops"].items():
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
        dg = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
        assert dg(prod) == g["multiply"] and dg(keys.relin) == g["relin_keys"]
        rel = ctx.relinearize(prod, keys.relin)
        assert dg(rel) == g["relinearize"] and dg(ctx.mod_switch_next(rel)) == g["mod_switch"]


@pytest.mark.parametrize("name", ["256K-512", "1M-4096-com", "100K-1"])
def test_query_semantics_and_golden_results(name):
    """decode(decrypt(result)) - mask == P_bin(x) slot-wise; planted members decrypt to exactly the mask;
    result digests equal the committed fixtures."""
    import hashlib
    from harness import Scenario
    g = GOLD["queries"][name]
    sc = Scenario(name, g["degrees"], planted=8)
    dg = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
    assert dg(sc.cts) == g["cts"] and dg(sc.masks) == g["masks"]
    ses = sc.db.run_query(sc.src_powers, sc.cts, sc.relin, sc.masks, threads=4)
    res = ses.results()
    assert len(res) == sum(len(d) for d in g["degrees"])
    for b, c, ct in res:
        assert dg(ct) == g["results"][f"{b},{c}"]
        ok, budget, got, exp = sc.check_result(b, c, ct)
        assert ok and budget > 0
        r = sc.mask_values[sc.pack_idx(b, c)]
        assert all(got[s] == r[s] for s in sc.planted[(b, c)])


# ---- SEAL's default generator restated (oracle/prng_restate.hpp): rows f2/f3 --------------------------------------
def test_blake2b_rfc7693_vector_and_hashlib():
    """BLAKE2b core of the oracle: the RFC 7693 appendix A vector, and hashlib for keyed / truncated digests."""
    import hashlib
    rfc = ("ba80a53f981c4d0d6a2797b69f12f6e94c212f14685ac4b74b12bb6fdbffa2d1"
           "7d87c5392aab792dc252d5de4533cc9518d38aa8dbf1925ab92386edd4009923")
    assert O.blake2b(b"abc").hex() == rfc
    rng = np.random.default_rng(7693)
    for n in (0, 1, 64, 127, 128, 129, 255, 256, 1000):
        d = rng.bytes(n)
        for kl in (0, 16, 64):
            k = rng.bytes(kl)
            for ol in (20, 32, 64):
                assert O.blake2b(d, ol, k) == hashlib.blake2b(d, digest_size=ol, key=k).digest(), (n, kl, ol)


def test_blake2xb_against_independent_model():
    """BLAKE2Xb: hashlib cannot express the expansion nodes (fanout = depth = 0), so the pin is a pure-Python model
    of the compression function with an explicit parameter block (tests/blake2_model.py), itself checked against
    hashlib on every parameter block hashlib accepts (tree parameters, node_offset carrying xof_length)."""
    import hashlib
    import blake2_model as M
    rng = np.random.default_rng(2)
    for n in (0, 3, 128, 129, 300):
        d = rng.bytes(n)
        for k in (b"", rng.bytes(64)):
            assert M.blake2b_param(d, M.param_block(64, len(k)), k) == hashlib.blake2b(d, key=k).digest()
            assert M.blake2b_param(d, M.param_block(48, len(k), 2, 3, 64, 7, 4096, 1, 64), k) == hashlib.blake2b(
                d, digest_size=48, key=k, fanout=2, depth=3, leaf_size=64, node_offset=7 | (4096 << 32), node_depth=1, inner_size=64).digest()
            for ol in (1, 64, 65, 200, 4096):
                assert O.blake2xb(d, ol, k) == M.blake2xb(d, ol, k), (n, ol)


def test_prng_stream_and_sample_poly_uniform():
    """Blake2xbPRNG = concatenated 4096-byte refills blake2xb(key = seed, input = refill counter); sample_poly_uniform =
    bulk fill + sequential redraws of rejected words, against a Python model of the same rule on the model's stream.
    The second modulus sits far below a power of two, so that rejections actually happen."""
    import blake2_model as M
    seed = bytes(range(64))
    s = O.prng_bytes(seed, 9000)
    assert s[:4096] == M.blake2xb((0).to_bytes(8, "little"), 4096, seed)
    assert s[4096:8192] == M.blake2xb((1).to_bytes(8, "little"), 4096, seed)
    assert s[8192:] == M.blake2xb((2).to_bytes(8, "little"), 4096, seed)[:808]
    N, moduli = 256, [0xFFFFFFFFFFC0001, 0xB000000000000001 >> 4]
    got = O.sample_poly_uniform(seed, moduli, N)
    stream = b"".join(M.blake2xb(c.to_bytes(8, "little"), 4096, seed) for c in range(4))
    words = np.frombuffer(stream, dtype="<u8")
    nxt, rejected = len(moduli) * N, 0
    for j, q in enumerate(moduli):
        mm = (2 ** 64 - 1) - ((2 ** 64 - 1) % q) - 1
        for i in range(N):
            r = int(words[j * N + i])
            while r >= mm:
                r = int(words[nxt])
                nxt += 1
                rejected += 1
            assert int(got[j, i]) == r % q, (j, i)
    assert rejected > 0


def test_seal_kat_expected_is_current():
    """tests/golden/seal_kat_expected.json (what tools/seal_kat/seal_kat.cpp must print when run against real SEAL 3.7)
    is what the oracle computes today — one configuration here, all four in tools/seal_kat/make_kat_expected.py."""
    import json
    import pathlib
    import sys
    root = pathlib.Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root / "tools" / "seal_kat"))
    import make_kat_expected as MK
    exp = json.loads((root / "tests" / "golden" / "seal_kat_expected.json").read_text())
    assert MK.config_vectors("256K-512") == exp["configs"]["256K-512"]
