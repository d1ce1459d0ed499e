"""GPU parity of the stand-alone evaluator operations (K2-K6) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

CONFIGS = ["256K-512", "1M-4096-com", "16M-4096", "256M-4096", "100K-1"]


def _rand_polys(rng, primes, shape_prefix, N):
    out = np.zeros(tuple(shape_prefix) + (len(primes), N), dtype=np.uint64)
    for j, q in enumerate(primes):
        out[..., j, :] = rng.integers(0, q, size=tuple(shape_prefix) + (N,), dtype=np.uint64)
    return out


@pytest.fixture(scope="module", params=CONFIGS)
def env(request):
    import apsu_b200
    p = O.Params.load(request.param)
    ctx = O.Context.from_params(p)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(p.to_json()), 0)
    rx = apsu_b200.Receiver(db)
    yield p, ctx, db, rx
    db.close()


def test_ntt_roundtrip_and_parity(env):
    p, ctx, db, rx = env
    rng = np.random.default_rng(1)
    K = len(p.primes)
    polys = _rand_polys(rng, p.primes, (3,), p.N)  # [3][K][N]
    pat = [rx.modulus_index(0, i) for i in range(K)]
    fwd = rx.op_ntt(polys, pat, inverse=False)
    for c in range(3):
        for j in range(K):
            assert np.array_equal(fwd[c, j], ctx.ntt(j, polys[c, j])), (c, j)
    back = rx.op_ntt(fwd, pat, inverse=True)
    assert np.array_equal(back, polys)
    # plain modulus and BEHZ auxiliary primes
    aux = ctx.aux_base(ctx.first_L)
    for kind, i, q in [(3, 0, p.t), (1, 0, aux["m_sk"])] + [(2, i, b) for i, b in enumerate(aux["B"])]:
        x = rng.integers(0, q, size=(2, p.N), dtype=np.uint64)
        y = rx.op_ntt(x, [rx.modulus_index(kind, i)])
        for r in range(2):
            assert np.array_equal(y[r], O.ntt_mod(p.N, q, x[r])), (kind, i)
        assert np.array_equal(rx.op_ntt(y, [rx.modulus_index(kind, i)], inverse=True), x)


def test_mod_switch(env):
    p, ctx, db, rx = env
    rng = np.random.default_rng(2)
    for L in range(2, ctx.first_L + 1):
        x = _rand_polys(rng, p.primes[:L], (5,), p.N)
        got = rx.op_mod_switch_next(x)
        exp = ctx.mod_switch_next(x)  # treats [5][L][N] as a size-5 ciphertext: per-polynomial operation
        assert np.array_equal(got, exp), L


def test_multiply_and_relinearize(env):
    p, ctx, db, rx = env
    if ctx.K < 2:
        pytest.skip("single-prime parameters: no ciphertext multiplications on the path")
    rng = np.random.default_rng(3)
    keys = O.Keys(ctx, 77)
    rx.load_query  # noqa: B018 (API presence)
    import apsu_b200.capi as capi
    capi.check(capi.lib().apsu_b200_set_relin_keys(db._h, capi.ptr(np.ascontiguousarray(keys.relin))))
    for L in sorted({ctx.first_L, ctx.level_for_chain_idx(1)}):
        n_ops = 3
        a = _rand_polys(rng, p.primes[:L], (n_ops, 2), p.N)
        b = _rand_polys(rng, p.primes[:L], (n_ops, 2), p.N)
        b[2] = a[2]  # a square
        got = rx.op_multiply(a, b)
        for o in range(n_ops):
            exp = ctx.multiply(a[o], b[o])
            assert np.array_equal(got[o], exp), (L, o)
        rel = rx.op_relinearize(got)
        for o in range(n_ops):
            exp = ctx.relinearize(got[o], keys.relin)
            assert np.array_equal(rel[o], exp), (L, o)


_SHAPE_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import apsu_b200
from oracle import oracle as O
from harness import Scenario
rng = np.random.default_rng(5)
for name in ("16M-4096", "256K-512", "100K-1"):
    p = O.Params.load(name)
    ctx = O.Context.from_params(p)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(p.to_json()), 0)
    rx = apsu_b200.Receiver(db)
    K = len(p.primes)
    for count in (1, 5, 40):            # batches of count * K polynomials
        x = np.zeros((count, K, p.N), dtype=np.uint64)
        for j, q in enumerate(p.primes):
            x[:, j, :] = rng.integers(0, q, size=(count, p.N), dtype=np.uint64)
        pat = [rx.modulus_index(0, i) for i in range(K)]
        y = rx.op_ntt(x, pat, inverse=False)
        for c in (0, count - 1):
            for j in range(K):
                assert np.array_equal(y[c, j], ctx.ntt(j, x[c, j])), (name, count, c, j)
        assert np.array_equal(rx.op_ntt(y, pat, inverse=True), x), (name, count)
    db.close()
# one whole query (graphs replayed twice) under the same launch options
sc = Scenario("16M-4096", [[50], [], [46, 3], []], planted=4)
db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
for b in range(sc.p.bundle_idx_count):
    for c in range(len(sc.degrees[b])):
        db.add_bin_bundle(b, [a for (_, a) in sc.db.bundle_coeffs(b, c)])
exp = {{(b, c): ct for b, c, ct in sc.db.run_query(sc.src_powers, sc.cts, sc.relin, sc.masks, threads=4).results()}}
for _ in range(2):
    got = {{(r.bundle_idx, r.cache_idx): r.psu_result.reshape(2, -1)
           for r in apsu_b200.Receiver(db).RunQuery(apsu_b200.Query(sc.src_powers, sc.cts, sc.relin), sc.masks)}}
    assert set(got) == set(exp)
    for key in exp:
        assert np.array_equal(got[key], exp[key]), key
db.close()
print("ok")
"""


@pytest.mark.parametrize("env", [{"APSU_B200_NTT_SPLIT": "2"}, {"APSU_B200_NTT_SPLIT": "4"}, {"APSU_B200_NTT_SPLIT": "0"},
                                 {"APSU_B200_PDL": "1"}], ids=lambda e: ",".join(f"{k[10:]}={v}" for k, v in e.items()))
def test_alternative_launch_shapes_are_bit_exact(env):
    """The launch options that are process-wide switches — transforms cut into 2 / 4 cluster CTAs per polynomial
    (ntt.cuh: ntt_split_kernel; by default only small forward batches are split), never split, and programmatic
    dependent launch of the query kernels — produce the same words: transforms in both directions at three batch sizes
    and one whole query replayed twice, each in a fresh process."""
    import os
    import pathlib
    import subprocess
    import sys
    root = str(pathlib.Path(__file__).resolve().parent.parent)
    out = subprocess.run([sys.executable, "-c", _SHAPE_SCRIPT.format(root=root)], env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
