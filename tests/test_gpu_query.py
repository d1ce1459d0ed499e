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


def _all_parameter_files():
    import json
    import pathlib
    table = json.loads((pathlib.Path(__file__).resolve().parent / "golden" / "parameters.json").read_text())
    return sorted(k[:-5] for k in table)


def _sweep_degrees(p):
    """small BinBundles that still take every branch the parameter file allows: Paterson-Stockmeyer with and
    without a remainder polynomial plus a bundle below the PS threshold, or direct evaluation at full degree;
    first and last bundle index populated, the ones between empty (receiver_ddh.cpp:399-402)."""
    bic, maxdeg, ps = p.bundle_idx_count, p.max_items_per_bin - 1, p.ps_low_degree
    deg = [[] for _ in range(bic)]
    if ps:
        h = ps + 1
        d0 = min(maxdeg, (2 * h + ps // 2) if ps < 100 else h + ps // 4)
        d1 = min(maxdeg, 2 * h if ps < 100 else h)
        small = min(maxdeg, max(1, ps - 1) if ps < 100 else 20)
    else:
        d0, d1, small = maxdeg, max(1, maxdeg // 2), min(maxdeg, 3)
    deg[0] = [d0, small]
    if bic > 1:
        deg[bic - 1] = [d1]
    return deg


@pytest.mark.parametrize("name", [n for n in _all_parameter_files() if n not in CASES])
def test_query_parity_every_parameter_file(name):
    """north star: bit-exact result ciphertexts and identical decrypted output on EVERY parameters/*.json
    (the seven files in CASES get the deeper staged test above)."""
    import apsu_b200
    from oracle import oracle as O
    sc = Scenario(name, _sweep_degrees(O.Params.load(name)), planted=4)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
    try:
        _upload(sc, db)
        rx = apsu_b200.Receiver(db)
        exp = {(b, c): ct for b, c, ct in sc.db.run_query(sc.src_powers, sc.cts, sc.relin, sc.masks, threads=8).results()}
        got = {(r.bundle_idx, r.cache_idx): r.psu_result for r in rx.RunQuery(apsu_b200.Query(sc.src_powers, sc.cts, sc.relin), sc.masks)}
        assert set(got) == set(exp)
        for key in exp:
            assert np.array_equal(got[key].reshape(2, -1), exp[key]), key
            ok, budget, _, _ = sc.check_result(key[0], key[1], got[key].reshape(2, -1))
            assert ok and budget > 0, (key, budget)
    finally:
        db.close()


def rx_targets(p):
    if p.ps_low_degree:
        h = p.ps_low_degree + 1
        return list(range(1, p.ps_low_degree + 1)) + list(range(h, p.max_items_per_bin + 1, h))
    return list(range(1, p.max_items_per_bin + 1))


def test_results_match_committed_golden_digests():
    """the CUDA path against tests/golden/oracle_vectors.json (no oracle evaluation involved)."""
    import hashlib
    import json
    import pathlib
    import apsu_b200
    gold = json.loads((pathlib.Path(__file__).resolve().parent / "golden" / "oracle_vectors.json").read_text())
    for name in ("256K-512", "16M-4096", "1M-1024-com"):
        g = gold["queries"][name]
        sc = Scenario(name, g["degrees"], planted=8)
        db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
        try:
            _upload(sc, db)
            rx = apsu_b200.Receiver(db)
            res = rx.RunQuery(apsu_b200.Query(sc.src_powers, sc.cts, sc.relin), sc.masks)
            assert len(res) == len(g["results"])
            for r in res:
                d = hashlib.sha256(np.ascontiguousarray(r.psu_result).tobytes()).hexdigest()
                assert d == g["results"][f"{r.bundle_idx},{r.cache_idx}"], (name, r.bundle_idx, r.cache_idx)
        finally:
            db.close()


def test_db_roundtrip_synthetic_fill_and_mask_encode():
    """device-resident plaintexts read back equal what was uploaded; the synthetic fill equals the oracle's
    stream; BatchEncoder::encode on the device equals the oracle's."""
    import apsu_b200
    from oracle import oracle as O
    p = O.Params.load("1M-4096-com")
    ctx = O.Context.from_params(p)
    odb = O.ReceiverDB(ctx, p)
    odb.add_bundle_synthetic(0, 30, 1234)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(p.to_json()), 0)
    try:
        assert db.add_bin_bundle_synthetic(0, 30, 1234) == 0
        coeffs = odb.bundle_coeffs(0, 0)
        for k, (L, arr) in enumerate(coeffs):
            assert np.array_equal(db.bin_bundle_coeff(0, 0, k), arr), k
        assert db.add_bin_bundle(0, [a for _, a in coeffs]) == 1
        for k, (L, arr) in enumerate(coeffs):
            assert np.array_equal(db.bin_bundle_coeff(0, 1, k), arr), k
        assert db.get_bin_bundle_count(0) == 2 and db.get_bin_bundle_count() == 2
        rx = apsu_b200.Receiver(db)
        vals = np.random.default_rng(5).integers(0, p.t, size=(3, p.N), dtype=np.uint64)
        enc = rx.encode_masks(vals)
        for i in range(3):
            assert np.array_equal(enc[i], ctx.encode(vals[i]))
    finally:
        db.close()


@pytest.mark.parametrize("name,degrees", [("256K-512", [[63, 5]]), ("1M-4096-com", [[97, 9], [], [], [20], [0]]), ("100K-1", [[19]]),
                                          ("16M-4096", [[140], [], [], [44, 45]])])
def test_db_build_on_device(name, degrees):
    """row f1: BinBundle::regen_cache on the GPU (polyn_with_roots, encode, NTT) gives the plaintexts the oracle
    builds from the same bins, and a query against the device-built DB gives the oracle's result ciphertexts."""
    import apsu_b200
    sc = Scenario(name, degrees, planted=4)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
    try:
        for b in range(sc.p.bundle_idx_count):
            for c in range(len(sc.degrees[b])):
                assert db.add_bin_bundle_from_bins(b, sc.bins[(b, c)]) == c
                coeffs = sc.db.bundle_coeffs(b, c)
                for k, (L, arr) in enumerate(coeffs):
                    assert np.array_equal(db.bin_bundle_coeff(b, c, k), arr), (b, c, k)
        rx = apsu_b200.Receiver(db)
        exp = {(b, c): ct for b, c, ct in sc.db.run_query(sc.src_powers, sc.cts, sc.relin, sc.masks, threads=8).results()}
        got = {(r.bundle_idx, r.cache_idx): r.psu_result for r in rx.RunQuery(apsu_b200.Query(sc.src_powers, sc.cts, sc.relin), sc.masks)}
        assert set(got) == set(exp)
        for key in exp:
            assert np.array_equal(got[key].reshape(2, -1), exp[key]), key
        with pytest.raises(ValueError):  # a bin with max_items_per_bin items (receiver_db.cpp:388-389)
            bins = [[] for _ in range(sc.p.bins_per_bundle)]
            bins[3] = list(range(sc.p.max_items_per_bin))
            db.add_bin_bundle_from_bins(0, bins)
        with pytest.raises(ValueError):  # an item that is not a field element
            bins = [[] for _ in range(sc.p.bins_per_bundle)]
            bins[5] = [1, sc.p.t, 2]
            db.add_bin_bundle_from_bins(0, bins)
    finally:
        db.close()


@pytest.mark.parametrize("name,degrees", [("1M-4096-com", [[30, 9], [20], [], [12], [18]]), ("256K-512", [[20]]), ("16M-4096", [[50], [], [46, 3], []])])
def test_mask_generation_on_device(name, degrees):
    """row f3: masks drawn, encoded and packed into PEQT blocks on the device (receiver_ddh.cpp:241-283) equal the
    restatement; a query evaluated with the device-resident masks equals one given the same masks explicitly."""
    import apsu_b200
    from harness import ref_generate_masks
    sc = Scenario(name, degrees, planted=4)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
    try:
        _upload(sc, db)
        rx = apsu_b200.Receiver(db)
        counts = [len(d) for d in degrees]
        seed = bytes((7 * i + 3) & 0xFF for i in range(64))
        blocks, values = rx.generate_masks(seed, want_values=True)
        rvalues, rblocks, padded = ref_generate_masks(sc.p, seed, counts)
        assert np.array_equal(values, rvalues)
        assert np.array_equal(blocks, rblocks)
        # evaluate with the resident masks, then with the same masks passed in
        q = apsu_b200.Query(sc.src_powers, sc.cts, sc.relin)
        rx.load_query(q)
        rx.ComputePowers()
        rx.ProcessBinBundleCaches()
        got = {(r.bundle_idx, r.cache_idx): r.psu_result.copy() for r in rx.results()}
        masks = np.stack([sc.ctx.encode(v) for v in rvalues])
        again = {(r.bundle_idx, r.cache_idx): r.psu_result for r in rx.RunQuery(q, masks)}
        assert set(got) == set(again) and len(got) == sum(counts)
        for key in got:
            assert np.array_equal(got[key], again[key]), key
    finally:
        db.close()


@pytest.mark.parametrize("name,degrees,size", [("16M-4096", [[140, 45]], 2), ("1M-4096-com", [[97, 12]], 3), ("256K-512", [[63]], 2)])
def test_split_powers_dag(name, degrees, size):
    """collective C2 on one GPU: `size` receivers (one context each) split the PowersDag of a bundle index, the
    all-gather between DAG levels is emulated with device copies, and every receiver ends with the oracle's powers
    and result ciphertexts."""
    import torch
    import apsu_b200
    from apsu_b200.sharding import _DeviceRegion
    bic_deg = degrees + [[] for _ in range(100)]
    from oracle import oracle as O
    p0 = O.Params.load(name)
    sc = Scenario(name, bic_deg[:p0.bundle_idx_count], planted=4)
    dbs = [apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0) for _ in range(size)]
    try:
        rxs = []
        q = apsu_b200.Query(sc.src_powers, sc.cts, sc.relin)
        for r, db in enumerate(dbs):
            _upload(sc, db)
            rx = apsu_b200.Receiver(db)
            rx.set_powers_partition(r, size)
            rx.load_query(q)
            rx.set_masks(sc.masks)
            rxs.append(rx)
        lib = apsu_b200.capi.lib()
        import ctypes as C
        n = C.c_uint32()
        apsu_b200.capi.check(lib.apsu_b200_powers_stage_count(dbs[0]._h, C.byref(n)))
        for s in range(n.value):
            for db in dbs:
                apsu_b200.capi.check(lib.apsu_b200_compute_powers_stage(db._h, s))
            for db in dbs:
                apsu_b200.capi.check(lib.apsu_b200_ctx_synchronize(db._h))
            if s + 1 < n.value:
                regs = [rx.powers_exchange_regions(s + 1) for rx in rxs]
                for k in range(len(regs[0])):
                    words = regs[0][k][1] // 8
                    full = [torch.as_tensor(_DeviceRegion(regs[r][k][0], regs[r][k][1] * size), device="cuda") for r in range(size)]
                    for src in range(size):
                        for dst in range(size):
                            if src != dst:
                                full[dst][src * words:(src + 1) * words] = full[src][src * words:(src + 1) * words]
                torch.cuda.synchronize()
        pses = sc.db.run_query(sc.src_powers, sc.cts, sc.relin, None, threads=8, powers_only=True)
        exp = {(b, c): ct for b, c, ct in sc.db.run_query(sc.src_powers, sc.cts, sc.relin, sc.masks, threads=8).results()}
        for rx in rxs:
            for e in rx_targets(sc.p):
                want = pses.power(0, e)
                L, ntt, got = rx.get_power(0, e)
                assert (L, ntt) == (want[0], want[1]) and np.array_equal(got, want[2]), e
            rx.ProcessBinBundleCaches()
            got = {(r.bundle_idx, r.cache_idx): r.psu_result for r in rx.results()}
            assert set(got) == set(exp)
            for key in exp:
                assert np.array_equal(got[key].reshape(2, -1), exp[key]), key
        with pytest.raises(AssertionError):  # logic_error: the one-call ComputePowers needs an unsplit DAG
            apsu_b200.capi.check(lib.apsu_b200_compute_powers(dbs[0]._h))
    finally:
        for db in dbs:
            db.close()


@pytest.mark.parametrize("name,degrees", [("1M-4096-com", [[30, 9], [20], [], [12], [18]]), ("256K-512", [[40]]), ("16M-4096", [[60], [], [], [46]])])
def test_sender_decrypt_on_device(name, degrees):
    """row f4: ResultPackage::extract for a batch (decrypt at the last level, noise budget, BatchEncoder::decode) and
    the items' blocks on the device equal the oracle's decrypt + decode and the restated vec_to_std_block; the decoded
    slots are P_bin(x) + r, i.e. the whole receiver -> sender exchange closes on the GPU."""
    import apsu_b200
    from harness import ref_vec_to_std_block
    sc = Scenario(name, degrees, planted=4)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
    try:
        _upload(sc, db)
        rx = apsu_b200.Receiver(db)
        res = rx.RunQuery(apsu_b200.Query(sc.src_powers, sc.cts, sc.relin), sc.masks)
        cts = np.stack([r.psu_result.reshape(2, -1) for r in res])
        q0 = sc.p.primes[0]
        s_lift = np.array([int(v) % q0 for v in sc.keys.secret.astype(np.int64)], dtype=np.uint64)
        s_ntt = sc.ctx.ntt(0, s_lift)
        values, blocks, budget = rx.decrypt_results(s_ntt, cts)
        f, t = sc.p.felts_per_item, sc.p.t
        for k, r in enumerate(res):
            plain, b = sc.keys.decrypt_last(cts[k].reshape(2, 1, -1))
            assert np.array_equal(values[k], sc.ctx.decode(plain)), k
            assert int(budget[k]) == b and b > 0
            assert np.array_equal(values[k], sc.expected_slots(r.bundle_idx, r.cache_idx))
            for item in (0, 1, sc.p.N // f - 1):
                lo, hi = ref_vec_to_std_block([int(v) for v in values[k, item * f:(item + 1) * f]], f, t)
                assert (int(blocks[k, item, 0]), int(blocks[k, item, 1])) == (lo, hi), (k, item)
    finally:
        db.close()


def test_successive_queries_replay_the_captured_graphs():
    """the launch sequences are captured into CUDA graphs on the first query; later queries with other ciphertexts,
    keys and masks, and a query after the DB grew (new plan, new graphs), still match the oracle."""
    import apsu_b200
    sc = Scenario("1M-4096-com", [[30, 9], [20], [], [12], [18]], planted=4)
    other = Scenario("1M-4096-com", [[1], [], [], [], []], planted=1, seed=5, build_db=False)  # another sender: keys, query, masks
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
    try:
        _upload(sc, db)
        rx = apsu_b200.Receiver(db)

        def check(q_sc):
            masks = q_sc.masks
            if masks.shape[0] < sc.masks.shape[0]:
                masks = np.concatenate([masks] * (sc.masks.shape[0] // masks.shape[0] + 1))[:sc.masks.shape[0]]
            exp = {(b, c): ct for b, c, ct in sc.db.run_query(q_sc.src_powers, q_sc.cts, q_sc.relin, masks, threads=8).results()}
            got = {(r.bundle_idx, r.cache_idx): r.psu_result for r in rx.RunQuery(apsu_b200.Query(q_sc.src_powers, q_sc.cts, q_sc.relin), masks)}
            assert set(got) == set(exp)
            for key in exp:
                assert np.array_equal(got[key].reshape(2, -1), exp[key]), key
        check(sc)
        check(other)
        check(sc)
        # the DB grows: plan and graphs are rebuilt
        bins = sc.bins[(0, 1)]
        assert sc.db.add_bundle_from_bins(2, bins) == 0
        assert db.add_bin_bundle_from_bins(2, bins) == 0
        check(other)
    finally:
        db.close()


@pytest.mark.parametrize("name,big,few", [("16M-4096", 1303, 3), ("256M-4096", 2100, 2), ("1M-1-32", 227, 5)])
def test_db_build_large_degrees(name, big, few):
    """row f1 at the degrees the small scenarios do not reach: a full 16M-4096 bin (1303 items: the 44-register
    instance of k_polyn_with_roots_reg and every active-register case below it) and a 2100-item bin of 256M-4096 (more
    than 2048 coefficients: the shared-memory kernel).  Device-built plaintexts equal the oracle-built ones."""
    import apsu_b200
    from oracle import oracle as O
    p = O.Params.load(name)
    ctx = O.Context.from_params(p)
    rng = np.random.default_rng(11)
    bins = [[] for _ in range(p.bins_per_bundle)]
    for i in range(few):
        bins[i * 97 % p.bins_per_bundle] = rng.integers(0, p.t, size=big - i, dtype=np.uint64).tolist()
    for i in range(few, 40):
        bins[(i * 131 + 7) % p.bins_per_bundle] = rng.integers(0, p.t, size=int(rng.integers(0, 70)), dtype=np.uint64).tolist()
    odb = O.ReceiverDB(ctx, p)
    assert odb.add_bundle_from_bins(0, bins) == 0
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(p.to_json()), 0)
    try:
        assert db.add_bin_bundle_from_bins(0, bins) == 0
        coeffs = odb.bundle_coeffs(0, 0)
        assert len(coeffs) == big + 1
        for k, (L, arr) in enumerate(coeffs):
            assert np.array_equal(db.bin_bundle_coeff(0, 0, k), arr), k
    finally:
        db.close()


def test_error_behaviour_on_device():
    import apsu_b200
    sc = Scenario("256K-512", [[5]], planted=2)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
    try:
        rx = apsu_b200.Receiver(db)
        with pytest.raises(AssertionError):  # logic_error: powers before a query
            rx.ComputePowers()
        _upload(sc, db)
        with pytest.raises(ValueError):  # invalid_argument: wrong source powers (query.cpp:68-111)
            rx.load_query(apsu_b200.Query(sc.src_powers[:-1], sc.cts[:-1], sc.relin))
        with pytest.raises(ValueError):  # more coefficients than max_items_per_bin
            db.add_bin_bundle(0, [np.zeros(sc.p.N, dtype=np.uint64)] + [np.zeros((2, sc.p.N), dtype=np.uint64)] * sc.p.max_items_per_bin)
        rx.load_query(apsu_b200.Query(sc.src_powers, sc.cts, sc.relin))
        rx.ComputePowers()
        with pytest.raises(ValueError):  # no masks
            rx.ProcessBinBundleCaches()
    finally:
        db.close()


def test_cpp_facade_on_gpu(tmp_path):
    """Receiver::RunQuery of the C++ facade == the C-ABI path driven from Python, on the same synthetic inputs."""
    import json
    import pathlib
    import subprocess
    import apsu_b200
    root = pathlib.Path(__file__).resolve().parent.parent
    exe = tmp_path / "test_facade"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-o", str(exe), str(root / "tests" / "cpp" / "test_facade.cpp"),
                           f"-L{root / 'apsu_b200'}", "-lapsu_b200", f"-Wl,-rpath,{root / 'apsu_b200'}"])
    table = json.loads((root / "tests" / "golden" / "parameters.json").read_text())
    pj = tmp_path / "p.json"
    pj.write_text(json.dumps(table["1M-4096-com.json"]))
    out = subprocess.check_output([str(exe), "gpu", str(pj), "40", "99"], text=True)
    lines = [l for l in out.splitlines() if l.startswith("bundle_idx=")]
    assert len(lines) == 5
    # same inputs from Python: splitmix64 stream of the C++ test
    params = apsu_b200.PSUParams.Load(pj.read_text())
    N, t, primes = params.poly_modulus_degree(), params.plain_modulus(), params.coeff_modulus()
    K, L, bic = len(primes), len(primes) - 1, params.bundle_idx_count()
    state = [99]

    def sm():
        state[0] = (state[0] + 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = state[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        return z ^ (z >> 31)
    qp = params.query_powers()
    # the C++ test iterates an unordered_map for upload but draws values in std::set order of the powers
    cts = np.zeros((len(qp), bic, 2, L, N), dtype=np.uint64)
    for k in range(len(qp)):
        for b in range(bic):
            for c in range(2):
                for j in range(L):
                    cts[k, b, c, j] = [sm() % primes[j] for _ in range(N)]
    relin = np.array([sm() % primes[(i // N) % K] for i in range((K - 1) * 2 * K * N)], dtype=np.uint64).reshape(K - 1, 2, K, N)
    masks = np.array([sm() % t for _ in range(bic * N)], dtype=np.uint64).reshape(bic, N)
    db = apsu_b200.ReceiverDB(params, 0)
    try:
        for b in range(bic):
            db.add_bin_bundle_synthetic(b, 40, 99 + b)
        res = apsu_b200.Receiver(db).RunQuery(apsu_b200.Query(qp, cts, relin), masks)

        def fnv(a):
            h = 1469598103934665603
            for x in a.reshape(-1):
                h = ((h ^ int(x)) * 1099511628211) & (2**64 - 1)
            return h
        exp = {f"bundle_idx={r.bundle_idx} cache_idx={r.cache_idx} fnv={fnv(r.psu_result):016x}" for r in res}
        assert set(lines) == exp
    finally:
        db.close()


def test_results_are_delivered_per_binbundle():
    """apsu_b200_eval_all_stream: one delivery per BinBundle, chunk by chunk (direct BinBundles first, then the
    Paterson-Stockmeyer ones two at a time), every ciphertext identical to the batch call's."""
    import apsu_b200
    degrees = [[30, 9, 25, 7, 22], [20], [], [12, 15], [18]]  # 1M-4096-com: ps_low = 8, so degree <= 8 evaluates directly
    sc = Scenario("1M-4096-com", degrees, planted=4)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
    try:
        _upload(sc, db)
        rx = apsu_b200.Receiver(db)
        rx.load_query(apsu_b200.Query(sc.src_powers, sc.cts, sc.relin))
        rx.set_masks(sc.masks)
        rx.ComputePowers()
        rx.ProcessBinBundleCaches()
        want = {(r.bundle_idx, r.cache_idx): r.psu_result.copy() for r in rx.results()}
        seen = []
        rx.set_eval_chunk(2)
        rx.load_query(apsu_b200.Query(sc.src_powers, sc.cts, sc.relin))
        rx.ComputePowers()
        got = rx.ProcessBinBundleCachesStreamed(on_result=lambda rp: seen.append((rp.bundle_idx, rp.cache_idx)))
        assert len(got) == len(want) == 9 and len(set(seen)) == 9
        assert seen[0] == (0, 3)  # the only directly evaluated BinBundle (degree 7) is delivered first
        for rp in got:
            assert np.array_equal(rp.psu_result, want[(rp.bundle_idx, rp.cache_idx)]), (rp.bundle_idx, rp.cache_idx)
        exp = {(b, c): ct for b, c, ct in sc.db.run_query(sc.src_powers, sc.cts, sc.relin, sc.masks, threads=4).results()}
        for rp in got:
            assert np.array_equal(rp.psu_result.reshape(2, -1), exp[(rp.bundle_idx, rp.cache_idx)])
    finally:
        db.close()


def test_fused_transform_prologues_are_bit_exact(monkeypatch):
    """APSU_B200_FUSE=1: extension / tensor product / key-switch inner product computed inside the transforms that
    consume them (ntt.cuh: NttFuse) — an alternative schedule of the same arithmetic, kept for A/B measurements."""
    import apsu_b200
    monkeypatch.setenv("APSU_B200_FUSE", "1")
    for name, degrees in [("16M-4096", [[50], [], [46, 3], []]), ("256K-512", [[20, 63]])]:
        sc = Scenario(name, degrees, planted=4)
        db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), 0)
        try:
            _upload(sc, db)
            got = {(r.bundle_idx, r.cache_idx): r.psu_result.reshape(2, -1)
                   for r in apsu_b200.Receiver(db).RunQuery(apsu_b200.Query(sc.src_powers, sc.cts, sc.relin), sc.masks)}
            exp = {(b, c): ct for b, c, ct in sc.db.run_query(sc.src_powers, sc.cts, sc.relin, sc.masks, threads=4).results()}
            assert set(got) == set(exp)
            for key in exp:
                assert np.array_equal(got[key], exp[key]), (name, key)
        finally:
            db.close()
