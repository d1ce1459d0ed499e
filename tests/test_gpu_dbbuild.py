"""Row f1 (GPU): the whole ReceiverDB build on the device — first-fit insertion of the algebraised items into BinBundles
(receiver/apsu/receiver_db.cpp:330-438) + every BinBundle cache (bin_bundle.cpp:934-1041) — against a sequential
restatement of insert_or_assign_worker (this file) feeding the oracle's cache build.  Which BinBundle an item lands in
depends on the arrival order (BinBundles are scanned newest first), so three orders are tested: the reference's own
(location-major, preprocess_unlabeled_data :292-301), a random one and one that interleaves the slots round-robin."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def sequential_first_fit(p, felts, cuckoo_idx):
    """insert_or_assign_worker for every bundle index, unlabeled data, empty DB: -> bundles[b] = list of BinBundles, each a
    list (len bins_per_bundle) of lists of felts."""
    F, nbins = p.felts_per_item, p.bins_per_bundle
    bundles = [[] for _ in range(p.bundle_idx_count)]
    for item, cidx in zip(felts, cuckoo_idx):
        bin_idx = int(cidx) % nbins                     # unpack_cuckoo_idx (:93-106)
        b = (int(cidx) - bin_idx) // nbins
        written = False
        for bundle in reversed(bundles[b]):             # rbegin .. rend (:370)
            largest = max(len(bundle[bin_idx + f]) + 1 for f in range(F))   # multi_insert_dry_run
            if 0 < largest < p.max_items_per_bin:       # :388-389
                for f in range(F):
                    bundle[bin_idx + f].append(int(item[f]))
                written = True
                break
        if not written:                                 # :407-432
            fresh = [[] for _ in range(nbins)]
            for f in range(F):
                fresh[bin_idx + f].append(int(item[f]))
            bundles[b].append(fresh)
    return bundles


def _items(p, order, seed):
    rng = np.random.default_rng(seed)
    F, ipb, bic = p.felts_per_item, p.items_per_bundle, p.bundle_idx_count
    cap = p.max_items_per_bin - 1
    locs = []
    for b in range(bic):
        if b == 2:
            continue  # an empty bundle index
        hot = rng.choice(ipb, size=6, replace=False)
        for k, s in enumerate(hot):
            locs += [b * ipb + int(s)] * int(cap * (0.6 + 0.55 * k))   # up to ~3.4 BinBundles deep
        locs += [b * ipb + int(s) for s in rng.integers(0, ipb, size=300)]
    locs = np.array(locs, dtype=np.uint64)
    if order == "location-major":
        locs = np.sort(locs, kind="stable")
    elif order == "random":
        rng.shuffle(locs)
    else:  # round-robin over the slots: every slot advances one item at a time
        uniq, cnt = np.unique(locs, return_counts=True)
        rounds = [uniq[cnt > r] for r in range(int(cnt.max()))]
        locs = np.concatenate(rounds)
    felts = rng.integers(0, p.t, size=(len(locs), F), dtype=np.uint64)
    return felts, locs * np.uint64(F)


def _ncoeffs(db, b, c):
    import ctypes as C
    from apsu_b200 import capi
    n = C.c_uint32()
    capi.check(capi.lib().apsu_b200_db_binbundle_ncoeffs(db._h, b, c, C.byref(n)))
    return n.value


@pytest.mark.parametrize("order", ["location-major", "random", "round-robin"])
def test_device_db_build_equals_sequential_first_fit(order):
    import apsu_b200
    p = O.Params.load("1M-4096-com")
    ctx = O.Context.from_params(p)
    felts, cidx = _items(p, order, 31)
    want = sequential_first_fit(p, felts, cidx)
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(p.to_json()), 0)
    try:
        counts = db.set_data(felts, cidx)
        assert counts == [len(x) for x in want]
        assert max(counts) >= 3 and counts[2] == 0
        odb = O.ReceiverDB(ctx, p)
        for b in range(p.bundle_idx_count):
            for c, bins in enumerate(want[b]):
                assert odb.add_bundle_from_bins(b, bins) == c
                ref = odb.bundle_coeffs(b, c)
                assert _ncoeffs(db, b, c) == len(ref), (b, c)
                for k, (_, arr) in enumerate(ref):
                    assert np.array_equal(db.bin_bundle_coeff(b, c, k), arr), (order, b, c, k)
    finally:
        db.close()


def test_set_data_rejects_bad_items():
    import apsu_b200
    p = O.Params.load("256K-512")
    db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(p.to_json()), 0)
    try:
        felts = np.zeros((4, p.felts_per_item), dtype=np.uint64)
        with pytest.raises(ValueError):
            db.set_data(felts, np.array([0, 1, 5, 10], dtype=np.uint64))  # 1 is not the first bin of a slot
        felts[2, 1] = p.t
        with pytest.raises(ValueError):
            db.set_data(felts, np.array([0, 5, 10, 15], dtype=np.uint64))  # not a field element
        felts[2, 1] = p.t - 1
        assert db.set_data(felts, np.array([0, 5, 10, 15], dtype=np.uint64)) == [1]
    finally:
        db.close()
