"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the C++ multi-GPU path (apsu_b200_mgpu_*, csrc/mgpu.cu) with
REAL NCCL between two B200s — query scatter (ncclSend/ncclRecv per bundle index + ncclBroadcast of the keys), the
PowersDag split with its per-level in-place ncclAllGather (collective C2), the unpadded result gather and the local
(ungathered) result delivery — compared
bit for bit with the CPU oracle.  The two ranks are two threads of this process, each with its own context on its own
GPU (ctypes releases the GIL, so both block inside NCCL concurrently); bench.py runs the same calls with one process
per GPU under torchrun."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from harness import Scenario

pytestmark = pytest.mark.gpu


def _gpu_count():
    import torch
    return torch.cuda.device_count()


def _run_ranks(world, fn):
    with ThreadPoolExecutor(world) as ex:
        futs = [ex.submit(fn, r) for r in range(world)]
        return [f.result(timeout=600) for f in futs]


@pytest.mark.parametrize("name,degrees,dag_split,expect_group", [
    ("1M-4096-com", [[30, 9], [20], [], [12], [18]], -1, 1),   # ranks own different bundle indices: scatter by index
    ("16M-4096", [[], [140, 45], [], []], 1, 2),                # both ranks on ONE bundle index: PowersDag split (C2)
    ("256K-512", [[63, 20, 5]], 1, 2),                          # direct evaluation, depth-1 DAG, split
    ("16M-4096", [[50], [], [46, 3], []], 0, 1),                # empty bundle indices, no split
])
def test_two_gpu_query_matches_oracle(name, degrees, dag_split, expect_group):
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    import apsu_b200
    from apsu_b200.sharding import MultiGpu, shard_bundles
    world = 2
    sc = Scenario(name, degrees, planted=4)
    bic = sc.p.bundle_idx_count
    exp = {(b, c): ct for b, c, ct in sc.db.run_query(sc.src_powers, sc.cts, sc.relin, sc.masks, threads=4).results()}
    parts = shard_bundles(degrees, world)
    uid = MultiGpu.unique_id()

    def rank_main(r):
        db = apsu_b200.ReceiverDB(apsu_b200.PSUParams.Load(sc.p.to_json()), r)
        mg = None
        try:
            local, gidx = {}, []
            for (b, c, _) in parts[r]:
                local[(b, c)] = db.add_bin_bundle(b, [a for (_, a) in sc.db.bundle_coeffs(b, c)])
                gidx.append(c)
            alpha = (max(local.values()) + 1) if local else 1
            masks = np.zeros((alpha * bic, sc.p.N), dtype=np.uint64)
            for (b, c), lc in local.items():
                masks[b + lc * bic] = sc.masks[b + c * bic]
            mg = MultiGpu(db, uid, r, world)
            info = mg.commit(gidx, dag_split)
            outs = []
            for shared in (False, False, True):  # twice: the second query replays the captured graphs and reuses the staging
                # buffers; then the shared-query call: every rank reads the query and uploads its own part
                have = shared or r == 0
                outs.append(mg.run_query(sc.src_powers, sc.cts if have else None, sc.relin if have else None, masks, shared=shared))
            if r == 0:
                for k in range(3):
                    assert np.array_equal(outs[0][k], outs[2][k])  # results and indices of the scattered and the shared call
            # local delivery: every rank receives the results of its own BinBundles in its own host buffer, nothing gathered
            loc = mg.run_query_local(sc.src_powers, sc.cts, sc.relin, masks)
            assert mg.local_count() == len(parts[r]) == loc[0].shape[0]
            return info, outs[1], loc
        finally:
            if mg is not None:
                mg.close()
            db.close()

    res = _run_ranks(world, rank_main)
    info0, (out, bidx, cidx), _ = res[0]
    print("multi-GPU info:", info0)
    assert info0["total_bin_bundles"] == len(exp)
    assert info0["dag_group_size"] == expect_group and res[1][0]["dag_group_size"] == expect_group
    got = {(int(bidx[k]), int(cidx[k])): out[k] for k in range(len(exp))}
    assert set(got) == set(exp)
    for key in exp:
        assert np.array_equal(got[key], exp[key]), key
    got_local = {}
    for r in range(world):
        lout, lb, lc = res[r][2]
        for k in range(lout.shape[0]):
            got_local[(int(lb[k]), int(lc[k]))] = lout[k]
    assert set(got_local) == set(exp)
    for key in exp:
        assert np.array_equal(got_local[key], exp[key]), key


def test_cpp_multi_gpu_facade_matches_single_gpu(tmp_path):
    """apsu::receiver::MultiGpuReceiver (C++ facade, one host thread per GPU over apsu_b200_mgpu_*) returns the same
    ResultPackages as apsu::receiver::Receiver on one GPU."""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    import json
    import pathlib
    import subprocess
    root = pathlib.Path(__file__).resolve().parent.parent
    exe = tmp_path / "test_facade"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-o", str(exe), str(root / "tests" / "cpp" / "test_facade.cpp"),
                           f"-L{root / 'apsu_b200'}", "-lapsu_b200", f"-Wl,-rpath,{root / 'apsu_b200'}"])
    table = json.loads((root / "tests" / "golden" / "parameters.json").read_text())
    pj = tmp_path / "p.json"
    pj.write_text(json.dumps(table["1M-4096-com.json"]))
    one = subprocess.check_output([str(exe), "gpu", str(pj), "40", "99"], text=True)
    two = subprocess.check_output([str(exe), "mgpu", str(pj), "40", "99", "2"], text=True)
    a = sorted(l for l in one.splitlines() if l.startswith("bundle_idx="))
    b = sorted(l for l in two.splitlines() if l.startswith("bundle_idx="))
    assert len(a) == 5 and a == b
