"""CPU tests of the product's host side: the C-ABI library loads and exports every declared symbol, the
parameter loader / PowersDag agree with the oracle for all 36 parameter sets, error classes map to the
reference's exception types, the path fails loudly without a GPU, and BinBundle sharding + the
torch.distributed exchange work on world_size-2 gloo."""
import ctypes as C
import json
import os
import pathlib
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O

ROOT = pathlib.Path(__file__).resolve().parent.parent
TABLE = json.loads((ROOT / "tests" / "golden" / "parameters.json").read_text())


def test_library_exports_every_declared_symbol():
    from apsu_b200 import capi
    header = (ROOT / "include" / "apsu_b200.h").read_text()
    declared = set(re.findall(r"\b(apsu_b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 40
    lib = C.CDLL(str(capi.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/apsu_b200.h but not exported"
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    assert b"sm_100a" in capi.lib().apsu_b200_version()


def test_header_is_plain_c_and_cites_the_reference():
    """the drop-in boundary is a C ABI: the header compiles as C99 and every entry point's comment names the reference
    interface it replaces."""
    hdr = ROOT / "include" / "apsu_b200.h"
    subprocess.check_call(["gcc", "-x", "c", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", str(hdr)])
    text = hdr.read_text()
    assert "#include <torch" not in text and "at::Tensor" not in text  # plain pointers and sizes only
    for cite in ("receiver_ddh.cpp", "bin_bundle.cpp", "psu_params.cpp", "powers.cpp", "result_package.cpp", "receiver_db.cpp"):
        assert cite in text, cite


def test_db_stream_lane_bounds_hold_for_every_prime_size():
    """the Karatsuba lanes of the DB-stream kernel (csrc/db_stream.cuh, split / fold period as computed in
    Engine::Engine): for every prime size the kk lane cannot overflow between folds and the rebuilt sum stays below
    2^(64+sh), the precondition of the one-word Barrett of the epilogue."""
    TS = 4
    for b in range(12, 61):
        s = max(1, (b + 1) // 2)
        sum_max = (2**s - 1) + (2**max(0, b - s) - 1)
        cap = ((2**64 - 1) - 2**(s + 1)) // (sum_max * sum_max)
        assert cap >= TS, b
        terms = (cap // TS) * TS                       # terms between two folds
        q = 2**b - 1                                   # any modulus of b bits is below this
        # lanes after a fold: ll < 2^s, kk = lo + hi < 2^(s+1), hh = 0; then `terms` products
        assert 2**(s + 1) + terms * sum_max * sum_max < 2**64, b
        assert 2**s + terms * (2**s - 1)**2 < 2**64 and terms * (2**max(0, b - s) - 1)**2 < 2**64, b
        # value of the lanes: the folded residue plus `terms` products of residues
        assert q + terms * q * q < 2**(64 + b - 1), b


def test_vec_to_std_block_restatement_matches_a_hand_computed_block():
    """tests/harness.py::ref_vec_to_std_block (the checker of row f3/f4) against a value worked out by hand from
    receiver/apsu/receiver_ddh.cpp:70-92: t = 65537 -> len 17, masks 0x1FFFF / 0xFF / 0x1FF00; felts [1,2,3,4,5]."""
    sys.path.insert(0, str(ROOT / "tests"))
    from harness import ref_vec_to_std_block
    # odd count: lower = 5 & 0xFF = 5, higher = (5 & 0x1FF00) >> 7 = 0
    # pla = 0: lower = 1 | 5 << 17 = 0xA0001, higher = 2 | 0 << 17 = 2
    # pla = 2: lower = 3 | 0xA0001 << 17 = 0x1400020003, higher = 4 | 2 << 17 = 0x40004
    assert ref_vec_to_std_block([1, 2, 3, 4, 5], 5, 65537) == (0x1400020003, 0x40004)
    # even count, the shift quirk of the odd branch is not taken: [7, 9] -> lower 7, higher 9
    assert ref_vec_to_std_block([7, 9], 2, 65537) == (7, 9)
    # a high felt bit reaches `higher` through mask_higher >> (len/2 - 1): 0x1FF00 >> 7 = 0x3FE
    assert ref_vec_to_std_block([0x1FF00], 1, 65537) == (0, 0x3FE)


def test_params_and_powers_dag_match_oracle_for_all_parameter_sets():
    import apsu_b200
    for name, obj in TABLE.items():
        ref = O.Params(obj, name)
        p = apsu_b200.PSUParams.Load(json.dumps(obj))
        assert p.coeff_modulus() == ref.primes and p.plain_modulus() == ref.t, name
        assert (p.bundle_idx_count(), p.items_per_bundle(), p.bins_per_bundle(), p.item_bit_count()) == (
            ref.bundle_idx_count, ref.items_per_bundle, ref.bins_per_bundle, ref.item_bit_count), name
        assert p.query_powers() == ref.query_powers
        dag = apsu_b200.PowersDag(p)
        exp = {n["power"]: (n["depth"], n["p1"], n["p2"]) for n in O.powers_dag(ref.ps_low_degree, ref.max_items_per_bin, ref.query_powers)}
        got = {n.power: (n.depth, n.parents[0], n.parents[1]) for n in dag.nodes.values()}
        assert got == exp, name


def test_error_mapping_and_validation():
    import apsu_b200
    good = TABLE["16M-4096.json"]
    with pytest.raises(RuntimeError):  # JSON / loader problems are runtime_error in the reference
        apsu_b200.PSUParams.Load('{"table_params": {}}')
    both = json.loads(json.dumps(good))
    both["seal_params"]["plain_modulus"] = 65537
    with pytest.raises(RuntimeError):
        apsu_b200.PSUParams.Load(json.dumps(both))
    for path, value in [(("table_params", "table_size"), 6553), (("item_params", "felts_per_item"), 1),
                        (("table_params", "hash_func_count"), 9), (("query_params", "query_powers"), [1, 46]),
                        (("table_params", "max_items_per_bin"), 0)]:
        bad = json.loads(json.dumps(good))
        bad[path[0]][path[1]] = value
        with pytest.raises(ValueError):  # invalid_argument
            apsu_b200.PSUParams.Load(json.dumps(bad))


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import apsu_b200
    p = apsu_b200.PSUParams.Load(json.dumps(TABLE["256K-512.json"]))
    with pytest.raises(apsu_b200.CudaUnavailable):
        apsu_b200.ReceiverDB(p, 0)


def test_product_never_imports_the_oracle():
    for path in list((ROOT / "apsu_b200").rglob("*.py")) + list((ROOT / "apsu_b200" / "csrc").glob("*")) + list((ROOT / "apsu_b200" / "host").glob("*")):
        if path.suffix in (".py", ".cu", ".cuh", ".hpp", ".cpp", ".h"):
            text = path.read_text()
            assert "oracle" not in text.lower() or path.name in ("capi.py", "eval_kernels.cuh", "capi.cu", "apsu_b200.h", "__init__.py"), path
            assert "import oracle" not in text and "from oracle" not in text and "liborc" not in text, path


def test_cpp_facade_host_side(tmp_path):
    exe = tmp_path / "test_facade"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-o", str(exe), str(ROOT / "tests" / "cpp" / "test_facade.cpp"),
                           f"-L{ROOT / 'apsu_b200'}", "-lapsu_b200", f"-Wl,-rpath,{ROOT / 'apsu_b200'}"])
    pj = tmp_path / "p.json"
    pj.write_text(json.dumps(TABLE["16M-4096.json"]))
    out = subprocess.check_output([str(exe), "host", str(pj)], text=True)
    assert "N=8192 t=4079617 K=4 bundle_idx_count=4 items_per_bundle=1638 depth=3 sources=6 targets=72" in out
    assert "exceptions=7" in out


def test_shard_partitions_binbundles():
    from apsu_b200 import sharding
    degrees = [[1303] * 6 + [153], [1303] * 6 + [196], [1303] * 6 + [156], [1303] * 6 + [116]]
    for world in (1, 2, 3, 4, 8):
        parts = sharding.shard_bundles(degrees, world)
        flat = sorted((b, c) for part in parts for (b, c, _) in part)
        assert flat == sorted((b, c) for b, row in enumerate(degrees) for c in range(len(row)))
        loads = [sum(d + 1 for _, _, d in part) for part in parts]
        assert max(loads) <= 1.35 * (sum(loads) / world) + 1304
        # a rank touches as few bundle indices as possible
        assert max(len({b for b, _, _ in part}) for part in parts) <= max(1, -(-4 // world)) + (1 if world == 3 else 0)


_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from apsu_b200 import sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
degrees = [[5, 3, 1], [4], [2, 2]]
parts = sharding.shard_bundles(degrees, world)
N = 16
q = torch.arange(24, dtype=torch.int64) * 7 if rank == 0 else torch.zeros(24, dtype=torch.int64)
sharding.broadcast_query([q], src=0)
assert torch.equal(q, torch.arange(24, dtype=torch.int64) * 7)
mine = parts[rank]
local = torch.tensor([[b * 100 + c] * (2 * N) for (b, c, _) in mine], dtype=torch.int64).reshape(len(mine), 2, N)
gathered = sharding.gather_results(local, [len(p) for p in parts], N, dst=0)
if rank == 0:
    order = [(b, c) for p in parts for (b, c, _) in p]
    assert gathered.shape[0] == len(order)
    for k, (b, c) in enumerate(order):
        assert int(gathered[k, 0, 0]) == b * 100 + c
    print("OK")
# the preallocated variant used on the per-query path
gt = sharding.ResultGatherer([len(p) for p in parts], N, "cpu", dst=0)
gt.local_buffer()[: local.shape[0]] = local
host = gt.gather()
if rank == 0:
    assert torch.equal(host, gathered)
# PowersDag split (collective C2): both ranks own BinBundles of bundle index 0 only -> one group of two
parts1 = sharding.shard_bundles([[5, 3, 1, 4]], world)
group, index, groups = sharding.powers_partition(parts1, rank)
assert group == [0, 1] and index == rank and groups == [[0, 1]]
assert sharding.powers_partition(parts, rank)[0] == [rank]   # ranks owning several indices do not split
assert sharding.worth_splitting(311, 2) and not sharding.worth_splitting(66, 2) and not sharding.worth_splitting(311, 1)
pg = dist.new_group(group)
full = torch.full((2 * 6,), -1, dtype=torch.int64)
full[index * 6:(index + 1) * 6] = torch.arange(6) + 100 * rank
sharding.allgather_region(full, index, len(group), pg)
assert torch.equal(full, torch.cat([torch.arange(6), torch.arange(6) + 100])), full
if rank == 0:
    print("OK2")
dist.destroy_process_group()
'''


def test_sharded_exchange_on_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517")
    out = subprocess.check_output([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29517", str(script), str(ROOT)],
                                  env=env, text=True, stderr=subprocess.STDOUT, timeout=300)
    assert "OK" in out and "OK2" in out


def test_seal_wire_formats(tmp_path):
    """Row f2 (apsu_b200/host/seal_wire.hpp): SEAL object framing (expanded and seeded ciphertexts), the FlatBuffers
    envelopes of QueryRequest and ResultPackage round-tripped through reader and writer, error paths, and parms_id =
    BLAKE2b-256 of the parameter words against hashlib.  The byte layouts are SEAL-3.7 recall (the header says so)."""
    import hashlib
    import struct
    exe = tmp_path / "test_wire"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", str(exe), str(ROOT / "tests" / "cpp" / "test_wire.cpp"), "-ldl"])
    out = subprocess.check_output([str(exe)], text=True)
    assert "ok=1" in out
    words = [1, 8192, 0xfffffffff70001, 0xfffffffff78001, 0xfffffffffb4001, 0x3ffffffffc001, 4079617]
    d = hashlib.blake2b(struct.pack("<7Q", *words), digest_size=32).digest()
    assert "parms_id=" + ",".join("%016x" % x for x in struct.unpack("<4Q", d)) in out
    assert f"ct_bytes={16 + 32 + 1 + 5 * 8 + 16 + 8 + 2 * 3 * 64 * 8}" in out


def test_seal_wire_compressed_objects(tmp_path):
    """SEAL's default compr_mode when built with zstd (the vcpkg port APSU uses) is zstd: a compressed object is its
    SEALHeader + one zstd frame / zlib stream of the members.  seal_wire.hpp inflates both through the system libraries
    (dlopen); the compressed inputs are made here with Python's zlib and with libzstd through ctypes."""
    import ctypes
    import struct
    import zlib
    exe = tmp_path / "test_wire"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", str(exe), str(ROOT / "tests" / "cpp" / "test_wire.cpp"), "-ldl"])
    plain = tmp_path / "ct_none.bin"
    subprocess.check_call([str(exe), "dump", str(plain)])
    raw = plain.read_bytes()
    members = raw[16:]

    def wrap(mode, payload):
        return raw[:5] + bytes([mode]) + raw[6:8] + struct.pack("<Q", 16 + len(payload)) + payload

    files = [plain]
    z = tmp_path / "ct_zlib.bin"
    z.write_bytes(wrap(1, zlib.compress(members)))
    files.append(z)
    try:
        zs = ctypes.CDLL("libzstd.so.1")
    except OSError:
        zs = None
    if zs is not None:
        zs.ZSTD_compressBound.restype = ctypes.c_size_t
        zs.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
        zs.ZSTD_compress.restype = ctypes.c_size_t
        zs.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        cap = zs.ZSTD_compressBound(len(members))
        dst = ctypes.create_string_buffer(cap)
        n = zs.ZSTD_compress(dst, cap, members, len(members), 3)
        assert not zs.ZSTD_isError(n)
        f = tmp_path / "ct_zstd.bin"
        f.write_bytes(wrap(2, dst.raw[:n]))
        files.append(f)
    out = subprocess.check_output([str(exe), "check"] + [str(f) for f in files], text=True)
    assert out.count("parsed") == len(files) and "compr_mode=1" in out
    assert zs is None or "compr_mode=2" in out


def test_split_transform_index_model():
    """Model of csrc/ntt.cuh ntt_split_kernel in plain integers: a transform cut into C = 2 / 4 slices equals the whole
    transform when (forward) slice c = positions [c*M, (c+1)*M) runs the M-point passes with tw[(C + c) * 2^s + group]
    after the first log2 C stages, and (inverse) the positions congruent to c mod C run the M-point inverse with the
    SAME table indices tw[2^s + group] after their first log2 C stages.  (The kernel itself is checked against the
    oracle on the GPU: tests/test_gpu_ops.py.)"""
    import random
    q, logn = 12289, 6
    N = 1 << logn
    bitrev = lambda i: int(format(i, "0%db" % logn)[::-1], 2)
    psi = next(g for g in range(2, q) if pow(g, N, q) == q - 1)
    tw = [pow(psi, bitrev(i), q) for i in range(N)]
    twi = [pow(psi, -bitrev(i), q) for i in range(N)]
    ninv = pow(N, -1, q)

    def fwd(a, M=N, root=1):
        a, t, m = a[:], M, 1
        while m < M:
            t //= 2
            for i in range(m):
                w = tw[root * m + i]
                for j in range(2 * i * t, 2 * i * t + t):
                    u, v = a[j], a[j + t] * w % q
                    a[j], a[j + t] = (u + v) % q, (u - v) % q
            m *= 2
        return a

    def inv(a, M=N):
        a, t, m = a[:], 1, M // 2
        while m >= 1:
            for i in range(m):
                w = twi[m + i]
                for j in range(2 * i * t, 2 * i * t + t):
                    u, v = a[j], a[j + t]
                    a[j], a[j + t] = (u + v) % q, (u - v) * w % q
            t, m = t * 2, m // 2
        return [x * ninv % q for x in a]

    rnd = random.Random(7)
    x = [rnd.randrange(q) for _ in range(N)]
    y = fwd(x)
    assert inv(y) == x
    for LC in (1, 2):
        C = 1 << LC
        M, H = N // C, N // 2
        out = [0] * N
        for c in range(C):
            sl = []
            for j in range(M):
                xs = [x[j + i * M] for i in range(C)]
                if LC == 1:
                    t = xs[1] * tw[1] % q
                    sl.append((xs[0] - t) % q if c else (xs[0] + t) % q)
                else:
                    t2, t3 = xs[2] * tw[1] % q, xs[3] * tw[1] % q
                    A = (xs[0] - t2) % q if c & 2 else (xs[0] + t2) % q
                    B = (xs[1] - t3) % q if c & 2 else (xs[1] + t3) % q
                    v = B * tw[2 + (c >> 1)] % q
                    sl.append((A - v) % q if c & 1 else (A + v) % q)
            out[c * M:(c + 1) * M] = fwd(sl, M, C + c)
        assert out == y, LC
        out = [0] * N
        for c in range(C):
            sl = []
            for mm in range(M):
                xs = y[mm * C:(mm + 1) * C]
                if LC == 1:
                    sl.append((xs[0] + xs[1]) % q if not c else (xs[0] - xs[1]) * twi[H + mm] % q)
                else:
                    if c & 1:
                        A, B = (xs[0] - xs[1]) * twi[H + 2 * mm] % q, (xs[2] - xs[3]) * twi[H + 2 * mm + 1] % q
                    else:
                        A, B = (xs[0] + xs[1]) % q, (xs[2] + xs[3]) % q
                    sl.append((A + B) % q if not c & 2 else (A - B) * twi[H // 2 + mm] % q)
            r = inv(sl, M)
            for mm in range(M):
                out[mm * C + c] = r[mm]
        assert out == x, LC
