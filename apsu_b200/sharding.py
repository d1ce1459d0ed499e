"""BinBundle sharding across the GPUs of one box.

BinBundles are independent units (receiver/apsu/receiver_ddh.cpp:340-364): each needs only the query powers
of its bundle index and its own mask.  A rank owns a contiguous run of the (bundle_idx, cache_idx) list so
that it touches as few bundle indices as possible (it recomputes the powers of every index it owns — no
collective on the data path).  The exchanges are the query broadcast and the result gather and, when there are more ranks than bundle indices, the
all-gather of the ciphertext powers between the ranks that share a bundle index and split its PowersDag
(collective C2), all done with torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests)."""
from __future__ import annotations


def shard_bundles(degrees, world: int):
    """degrees[bundle_idx][cache_idx] -> per-rank lists of (bundle_idx, cache_idx, degree), balanced by the
    number of plaintexts (degree + 1).  A rank recomputes the query powers of every bundle index it holds BinBundles of,
    so: with at least as many ranks as populated bundle indices every rank gets BinBundles of ONE index only (ranks are
    dealt to the indices in proportion to their plaintexts; the ranks of an index can then split its PowersDag, collective
    C2); with fewer ranks the (bundle_idx, cache_idx) list is cut into contiguous runs."""
    weights = [sum(d + 1 for d in row) for row in degrees]
    active = [b for b, w in enumerate(weights) if w]
    parts = [[] for _ in range(world)]
    if active and world >= len(active):
        total = sum(weights)
        ranks = {b: 1 for b in active}
        for _ in range(world - len(active)):  # one more rank to the index with the most plaintexts per rank
            b = max(active, key=lambda i: (weights[i] / ranks[i], -i))
            ranks[b] += 1
        r0 = 0
        for b in active:
            acc = 0
            for c, d in enumerate(degrees[b]):
                w = d + 1
                k = min(ranks[b] - 1, int((acc + w / 2) * ranks[b] / weights[b]))
                parts[r0 + k].append((b, c, d))
                acc += w
            r0 += ranks[b]
        assert total and r0 == world
        return parts
    flat = [(b, c, d) for b, row in enumerate(degrees) for c, d in enumerate(row)]
    total = sum(d + 1 for _, _, d in flat)
    acc = 0
    for b, c, d in flat:
        w = d + 1
        r = min(world - 1, int((acc + w / 2) * world / total)) if total else 0
        parts[r].append((b, c, d))
        acc += w
    return parts


class MultiGpu:
    """One rank of the C++ multi-GPU path (include/apsu_b200.h, apsu_b200_mgpu_*; csrc/mgpu.cu): query scatter, optional
    PowersDag split and result gather all run inside the library over NCCL.  This class only forwards — ranks are
    processes (torchrun) or threads, each with its own ReceiverDB on its own GPU."""

    def __init__(self, db, unique_id: bytes, rank: int, world: int):
        import ctypes as C
        import numpy as np
        from . import capi
        self._C, self._np, self._capi, self._L = C, np, capi, capi.lib()
        self.db, self.rank, self.world = db, rank, world
        idb = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        h = C.c_void_p()
        capi.check(self._L.apsu_b200_mgpu_create(db._h, capi.ptr(idb), rank, world, C.byref(h)))
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        import numpy as np
        from . import capi
        idb = np.zeros(128, dtype=np.uint8)
        capi.check(capi.lib().apsu_b200_mgpu_unique_id(capi.ptr(idb)))
        return idb.tobytes()

    def close(self):
        if getattr(self, "_h", None):
            self._L.apsu_b200_mgpu_destroy(self._h)
            self._h = None

    def commit(self, global_cache_idx=None, dag_split: int = -1):
        g = None if global_cache_idx is None else self._np.ascontiguousarray(global_cache_idx, dtype=self._np.uint32)
        self._capi.check(self._L.apsu_b200_mgpu_commit(self._h, self._capi.ptr(g), dag_split))
        return self.info()

    def info(self) -> dict:
        C = self._C
        total, group, exch, ver = C.c_uint32(), C.c_uint32(), C.c_int(), C.c_int()
        self._capi.check(self._L.apsu_b200_mgpu_info(self._h, C.byref(total), C.byref(group), C.byref(exch), C.byref(ver)))
        return {"total_bin_bundles": total.value, "dag_group_size": group.value, "nccl_version": ver.value,
                "dag_exchange": ["none", "ncclAllGather per level", "NVLink peer-memory stores + flag barrier"][exch.value]}

    def compute_powers(self):
        self._capi.check(self._L.apsu_b200_mgpu_compute_powers(self._h))

    def local_count(self) -> int:
        n = self._C.c_uint32()
        self._capi.check(self._L.apsu_b200_mgpu_local_count(self._h, self._C.byref(n)))
        return n.value

    def run_query_local(self, src_powers, cts, relin_keys, masks_local, out=None, bundle_idx=None, cache_idx=None):
        """apsu_b200_mgpu_run_query_local: every rank holds the query (as with shared=True) and receives the result
        ciphertexts of ITS OWN BinBundles in its own host buffer `out` [local_count][2][N]; nothing is gathered.
        Returns (out, bundle_idx, cache_idx) with the global cache indices, on every rank."""
        np, ptr = self._np, self._capi.ptr
        sp = np.ascontiguousarray(list(src_powers), dtype=np.uint32)
        n = self.local_count()
        if out is None:
            out = np.zeros((max(n, 1), 2, self.db.params.poly_modulus_degree()), dtype=np.uint64)
        if bundle_idx is None:
            bundle_idx, cache_idx = np.zeros(max(n, 1), dtype=np.uint32), np.zeros(max(n, 1), dtype=np.uint32)
        npack = 0 if masks_local is None else masks_local.shape[0]
        self._capi.check(self._L.apsu_b200_mgpu_run_query_local(
            self._h, sp, len(sp), ptr(cts), ptr(relin_keys), ptr(masks_local), npack, ptr(out), ptr(bundle_idx), ptr(cache_idx)))
        return out[:n], bundle_idx[:n], cache_idx[:n]

    def run_query(self, src_powers, cts, relin_keys, masks_local, out=None, bundle_idx=None, cache_idx=None, shared: bool = False):
        """root (rank 0): cts / relin_keys host arrays and `out` [total][2][N] (allocated when None); other ranks pass
        None — or, with shared=True, the same query (one host buffer every rank can read): every rank then uploads its own
        part and nothing is scattered.  Returns (out, bundle_idx, cache_idx) on root, None elsewhere."""
        np, ptr = self._np, self._capi.ptr
        sp = np.ascontiguousarray(list(src_powers), dtype=np.uint32)
        root = self.rank == 0
        if root and out is None:
            total = self.info()["total_bin_bundles"]
            out = np.zeros((max(total, 1), 2, self.db.params.poly_modulus_degree()), dtype=np.uint64)
        if root and bundle_idx is None:
            bundle_idx, cache_idx = np.zeros(out.shape[0], dtype=np.uint32), np.zeros(out.shape[0], dtype=np.uint32)
        npack = 0 if masks_local is None else masks_local.shape[0]
        fn = self._L.apsu_b200_mgpu_run_query_shared if shared else self._L.apsu_b200_mgpu_run_query
        have = root or shared
        self._capi.check(fn(
            self._h, sp, len(sp), ptr(cts) if have else None, ptr(relin_keys) if have else None, ptr(masks_local), npack,
            ptr(out) if root else None, ptr(bundle_idx) if root else None, ptr(cache_idx) if root else None))
        return (out, bundle_idx, cache_idx) if root else None


def broadcast_query(tensors, src: int = 0):
    """query ciphertexts + relinearisation keys from the rank that received them (C1 of SURVEY.md §2.2)."""
    import torch.distributed as dist
    for t in tensors:
        dist.broadcast(t, src)


def gather_results(local, counts, N: int, dst: int = 0):
    """result ciphertexts [n_local][2][N] of every rank -> [sum(counts)][2][N] on dst, in rank order (C3)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    pad = max(max(counts), 1)
    buf = torch.zeros((pad, 2, N), dtype=local.dtype, device=local.device)
    if local.shape[0]:
        buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in counts] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat([o[:n] for o, n in zip(out, counts)], dim=0)


class ResultGatherer:
    """gather_results with every buffer allocated once (the call is on the per-query path): padded per-rank device
    buffers, the NCCL gather, and a pinned host buffer on `dst` for the final device-to-host copy."""

    def __init__(self, counts, N: int, device, dst: int = 0):
        import torch
        import torch.distributed as dist
        self.counts, self.N, self.dst, self.rank = list(counts), N, dst, dist.get_rank()
        self.pad = max(max(counts), 1)
        self.buf = torch.zeros((self.pad, 2, N), dtype=torch.int64, device=device)
        self.out = [torch.empty_like(self.buf) for _ in counts] if self.rank == dst else None
        pin = device is not None and str(device) != "cpu"
        self.host = torch.empty((sum(counts), 2, N), dtype=torch.int64, pin_memory=pin) if self.rank == dst else None

    def local_buffer(self):
        """device tensor [pad][2][N] the rank writes its results into (first counts[rank] rows)"""
        return self.buf

    def gather(self):
        """-> host tensor [sum(counts)][2][N] on dst (rank order), None elsewhere; the copies to the host are queued on
        the current stream, the caller synchronises."""
        import torch.distributed as dist
        dist.gather(self.buf, self.out, dst=self.dst)
        if self.rank != self.dst:
            return None
        o = 0
        for t, n in zip(self.out, self.counts):
            if n:
                self.host[o:o + n].copy_(t[:n], non_blocking=True)
            o += n
        return self.host


def powers_partition(parts, rank: int):
    """PowersDag split (collective C2): if every rank owns BinBundles of exactly one bundle index, the ranks owning
    the same index form a group that splits ComputePowers.  parts = shard_bundles(...).
    -> (sorted ranks of this rank's group, index of `rank` in it, all groups) ; groups of one rank mean no split."""
    owner = []
    for part in parts:
        idx = sorted({b for (b, _, _) in part})
        if len(idx) != 1:
            return [rank], 0, [[r] for r in range(len(parts))]
        owner.append(idx[0])
    groups = {}
    for r, b in enumerate(owner):
        groups.setdefault(b, []).append(r)
    all_groups = [groups[b] for b in sorted(groups)]
    mine = groups[owner[rank]]
    return mine, mine.index(rank), all_groups


# One exchange (event hand-over to NCCL's stream, a two-rank all-gather of 3-8 MB, hand-back) costs about 0.1 ms
# per DAG level on B200/NVLink; halving the products of a level saves less than that unless the level holds a few
# hundred ciphertext products.  Measured, 8 GPUs, 16M-4096 (66 products per bundle index, depth 3): split 1.52 ms per
# query in the free-running loop against 1.31 ms recomputing (1.97 against 2.10 ms end to end).
MIN_PRODUCTS_TO_SPLIT = 128


def worth_splitting(n_products: int, group_size: int) -> bool:
    """split the PowersDag of a bundle index over `group_size` ranks only when it is large enough to pay for the
    per-level all-gathers (e.g. the 256M parameter sets: 311-325 products per bundle index)."""
    return group_size > 1 and n_products >= MIN_PRODUCTS_TO_SPLIT


class _DeviceRegion:
    """a device buffer known by address, for torch.as_tensor (CUDA array interface)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def allgather_region(full, index: int, size: int, group=None):
    """`full` = flat tensor of `size` equal chunks, this rank's data in chunk `index`: fills the other chunks."""
    import torch.distributed as dist
    words = full.numel() // size
    mine = full[index * words:(index + 1) * words].clone()
    dist.all_gather(list(full.view(size, words).unbind(0)), mine, group=group)


_REGION_VIEWS = {}


def exchange_powers(regions, index: int, size: int, group=None, engine_stream: int | None = None):
    """all-gather of one DAG level between the `size` ranks of a group: every region is a device buffer of `size`
    chunks, this rank's products in chunk `index` (apsu_b200_powers_exchange_regions).  In place (NCCL's in-place
    all-gather: the send buffer is this rank's chunk of the receive buffer); the tensor views of a region are built
    once, the call is on the per-query path of every rank."""
    import torch
    import torch.distributed as dist
    # The collective is issued on torch's CURRENT stream while the products were written (and will be read) on the
    # engine's stream (apsu_b200_ctx_get_stream): order the two in both directions unless they are the same stream.
    # The C++ path (MultiGpu / apsu_b200_mgpu_*) issues the all-gather on the engine's stream itself.
    cur = torch.cuda.current_stream()
    ext = None
    if engine_stream is not None and engine_stream != cur.cuda_stream:
        ext = torch.cuda.ExternalStream(engine_stream)
        cur.wait_stream(ext)
    try:
        _exchange_regions(regions, index, size, group)
    finally:
        if ext is not None:
            ext.wait_stream(cur)


def _exchange_regions(regions, index: int, size: int, group=None):
    import torch
    import torch.distributed as dist
    for ptr, chunk_bytes in regions:
        key = (ptr, chunk_bytes, index, size)
        views = _REGION_VIEWS.get(key)
        if views is None:
            full = torch.as_tensor(_DeviceRegion(ptr, chunk_bytes * size), device="cuda")
            words = chunk_bytes // 8
            views = _REGION_VIEWS[key] = (full, full[index * words:(index + 1) * words])
        dist.all_gather_into_tensor(views[0], views[1], group=group)
