"""BinBundle sharding across the GPUs of one box.

BinBundles are independent units (receiver/apsu/receiver_ddh.cpp:340-364): each needs only the query powers
of its bundle index and its own mask.  A rank owns a contiguous run of the (bundle_idx, cache_idx) list so
that it touches as few bundle indices as possible (it recomputes the powers of every index it owns — no
collective on the data path).  The only exchanges are the query broadcast and the result gather, done with
torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests)."""
from __future__ import annotations


def shard_bundles(degrees, world: int):
    """degrees[bundle_idx][cache_idx] -> per-rank lists of (bundle_idx, cache_idx, degree), balanced by the
    number of plaintexts (degree + 1)."""
    flat = [(b, c, d) for b, row in enumerate(degrees) for c, d in enumerate(row)]
    total = sum(d + 1 for _, _, d in flat)
    parts = [[] for _ in range(world)]
    acc = 0
    for b, c, d in flat:
        w = d + 1
        r = min(world - 1, int((acc + w / 2) * world / total)) if total else 0
        parts[r].append((b, c, d))
        acc += w
    return parts


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
