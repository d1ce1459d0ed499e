"""Independent pure-Python model of BLAKE2b with an explicit 64-byte parameter block (RFC 7693) and of BLAKE2Xb on
top of it (BLAKE2X specification).  hashlib.blake2b cannot express the XOF's expansion nodes (fanout = depth = 0), so
the compression function is modelled here, pinned against hashlib for every parameter block hashlib accepts, and then
used to pin the oracle's blake2xb (oracle/prng_restate.hpp).  Test infrastructure only."""
import struct

MASK = (1 << 64) - 1
IV = [0x6a09e667f3bcc908, 0xbb67ae8584caa73b, 0x3c6ef372fe94f82b, 0xa54ff53a5f1d36f1,
      0x510e527fade682d1, 0x9b05688c2b3e6c1f, 0x1f83d9abfb41bd6b, 0x5be0cd19137e2179]
SIGMA = [
    [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15], [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
    [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4], [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
    [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13], [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
    [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11], [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
    [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5], [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0],
]


def _rotr(x, n):
    return ((x >> n) | (x << (64 - n))) & MASK


def _compress(h, block, t, last):
    m = list(struct.unpack("<16Q", block))
    v = h[:] + IV[:]
    v[12] ^= t & MASK
    v[13] ^= t >> 64
    if last:
        v[14] ^= MASK

    def G(a, b, c, d, x, y):
        v[a] = (v[a] + v[b] + x) & MASK
        v[d] = _rotr(v[d] ^ v[a], 32)
        v[c] = (v[c] + v[d]) & MASK
        v[b] = _rotr(v[b] ^ v[c], 24)
        v[a] = (v[a] + v[b] + y) & MASK
        v[d] = _rotr(v[d] ^ v[a], 16)
        v[c] = (v[c] + v[d]) & MASK
        v[b] = _rotr(v[b] ^ v[c], 63)

    for r in range(12):
        s = SIGMA[r % 10]
        G(0, 4, 8, 12, m[s[0]], m[s[1]])
        G(1, 5, 9, 13, m[s[2]], m[s[3]])
        G(2, 6, 10, 14, m[s[4]], m[s[5]])
        G(3, 7, 11, 15, m[s[6]], m[s[7]])
        G(0, 5, 10, 15, m[s[8]], m[s[9]])
        G(1, 6, 11, 12, m[s[10]], m[s[11]])
        G(2, 7, 8, 13, m[s[12]], m[s[13]])
        G(3, 4, 9, 14, m[s[14]], m[s[15]])
    return [h[i] ^ v[i] ^ v[i + 8] for i in range(8)]


def param_block(digest_length=64, key_length=0, fanout=1, depth=1, leaf_length=0, node_offset=0, xof_length=0,
                node_depth=0, inner_length=0, salt=b"", person=b""):
    return struct.pack("<BBBBIIIBB14s16s16s", digest_length, key_length, fanout, depth, leaf_length, node_offset, xof_length,
                       node_depth, inner_length, b"", salt, person)


def blake2b_param(data: bytes, params: bytes, key: bytes = b"") -> bytes:
    assert len(params) == 64
    h = [IV[i] ^ w for i, w in enumerate(struct.unpack("<8Q", params))]
    if key:
        data = key.ljust(128, b"\0") + data
    t = 0
    while len(data) > 128:
        t += 128
        h = _compress(h, data[:128], t, False)
        data = data[128:]
    t += len(data)
    h = _compress(h, data.ljust(128, b"\0"), t, True)
    return struct.pack("<8Q", *h)[:params[0]]


def blake2xb(data: bytes, outlen: int, key: bytes = b"") -> bytes:
    root = blake2b_param(data, param_block(64, len(key), 1, 1, 0, 0, outlen), key)
    out, i = b"", 0
    while len(out) < outlen:
        bs = min(64, outlen - len(out))
        out += blake2b_param(root, param_block(bs, 0, 0, 0, 64, i, outlen, 0, 64))
        i += 1
    return out
