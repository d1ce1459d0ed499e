// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the random generator SEAL 3.7 uses by default
// (UniformRandomGeneratorFactory::DefaultFactory() -> Blake2xbPRNG) and of sample_poly_uniform, the two pieces
// behind (i) the masks of Receiver::RunQuery (receiver/apsu/receiver_ddh.cpp:221-225, 256-262) and (ii) the expansion
// of seeded ciphertexts / keys at load time (common/apsu/seal_object.h:161-219 -> Ciphertext::unsafe_load ->
// expand_seed).  SEAL is not in the reference tree (third-party, pinned 3.7, cmake/APSUConfig.cmake.in:44), so this
// restates published algorithms:
//   * BLAKE2b: RFC 7693 (pinned by its "abc" test vector and by Python's hashlib in tests/test_oracle.py);
//   * BLAKE2Xb: the BLAKE2X specification (root hash with xof_length set, output block i = BLAKE2b(root) with
//     node_offset = i, fanout = depth = 0, leaf_length = inner_length = 64), as in SEAL's bundled blake2xb.c;
//   * [SEAL-RECALL] Blake2xbPRNG::refill_buffer (native/src/seal/randomgen.cpp): a 4096-byte buffer filled with
//     blake2xb(out, 4096, &counter, 8, seed, 64), counter++ per refill; generate() serves buffer bytes in order;
//   * [SEAL-RECALL] sample_poly_uniform (native/src/seal/util/rlwe.cpp): bulk-fill L*N words, then per prime and
//     coefficient redraw 8 bytes while the word is >= max_multiple = 2^64-1 - ((2^64-1) mod q) - 1, reduce mod q.
// PARITY UNPINNED against SEAL for the two [SEAL-RECALL] items (tools/seal_kat/ prints the same digests from real SEAL).
// Written byte-oriented and streaming (init/update/final) on purpose: the product (apsu_b200/csrc/blake2.cuh) is a
// word-oriented, counter-addressed formulation of the same functions, so the two share no code.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace orc_prng {

struct Blake2bState {
    uint64_t h[8];
    uint64_t t[2];
    uint64_t f[2];
    uint8_t buf[128];
    size_t buflen;
    size_t outlen;
};

struct Blake2bParam { // 64 bytes, packed, little endian
    uint8_t digest_length, key_length, fanout, depth;
    uint8_t leaf_length[4];
    uint8_t node_offset[4];
    uint8_t xof_length[4];
    uint8_t node_depth, inner_length;
    uint8_t reserved[14];
    uint8_t salt[16];
    uint8_t personal[16];
};
static_assert(sizeof(Blake2bParam) == 64, "parameter block is 64 bytes");

inline uint64_t load64(const uint8_t *p)
{
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return v;
}
inline void store32(uint8_t *p, uint32_t v)
{
    for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (8 * i));
}
inline void store64(uint8_t *p, uint64_t v)
{
    for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i));
}
inline uint64_t rotr64(uint64_t w, unsigned c) { return (w >> c) | (w << (64 - c)); }

static const uint64_t kIV[8] = { 0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL };
static const uint8_t kSigma[12][16] = {
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 },
    { 11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4 }, { 7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8 },
    { 9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13 }, { 2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9 },
    { 12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11 }, { 13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10 },
    { 6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5 }, { 10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0 },
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 }
};

inline void compress(Blake2bState &S, const uint8_t block[128])
{
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) m[i] = load64(block + 8 * i);
    for (int i = 0; i < 8; i++) v[i] = S.h[i];
    for (int i = 0; i < 8; i++) v[8 + i] = kIV[i];
    v[12] ^= S.t[0];
    v[13] ^= S.t[1];
    v[14] ^= S.f[0];
    v[15] ^= S.f[1];
    auto G = [&](int r, int i, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d) {
        a = a + b + m[kSigma[r][2 * i]];
        d = rotr64(d ^ a, 32);
        c = c + d;
        b = rotr64(b ^ c, 24);
        a = a + b + m[kSigma[r][2 * i + 1]];
        d = rotr64(d ^ a, 16);
        c = c + d;
        b = rotr64(b ^ c, 63);
    };
    for (int r = 0; r < 12; r++) {
        G(r, 0, v[0], v[4], v[8], v[12]);
        G(r, 1, v[1], v[5], v[9], v[13]);
        G(r, 2, v[2], v[6], v[10], v[14]);
        G(r, 3, v[3], v[7], v[11], v[15]);
        G(r, 4, v[0], v[5], v[10], v[15]);
        G(r, 5, v[1], v[6], v[11], v[12]);
        G(r, 6, v[2], v[7], v[8], v[13]);
        G(r, 7, v[3], v[4], v[9], v[14]);
    }
    for (int i = 0; i < 8; i++) S.h[i] ^= v[i] ^ v[i + 8];
}

inline void init_param(Blake2bState &S, const Blake2bParam &P)
{
    std::memset(&S, 0, sizeof(S));
    const uint8_t *p = reinterpret_cast<const uint8_t *>(&P);
    for (int i = 0; i < 8; i++) S.h[i] = kIV[i] ^ load64(p + 8 * i);
    S.outlen = P.digest_length;
}

inline void update(Blake2bState &S, const uint8_t *in, size_t inlen)
{
    while (inlen > 0) {
        const size_t left = S.buflen, fill = 128 - left;
        if (inlen > fill) { // the buffered block is not the last one: compress it
            std::memcpy(S.buf + left, in, fill);
            S.t[0] += 128;
            if (S.t[0] < 128) S.t[1]++;
            compress(S, S.buf);
            S.buflen = 0;
            in += fill;
            inlen -= fill;
        } else {
            std::memcpy(S.buf + left, in, inlen);
            S.buflen += inlen;
            return;
        }
    }
}

inline void final(Blake2bState &S, uint8_t *out, size_t outlen)
{
    S.t[0] += S.buflen;
    if (S.t[0] < S.buflen) S.t[1]++;
    S.f[0] = ~0ULL;
    std::memset(S.buf + S.buflen, 0, 128 - S.buflen);
    compress(S, S.buf);
    uint8_t full[64];
    for (int i = 0; i < 8; i++) store64(full + 8 * i, S.h[i]);
    std::memcpy(out, full, outlen);
}

// plain (sequential-mode) BLAKE2b with optional key, RFC 7693
inline void blake2b(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen, const uint8_t *key, size_t keylen)
{
    if (!outlen || outlen > 64 || keylen > 64) throw std::invalid_argument("blake2b: bad lengths");
    Blake2bParam P;
    std::memset(&P, 0, sizeof(P));
    P.digest_length = (uint8_t)outlen;
    P.key_length = (uint8_t)keylen;
    P.fanout = 1;
    P.depth = 1;
    Blake2bState S;
    init_param(S, P);
    if (keylen) {
        uint8_t block[128] = { 0 };
        std::memcpy(block, key, keylen);
        update(S, block, 128);
    }
    update(S, in, inlen);
    final(S, out, outlen);
}

// BLAKE2Xb with key: arbitrary output length below 2^32 - 1
inline void blake2xb(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen, const uint8_t *key, size_t keylen)
{
    if (!outlen || outlen >= 0xFFFFFFFFull || keylen > 64) throw std::invalid_argument("blake2xb: bad lengths");
    Blake2bParam P;
    std::memset(&P, 0, sizeof(P));
    P.digest_length = 64;
    P.key_length = (uint8_t)keylen;
    P.fanout = 1;
    P.depth = 1;
    store32(P.xof_length, (uint32_t)outlen);
    Blake2bState S;
    init_param(S, P);
    if (keylen) {
        uint8_t block[128] = { 0 };
        std::memcpy(block, key, keylen);
        update(S, block, 128);
    }
    update(S, in, inlen);
    uint8_t root[64];
    final(S, root, 64);
    // expansion
    P.key_length = 0;
    P.fanout = 0;
    P.depth = 0;
    store32(P.leaf_length, 64);
    P.inner_length = 64;
    P.node_depth = 0;
    for (uint32_t i = 0; outlen > 0; i++) {
        const size_t bs = std::min<size_t>(outlen, 64);
        P.digest_length = (uint8_t)bs;
        store32(P.node_offset, i);
        Blake2bState C;
        init_param(C, P);
        update(C, root, 64);
        final(C, out + (size_t)i * 64, bs);
        outlen -= bs;
    }
}

// [SEAL-RECALL] seal::Blake2xbPRNG
class Blake2xbPRNG {
public:
    explicit Blake2xbPRNG(const std::array<uint64_t, 8> &seed) : seed_(seed), buf_(4096), head_(4096) {}
    void generate(size_t byte_count, uint8_t *dst)
    {
        while (byte_count) {
            const size_t cur = std::min(byte_count, buf_.size() - head_);
            std::memcpy(dst, buf_.data() + head_, cur);
            head_ += cur;
            dst += cur;
            byte_count -= cur;
            if (head_ == buf_.size()) {
                refill();
                head_ = 0;
            }
        }
    }
    uint32_t generate()
    {
        uint8_t b[4];
        generate(4, b);
        return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
    }

private:
    void refill()
    {
        uint8_t ctr[8], key[64];
        store64(ctr, counter_);
        for (int i = 0; i < 8; i++) store64(key + 8 * i, seed_[i]);
        blake2xb(buf_.data(), buf_.size(), ctr, 8, key, 64);
        counter_++;
    }
    std::array<uint64_t, 8> seed_;
    std::vector<uint8_t> buf_;
    size_t head_;
    uint64_t counter_ = 0;
};

// [SEAL-RECALL] seal::util::sample_poly_uniform (3.6+): destination [L][N]
inline void sample_poly_uniform(Blake2xbPRNG &prng, const uint64_t *moduli, size_t L, size_t N, uint64_t *dst)
{
    std::vector<uint8_t> raw(L * N * 8);
    prng.generate(raw.size(), raw.data());
    for (size_t k = 0; k < L * N; k++) dst[k] = load64(raw.data() + 8 * k);
    const uint64_t max_random = ~0ULL;
    for (size_t j = 0; j < L; j++) {
        const uint64_t q = moduli[j], max_multiple = max_random - (max_random % q) - 1;
        for (size_t i = 0; i < N; i++) {
            uint64_t r = dst[j * N + i];
            while (r >= max_multiple) {
                uint8_t b[8];
                prng.generate(8, b);
                r = load64(b);
            }
            dst[j * N + i] = r % q;
        }
    }
}

} // namespace orc_prng
