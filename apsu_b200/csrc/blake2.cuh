// BLAKE2b / BLAKE2Xb (RFC 7693; the XOF of the BLAKE2X paper) and, on top of it, SEAL's default
// UniformRandomGenerator (Blake2xbPRNG) as a COUNTER-ADDRESSABLE byte stream, for the two places the path
// needs SEAL's PRNG bit for bit:
//   * the masks of Receiver::RunQuery (receiver/apsu/receiver_ddh.cpp:221-225, 256-262): blake2xb PRNG seeded with
//     64 bytes of OS randomness, one generate() (32-bit word) per slot;
//   * the expansion of seeded ciphertexts / key-switching keys on load (common/apsu/seal_object.h:161-219 ->
//     seal::Ciphertext::unsafe_load -> expand_seed -> sample_poly_uniform).
// [SEAL-RECALL, SEAL 3.7 native/src/seal/randomgen.cpp, util/blake2xb.c]: the generator keeps a 4096-byte buffer;
// refill k (k = 0, 1, ...) is blake2xb(out = 4096 bytes, in = the 64-bit counter k (little endian), key = the 64-byte
// seed), and generate() hands out the buffer bytes in order.  So byte n of the stream is byte n % 4096 of refill
// n / 4096, and every 64-byte output block of a refill is ONE independent BLAKE2b compression of the refill's root
// hash — the whole stream is data parallel: one thread per 64-byte block, one CTA per refill.
#pragma once
#include "device_ctx.hpp"

namespace apsu_b200 {

constexpr int kPrngRefillBytes = 4096; // Blake2xbPRNG buffer size
constexpr int kPrngSeedWords = 8;      // prng_seed_type = std::array<uint64_t, 8>

struct PrngSeed {
    u64 w[kPrngSeedWords];
};

#define APSU_HD __host__ __device__ __forceinline__

APSU_HD u64 b2_rotr(u64 x, int n) { return (x >> n) | (x << (64 - n)); }

// one BLAKE2b compression: h (8 words) absorbs the 128-byte block m (16 little-endian words); t = bytes absorbed so
// far including this block (all our messages are far below 2^64 bytes), last = final block
APSU_HD void blake2b_compress(u64 (&h)[8], const u64 (&m)[16], u64 t, bool last)
{
    const u64 iv[8] = { 0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                        0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull };
    const unsigned char sigma[12][16] = {
        { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 },
        { 11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4 }, { 7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8 },
        { 9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13 }, { 2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9 },
        { 12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11 }, { 13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10 },
        { 6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5 }, { 10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0 },
        { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 }
    };
    u64 v[16];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = h[i], v[i + 8] = iv[i];
    v[12] ^= t;
    if (last) v[14] = ~v[14];
#define APSU_B2_G(a, b, c, d, x, y)                                                                                    \
    v[a] = v[a] + v[b] + (x);                                                                                          \
    v[d] = b2_rotr(v[d] ^ v[a], 32);                                                                                   \
    v[c] = v[c] + v[d];                                                                                                \
    v[b] = b2_rotr(v[b] ^ v[c], 24);                                                                                   \
    v[a] = v[a] + v[b] + (y);                                                                                          \
    v[d] = b2_rotr(v[d] ^ v[a], 16);                                                                                   \
    v[c] = v[c] + v[d];                                                                                                \
    v[b] = b2_rotr(v[b] ^ v[c], 63);
#pragma unroll
    for (int r = 0; r < 12; r++) {
        const unsigned char *s = sigma[r];
        APSU_B2_G(0, 4, 8, 12, m[s[0]], m[s[1]])
        APSU_B2_G(1, 5, 9, 13, m[s[2]], m[s[3]])
        APSU_B2_G(2, 6, 10, 14, m[s[4]], m[s[5]])
        APSU_B2_G(3, 7, 11, 15, m[s[6]], m[s[7]])
        APSU_B2_G(0, 5, 10, 15, m[s[8]], m[s[9]])
        APSU_B2_G(1, 6, 11, 12, m[s[10]], m[s[11]])
        APSU_B2_G(2, 7, 8, 13, m[s[12]], m[s[13]])
        APSU_B2_G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef APSU_B2_G
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

// parameter block words 0..2 (the rest is zero: no salt, no personalisation):
//   byte 0 digest_length, 1 key_length, 2 fanout, 3 depth, 4-7 leaf_length, 8-11 node_offset, 12-15 xof_length,
//   16 node_depth, 17 inner_length
APSU_HD void blake2b_init(u64 (&h)[8], unsigned digest_len, unsigned key_len, unsigned fanout, unsigned depth, unsigned leaf_len, unsigned node_offset,
                          unsigned xof_len, unsigned node_depth, unsigned inner_len)
{
    const u64 iv[8] = { 0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                        0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull };
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = iv[i];
    h[0] ^= (u64)digest_len | ((u64)key_len << 8) | ((u64)fanout << 16) | ((u64)depth << 24) | ((u64)leaf_len << 32);
    h[1] ^= (u64)node_offset | ((u64)xof_len << 32);
    h[2] ^= (u64)node_depth | ((u64)inner_len << 8);
}

// root hash of refill `counter`: BLAKE2b-512 keyed with the seed over the 8-byte counter, xof_length = 4096
// (blake2xb_init_key + blake2xb_update + the first half of blake2xb_final)
APSU_HD void prng_refill_root(const PrngSeed &seed, u64 counter, u64 (&root)[8])
{
    blake2b_init(root, 64, 64, 1, 1, 0, 0, kPrngRefillBytes, 0, 0);
    u64 m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = i < 8 ? seed.w[i] : 0; // the key, padded to one block
    blake2b_compress(root, m, 128, false);
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = 0;
    m[0] = counter;
    blake2b_compress(root, m, 128 + 8, true);
}

// 64-byte output block `block` (0..63) of a refill with root hash `root` (second half of blake2xb_final)
APSU_HD void prng_refill_block(const u64 (&root)[8], unsigned block, u64 (&out)[8])
{
    blake2b_init(out, 64, 0, 0, 0, 64, block, kPrngRefillBytes, 0, 64);
    u64 m[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m[i] = i < 8 ? root[i] : 0;
    blake2b_compress(out, m, 64, true);
}

#ifdef __CUDACC__
// Raw stream: refill counter0 + blockIdx.x, 64 threads = the 64 output blocks; out = u64 words in stream order.
// Thread 0 hashes the root once per refill (two compressions), every thread then does one.
__global__ void __launch_bounds__(64) k_prng_stream(u64 *__restrict__ out, PrngSeed seed, u64 counter0, size_t n_words)
{
    __shared__ u64 root_s[8];
    if (threadIdx.x == 0) {
        u64 root[8];
        prng_refill_root(seed, counter0 + blockIdx.x, root);
#pragma unroll
        for (int i = 0; i < 8; i++) root_s[i] = root[i];
    }
    __syncthreads();
    u64 root[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) root[i] = root_s[i];
    prng_refill_block(root, threadIdx.x, o);
    const size_t base = ((size_t)blockIdx.x * 64 + threadIdx.x) * 8;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (base + i < n_words) out[base + i] = o[i];
}

// Mask values of RunQuery (receiver_ddh.cpp:256-262): the s-th non-padded (cache_idx, bundle_idx) pair takes the
// 32-bit words [s*N, (s+1)*N) of the stream, r = word % plain_modulus.  grid (N/1024, npack), 64 threads: one refill
// (1024 words) per CTA.  seq[p] = s for pack index p, or 0xFFFFFFFF for a padded pair (values zeroed, no stream used).
__global__ void __launch_bounds__(64) k_prng_mask_values(u64 *__restrict__ values, const u32 *__restrict__ seq, PrngSeed seed, u64 t, int N)
{
    const size_t p = blockIdx.y;
    const u32 s = seq[p];
    u64 *dst = values + p * (size_t)N + (size_t)blockIdx.x * 1024 + (size_t)threadIdx.x * 16;
    if (s == 0xFFFFFFFFu) {
#pragma unroll
        for (int i = 0; i < 16; i++) dst[i] = 0;
        return;
    }
    __shared__ u64 root_s[8];
    if (threadIdx.x == 0) {
        u64 root[8];
        prng_refill_root(seed, (u64)s * (unsigned)(N / 1024) + blockIdx.x, root);
#pragma unroll
        for (int i = 0; i < 8; i++) root_s[i] = root[i];
    }
    __syncthreads();
    u64 root[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) root[i] = root_s[i];
    prng_refill_block(root, threadIdx.x, o);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        dst[2 * i] = (u64)(u32)o[i] % t;
        dst[2 * i + 1] = (u64)(u32)(o[i] >> 32) % t;
    }
}

// sample_poly_uniform (SEAL 3.7 util/rlwe.cpp) for `n_polys` seeded polynomials of L primes each: polynomial k takes
// the first L*N 64-bit words of the stream of seeds[k]; word (j, i) is accepted when below
// max_multiple_j = 2^64 - 1 - ((2^64 - 1) mod q_j) - 1 and then reduced mod q_j.  A rejected word is redrawn from the
// words FOLLOWING the bulk, in (j, i) order (sequential in SEAL) — for SEAL's primes (just below a power of two) that
// happens with probability ~2^-37 per word; rejected positions are only recorded here (rej_count[k], first kMaxRej
// positions in rej_pos[k]) and resolved by k_prng_fix_rejects.
// grid (L*N/512, n_polys), 64 threads: one refill (512 words) per CTA.  dst_idx[k] = arena index of polynomial k's
// [L][N] block.
constexpr int kMaxRej = 1024;
struct UniformMods {
    DMod q[kMaxQ];
    u64 max_multiple[kMaxQ];
    int L;
};
__global__ void __launch_bounds__(64)
k_prng_sample_uniform(u64 *A, const u32 *__restrict__ dst_idx, const PrngSeed *__restrict__ seeds, UniformMods mods, int N, u32 *__restrict__ rej_count,
                      u32 *__restrict__ rej_pos)
{
    const u32 k = blockIdx.y;
    __shared__ u64 root_s[8];
    if (threadIdx.x == 0) {
        u64 root[8];
        prng_refill_root(seeds[k], blockIdx.x, root);
#pragma unroll
        for (int i = 0; i < 8; i++) root_s[i] = root[i];
    }
    __syncthreads();
    u64 root[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) root[i] = root_s[i];
    prng_refill_block(root, threadIdx.x, o);
    const u32 w0 = blockIdx.x * 512 + threadIdx.x * 8; // word index inside the polynomial block [L][N]
    const int j = w0 / N;
    const DMod m = mods.q[j];
    const u64 mm = mods.max_multiple[j];
    u64 *dst = A + (size_t)dst_idx[k] * N + w0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (o[i] >= mm) {
            const u32 slot = atomicAdd(&rej_count[k], 1u);
            if (slot < (u32)kMaxRej) rej_pos[(size_t)k * kMaxRej + slot] = w0 + i;
        }
        dst[i] = barrett64(o[i], m);
    }
}
// resolves the recorded rejections of polynomial k sequentially: positions ascending, each takes the next stream
// words (from word L*N on) until one is accepted.  One thread per polynomial (the lists are almost always empty).
__global__ void k_prng_fix_rejects(u64 *A, const u32 *__restrict__ dst_idx, const PrngSeed *__restrict__ seeds, UniformMods mods, int N,
                                   u32 *__restrict__ rej_count, u32 *__restrict__ rej_pos, u32 n_polys, int *__restrict__ overflow)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_polys) return;
    const u32 n = rej_count[k];
    if (!n) return;
    if (n > (u32)kMaxRej) {
        atomicExch(overflow, 1);
        return;
    }
    u32 *pos = rej_pos + (size_t)k * kMaxRej;
    for (u32 a = 1; a < n; a++) { // insertion sort
        const u32 x = pos[a];
        u32 b = a;
        for (; b > 0 && pos[b - 1] > x; b--) pos[b] = pos[b - 1];
        pos[b] = x;
    }
    u64 next = (u64)mods.L * N; // stream word index of the next redraw
    u64 cur_refill = ~0ull, root[8], blk[8];
    unsigned cur_block = ~0u;
    for (u32 a = 0; a < n; a++) {
        const u32 w = pos[a];
        const int j = w / N;
        u64 r;
        do {
            const u64 refill = next / 512;
            const unsigned block = (unsigned)((next % 512) / 8);
            if (refill != cur_refill) {
                prng_refill_root(seeds[k], refill, root);
                cur_refill = refill;
                cur_block = ~0u;
            }
            if (block != cur_block) {
                prng_refill_block(root, block, blk);
                cur_block = block;
            }
            r = blk[next % 8];
            next++;
        } while (r >= mods.max_multiple[j]);
        A[(size_t)dst_idx[k] * N + w] = barrett64(r, mods.q[j]);
    }
    rej_count[k] = 0;
}
#endif // __CUDACC__

} // namespace apsu_b200
