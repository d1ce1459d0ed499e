// Kernels of the per-BinBundle polynomial evaluation (K1, K7, K8, K9, K10 of SURVEY.md §2.2):
//   k_db_mac_kt   — the DB stream: sum_j power_j ⊙ plaintext_j over NTT-form plaintexts (HBM-bound; db_stream.cuh)
//   k_db_mul_last — per-term last-prime products (Paterson-Stockmeyer i=0 polynomial, bin_bundle.cpp:314-324)
//   k_ms_sum_last — the per-term mod-switch of those terms, summed
//   k_finalize    — add_plain(coeff 0) + add_plain(mask) + mod-switch to the last level + clear bits
//   k_plain_lift  — fast plain lift of a coefficient-form plaintext to RNS (before its NTT)
//   k_fill_uniform, k_slot_scatter — synthetic DB fill, BatchEncoder slot permutation
#pragma once
#include "device_ctx.hpp"

namespace apsu_b200 {

constexpr int kMacThreads = 256;
constexpr int kTileCols = 128; // the DB-stream tile (db_stream.cuh)

// NTT-form DB plaintexts live TILE-MAJOR and SPLIT (db_stream.cuh): a buffer of `rows` plaintexts over L*N
// columns is [L*N/128][rows][128] words, each word split at bit `split` (low part in the low half of the
// 64-bit word, the remaining bits in the high half) — the DB-stream kernel multiplies the halves directly.
__host__ __device__ __forceinline__ u64 split_word(u64 w, int s) { return (w & ((1ull << s) - 1)) | ((w >> s) << 32); }
__host__ __device__ __forceinline__ u64 unsplit_word(u64 p, int s) { return (p & 0xFFFFFFFFull) | ((p >> 32) << s); }
__device__ __forceinline__ size_t tile_major_at(u32 rows, u32 row, size_t col) { return ((col / kTileCols) * rows + row) * kTileCols + col % kTileCols; }

// per-term products of the Paterson-Stockmeyer i=0 polynomial (bin_bundle.cpp:314-324), one job per BinBundle
struct MulTermsJob {
    const u64 *coeff; // tile-major split plaintexts of the BinBundle, `rows` per tile; terms are rows 0..nterms-1
    u32 rows;
    u32 out_idx;
    u32 pow_idx, pow_term_stride, pow_comp_stride; // powers in the standard layout (arena)
    u32 nterms;
};

// Last-prime variant: only the residues modulo the LAST prime of the level are produced,
// out[term][c][n] (one polynomial per (term, component)).  grid (N/256, nterms, n_bundles).
// Used for the PS i=0 polynomial when a mod-switch separates low and high powers: the per-term rounding of
// mod_switch_to_next only depends on each term's last-prime residue (see k_ms_sum_last).
__global__ void __launch_bounds__(kMacThreads)
k_db_mul_last(u64 *A, const MulTermsJob *__restrict__ jobs, LevelConsts c, int N, int split)
{
    pdl_enter();
    const MulTermsJob jb = jobs[blockIdx.z];
    const u32 j = blockIdx.y;
    if (j >= jb.nterms) return;
    const u32 n = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = c.L - 1;
    const DMod m = c.q[l];
    const size_t col = (size_t)l * N + n;
    const u64 w = unsplit_word(__ldcs(jb.coeff + tile_major_at(jb.rows, j, col)), split);
    const u64 *pw = A + ((size_t)jb.pow_idx + (size_t)j * jb.pow_term_stride) * N + col;
    u64 *o = A + ((size_t)jb.out_idx + (size_t)j * 2) * N + n;
    o[0] = mul_mod(pw[0], w, m);
    o[N] = mul_mod(pw[(size_t)jb.pow_comp_stride * N], w, m);
}

// sum_j mod_switch_to_next(t_j) for terms t_j given as (a) the coefficient-form SUM of all terms modulo the
// first L-1 primes and (b) each term's coefficient-form residue modulo the last prime q_k:
//   mod_switch(t_j)[i] = (t_j[i] - ((t_j[k] + half) mod q_k) mod q_i + half mod q_i) * q_k^-1   (mod q_i)
// is linear in t_j[i] once a_j = (t_j[k] + half) mod q_k is known, so
//   sum_j = (sum_j t_j[i] - sum_j (a_j mod q_i) + nterms * (half mod q_i)) * q_k^-1  (mod q_i)   — exact.
// grid (N/256, 2, n_bundles): sum_idx[b] -> [2][L][N] (component c at +c*L), last_idx[b] -> [nterms][2][N],
// dst_idx[b] -> [2][L-1][N].
__global__ void __launch_bounds__(kEwThreads)
k_ms_sum_last(u64 *A, const u32 *__restrict__ sum_idx, const u32 *__restrict__ last_idx, const u32 *__restrict__ dst_idx, u32 nterms,
              LevelConsts c, int N)
{
    pdl_enter();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 comp = blockIdx.y, b = blockIdx.z;
    const int L = c.L;
    const u64 qk = c.q[L - 1].q, half = qk >> 1;
    // sum_j (a_j mod q_i) is only needed modulo q_i, and that is (sum_j a_j) mod q_i: ONE integer sum of the a_j
    // (128 bits: nterms * q_k may exceed a word) and one reduction per prime at the end — no multiplication per term
    // (the first version reduced every a_j modulo every q_i: 2 Barrett reductions per term made this kernel
    // multiplier-bound, 91 us per 16M-4096 query)
    Acc128 sum{ 0, 0 };
    const u64 *t = A + ((size_t)last_idx[b] + comp) * N + n;
#pragma unroll 8
    for (u32 j = 0; j < nterms; j++) { // independent loads: unrolled so that eight are in flight
        const u64 a = add_mod(t[(size_t)j * 2 * N], half, qk);
        asm("add.cc.u64 %0, %0, %2;\n\taddc.u64 %1, %1, 0;" : "+l"(sum.lo), "+l"(sum.hi) : "l"(a));
    }
    u64 s[kMaxQ];
#pragma unroll
    for (int i = 0; i < kMaxQ - 1; i++) s[i] = (i + 1 < L) ? barrett128(sum.lo, sum.hi, c.q[i]) : 0;
    const u64 *x = A + ((size_t)sum_idx[b] + (size_t)comp * L) * N + n;
    u64 *o = A + ((size_t)dst_idx[b] + (size_t)comp * (L - 1)) * N + n;
#pragma unroll
    for (int i = 0; i < kMaxQ - 1; i++) {
        if (i + 1 < L) {
            const DMod m = c.q[i];
            const u64 corr = mul_mod(barrett64(nterms, m), c.half_mod[i], m); // nterms * (half mod q_i)
            u64 v = sub_mod(x[(size_t)i * N], s[i], m.q);
            v = add_mod(v, corr, m.q);
            o[(size_t)i * N] = mul_shoup(v, c.inv_qlast[i], m.q);
        }
    }
}

// ---- final assembly of one result ciphertext ----
struct FinalizeJob {
    u32 src[3];          // arena indices of size-2 ciphertexts [2][Ls][N] to add up (0xFFFFFFFF = none)
    u32 pack;            // mask index: bundle_idx + cache_idx * bundle_idx_count (receiver_ddh.cpp:346)
    const u64 *coeff0;   // constant-coefficient plaintext, coefficient form [N]
    u32 slot;            // result slot: out = results + slot*2*N
    u32 pad_;
};
constexpr u32 kNoSrc = 0xFFFFFFFFu;

// BFV scaling of one plaintext coefficient at the level described by c (multiply_add_plain_with_scaling_variant)
__device__ __forceinline__ u64 scaled_plain(u64 mcoef, const LevelConsts &c, int j, u64 t, u64 thr)
{
    // fix = floor((m * (q mod t) + (t+1)/2) / t); both factors < 2^61 so the numerator needs 128 bits
    u64 lo = mcoef * c.q_mod_t, hi = mulhi(mcoef, c.q_mod_t);
    lo += thr;
    hi += lo < thr;
    // mcoef < t and q_mod_t < t  =>  numerator < t^2 + t  =>  quotient < t + 1 fits 64 bits
    u64 fix;
    if (hi == 0) {
        fix = lo / t;
    } else {
        // 128/64 division: t < 2^61.  Long division by halves.
        unsigned __int128 num = ((unsigned __int128)hi << 64) | lo;
        fix = (u64)(num / t);
    }
    const DMod m = c.q[j];
    u64 s = mul_mod(mcoef, c.coeff_div_plain[j], m);
    return add_mod(s, barrett64(fix, m), m.q);
}

// grid (N/256, 2, n_jobs).  levels[L] = constants of the level with L primes; Ls = level of the sources.
__global__ void __launch_bounds__(kEwThreads)
k_finalize(u64 *A, const FinalizeJob *__restrict__ jobs, const LevelConsts *__restrict__ levels, const u64 *__restrict__ masks,
           u64 *__restrict__ results, int Ls, u64 t, u64 clear_mask, int N)
{
    pdl_enter();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 comp = blockIdx.y;
    const FinalizeJob jb = jobs[blockIdx.z];
    u64 v[kMaxQ];
    {
        const LevelConsts &c = levels[Ls];
#pragma unroll
        for (int i = 0; i < kMaxQ; i++) {
            if (i < Ls) {
                u64 s = 0;
                const u64 q = c.q[i].q;
#pragma unroll
                for (int k = 0; k < 3; k++)
                    if (jb.src[k] != kNoSrc) s = add_mod(s, A[((size_t)jb.src[k] + (size_t)comp * Ls + i) * N + n], q);
                v[i] = s;
            }
        }
        if (comp == 0) {
            const u64 thr = (t + 1) >> 1;
            const u64 m0 = jb.coeff0[n], m1 = masks[(size_t)jb.pack * N + n];
#pragma unroll
            for (int i = 0; i < kMaxQ; i++) {
                if (i < Ls) {
                    const u64 q = c.q[i].q;
                    v[i] = add_mod(v[i], scaled_plain(m0, c, i, t, thr), q);
                    v[i] = add_mod(v[i], scaled_plain(m1, c, i, t, thr), q);
                }
            }
        }
    }
    // mod_switch_to_next down to one prime
#pragma unroll
    for (int L = kMaxQ; L > 1; L--) { // compile-time indices keep v[] in registers
        if (L > Ls) continue;
        const LevelConsts &c = levels[L];
        const u64 ql = c.q[L - 1].q;
        const u64 a = add_mod(v[L - 1], ql >> 1, ql);
#pragma unroll
        for (int i = 0; i < kMaxQ - 1; i++) {
            if (i + 1 < L) {
                const DMod m = c.q[i];
                u64 tmp = sub_mod(reduce_known(a, m, c.last_kind[i]), c.half_mod[i], m.q);
                v[i] = mul_shoup(sub_mod(v[i], tmp, m.q), c.inv_qlast[i], m.q);
            }
        }
    }
    results[((size_t)jb.slot * 2 + comp) * N + n] = v[0] & clear_mask; // try_clear_irrelevant_bits
}

// fast plain lift: coefficient-form plaintext [N] mod t -> RNS [L][N] (then NTT'd by the caller).
// grid (N/256, L, n_plain): src plaintext p at in + rows[p]*N (rows == null: p), dst at out + p*L*N
__global__ void __launch_bounds__(kEwThreads)
k_plain_lift(const u64 *__restrict__ in, u64 *__restrict__ out, LevelConsts c, u64 t, int N, const u32 *__restrict__ rows = nullptr)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const size_t p = blockIdx.z;
    const u64 v = in[(rows ? (size_t)rows[p] : p) * N + n];
    const u64 thr = (t + 1) >> 1;
    out[(p * c.L + j) * N + n] = v >= thr ? v + (c.q[j].q - t) : v;
}

// counter-based splitmix64: word k of the stream seeded with `seed` (state after k+1 increments)
__device__ __forceinline__ u64 splitmix64_at(u64 seed, u64 k)
{
    u64 z = seed + (k + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// Synthetic BinBundle fill.  The stream visits the plaintexts in degree order (coefficient-form ones are
// N words mod t, NTT-form ones L*N words mod q_l) and word k of the stream is
// (splitmix64_at(seed, k) * modulus) >> 64 — the same words the oracle's synthetic fill produces.
// mode 0: `out` is the tile-major split NTT-form buffer of `rows` plaintexts; mode 1: the coefficient-form
// buffer [n_plain][N].  h = ps_low_degree + 1 (0 when Paterson-Stockmeyer is off).
struct FillMods {
    u64 q[kMaxQ];
    u64 t;
    int L;
};
__global__ void k_fill_db(u64 *__restrict__ out, size_t count, u64 seed, int mode, u32 h, FillMods mods, int N, u32 rows, int split)
{
    size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= count) return;
    const size_t LN = (size_t)mods.L * N;
    u64 pos, q;
    if (mode == 0) {
        size_t r = w / LN, within = w % LN;
        q = mods.q[within / N];
        if (h) {
            size_t blk = r / (h - 1), in_blk = r % (h - 1);
            pos = blk * (N + (h - 1) * LN) + N + in_blk * LN + within;
        } else {
            pos = N + r * LN + within;
        }
    } else {
        size_t r = w / N, within = w % N;
        q = mods.t;
        pos = (h ? r * (N + (size_t)(h - 1) * LN) : 0) + within;
    }
    const u64 v = mulhi(splitmix64_at(seed, pos), q);
    if (mode == 0)
        out[tile_major_at(rows, (u32)(w / LN), w % LN)] = split_word(v, split);
    else
        out[w] = v;
}

// polyn_with_roots (common/apsu/util/interpolate.cpp:27-80) for every bin of a BinBundle: P = prod (x - a) over
// the bin's roots, coefficients in degree-ascending order, written column-wise into the coefficient matrix
// M[degree][bin] that BatchedPlaintextPolyn's ctor gathers (bin_bundle.cpp:390-407); M is pre-zeroed.
// One warp per bin, the polynomial in shared memory; a root multiplies in place from the top coefficient
// down, 32 coefficients at a time (polyn[i] = polyn[i-1] - a*polyn[i] only needs the old values to its left).
// block = 32*warps, grid = ceil(nbins / warps), dynamic smem = warps * (max_deg + 2) words.
// SMALL: t < 2^32, products fit one word.
template <bool SMALL>
__global__ void k_polyn_with_roots(const u32 *__restrict__ bin_first, const u32 *__restrict__ bin_size, const u64 *__restrict__ roots, u64 *__restrict__ M,
                                   u32 nbins, u32 max_deg, DMod mt, int N, int *__restrict__ bad_input)
{
    extern __shared__ u64 poly_smem[];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const u32 bin = blockIdx.x * warps + warp;
    if (bin >= nbins) return;
    u64 *c = poly_smem + (size_t)warp * (max_deg + 2);
    const u32 d = bin_size[bin];
    const u64 *r = roots + bin_first[bin];
    const u64 t = mt.q;
    if (lane == 0) c[0] = 1;
    __syncwarp();
    for (u32 k = 0; k < d; k++) {
        const u64 a = r[k];
        if (a >= t) { // not a field element: reported by the host after the launch
            if (lane == 0) atomicExch(bad_input, 1);
            return;
        }
        const u64 neg_a = a ? t - a : 0; // negate_uint_mod
        // polynomial currently has k+1 coefficients c[0..k]; after this root k+2: c[0..k+1]
        // top coefficient first: c[k+1] = c[k] (old c[k+1] = 0)
        for (int hi = (int)k + 1; hi >= 0; hi -= 32) {
            const int i = hi - (int)lane;
            u64 v = 0;
            if (i >= 0) {
                const u64 ci = i <= (int)k ? c[i] : 0, cl = i > 0 ? c[i - 1] : 0;
                const u64 prod = SMALL ? barrett64(ci * neg_a, mt) : mul_mod(ci, neg_a, mt);
                v = add_mod(prod, cl, t);
            }
            __syncwarp();
            if (i >= 0) c[i] = v;
            __syncwarp();
        }
    }
    for (u32 i = lane; i <= d; i += 32) M[(size_t)i * N + bin] = c[i];
}

// polyn_with_roots for plain moduli below 2^30 and degrees up to 32*J - 1, the polynomial in REGISTERS: one warp
// per bin, coefficient i lives in register i/32 of lane i%32, so a root's update
// polyn[i] = polyn[i-1] - a*polyn[i] is one shuffle (the left neighbour's old value; lane 0 takes lane 31's value
// of the register below) and one 32-bit Shoup multiply-add per register: no shared memory, no barriers.  The
// Shoup quotients of 32 roots are computed by the 32 lanes at once.  The shared-memory kernel above (one step per
// 32 coefficients = two shared-memory round trips and two warp barriers, 64-bit Barrett) ran at 220 G steps/s
// and was 54 % of a BinBundle build.  block = 32*warps, grid = ceil(nbins / warps).
// one root applied to registers 0 .. A-1 (highest first: lane 0 needs the OLD value of the register below)
template <int J, int A>
__device__ __forceinline__ void polyn_apply_root(u32 (&c)[J], u32 neg_a, u32 quot, u32 lane, u32 t, u32 two_t)
{
#pragma unroll
    for (int j = (A < J ? A : J) - 1; j >= 0; j--) {
        u32 prev = __shfl_up_sync(0xffffffffu, c[j], 1);
        const u32 below = j > 0 ? __shfl_sync(0xffffffffu, c[j > 0 ? j - 1 : 0], 31) : 0u;
        if (lane == 0) prev = below;
        // c*neg_a mod t, lazy in [0, 2t); + prev < 3t < 2^32
        u32 v = c[j] * neg_a - __umulhi(c[j], quot) * t + prev;
        v = v >= two_t ? v - two_t : v;
        c[j] = v >= t ? v - t : v;
    }
}
template <int J, int A>
__device__ __forceinline__ void polyn_apply_block(u32 (&c)[J], u32 my_neg, u32 my_quot, u32 nk, u32 lane, u32 t, u32 two_t)
{
    for (u32 kk = 0; kk < nk; kk++)
        polyn_apply_root<J, A>(c, __shfl_sync(0xffffffffu, my_neg, kk), __shfl_sync(0xffffffffu, my_quot, kk), lane, t, two_t);
}
// dispatch on the number of active registers (a multiple of four, at most J)
template <int J>
__device__ __forceinline__ void polyn_apply_roots(u32 (&c)[J], u32 my_neg, u32 my_quot, u32 nk, u32 active, u32 lane, u32 t, u32 two_t)
{
    switch (active) {
#define APSU_POLYN_CASE(A) \
    case A: \
        if (A <= ((J + 3) & ~3)) polyn_apply_block<J, A>(c, my_neg, my_quot, nk, lane, t, two_t); \
        break;
        APSU_POLYN_CASE(4) APSU_POLYN_CASE(8) APSU_POLYN_CASE(12) APSU_POLYN_CASE(16) APSU_POLYN_CASE(20) APSU_POLYN_CASE(24)
        APSU_POLYN_CASE(28) APSU_POLYN_CASE(32) APSU_POLYN_CASE(36) APSU_POLYN_CASE(40) APSU_POLYN_CASE(44) APSU_POLYN_CASE(48)
        APSU_POLYN_CASE(52) APSU_POLYN_CASE(56) APSU_POLYN_CASE(60) APSU_POLYN_CASE(64)
#undef APSU_POLYN_CASE
    default: polyn_apply_block<J, J>(c, my_neg, my_quot, nk, lane, t, two_t);
    }
}

template <int J>
__global__ void __launch_bounds__(256, (J <= 16 ? 4 : J <= 44 ? 2 : 1))
k_polyn_with_roots_reg(const u32 *__restrict__ bin_first, const u32 *__restrict__ bin_size, const u64 *__restrict__ roots, u64 *__restrict__ M, u32 nbins,
                       u32 t, int N, int *__restrict__ bad_input)
{
    const u32 lane = threadIdx.x & 31, bin = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (bin >= nbins) return;
    const u32 d = bin_size[bin];
    const u64 *r = roots + bin_first[bin];
    u32 c[J];
#pragma unroll
    for (int j = 0; j < J; j++) c[j] = 0;
    if (lane == 0) c[0] = 1;
    const u32 two_t = 2 * t;
    for (u32 k0 = 0; k0 < d; k0 += 32) {
        // lane l prepares root k0 + l: -a mod t and its Shoup quotient floor(-a * 2^32 / t)
        u32 my_neg = 0, my_quot = 0;
        if (k0 + lane < d) {
            const u64 a = r[k0 + lane];
            if (a >= t) atomicExch(bad_input, 1);
            my_neg = a ? t - (u32)(a % t) : 0;
            my_quot = (u32)(((u64)my_neg << 32) / t);
        }
        const u32 nk = min(32u, d - k0);
        // registers that can hold a coefficient while these 32 roots are applied: 0 .. (k0 + 32) / 32, rounded up to
        // a multiple of four so that the update runs as one of J/4 branch-free instances
        const u32 active = min((u32)J, (((k0 >> 5) + 2) + 3) & ~3u);
        polyn_apply_roots<J>(c, my_neg, my_quot, nk, active, lane, t, two_t);
    }
#pragma unroll
    for (int j = 0; j < J; j++) {
        const u32 i = 32 * j + lane;
        if (i <= d) M[(size_t)i * N + bin] = c[j];
    }
}

// vec_to_std_block (receiver/apsu/receiver_ddh.cpp:70-92, sender/apsu/sender_ddh.cpp has the same helper): packs
// the felts of one item into a 128-bit block, returned as (low, high) words.  val(j) = felt j of the item.
template <typename F>
__device__ __forceinline__ void vec_to_std_block(F val, u32 felts_per_item, u64 t, u64 &lower, u64 &higher)
{
    u32 len = 1;
    while (((1ull << len) - 1) < t) len++;
    const u64 mask = (1ull << len) - 1, mask_lower = (1ull << (len >> 1)) - 1, mask_higher = mask - mask_lower;
    lower = higher = 0;
    if (felts_per_item & 1) {
        const u64 v = val(felts_per_item - 1);
        lower = v & mask_lower;
        higher = (v & mask_higher) >> ((len >> 1) - 1);
    }
    for (u32 pla = 0; pla + 1 < felts_per_item; pla += 2) {
        lower = (val(pla) & mask) | (lower << len);
        higher = (val(pla + 1) & mask) | (higher << len);
    }
}

// Mask generation of RunQuery (receiver/apsu/receiver_ddh.cpp:241-283), second half: the values r = prng32 % t were
// drawn by k_prng_mask_values (blake2.cuh: SEAL's blake2xb generator, as the reference uses); here every value is
// scattered to its BatchEncoder position (encode = this scatter + the inverse NTT mod t done by the caller) and the
// items' 128-bit blocks of the PEQT hand-off are packed (vec_to_std_block, :70-92); padded pairs (:247-252) get
// all-one blocks and no mask.
// grid (N/256, npack).  values/scattered: [npack][N]; blocks: [npack][items_per_bundle][2] = (low, high) words.
__global__ void __launch_bounds__(256)
k_masks_scatter_blocks(const u64 *__restrict__ values, u64 *__restrict__ scattered, u64 *__restrict__ blocks, const unsigned char *__restrict__ padded,
                       const u32 *__restrict__ map, u64 t, u32 felts_per_item, u32 items_per_bundle, int N)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t p = blockIdx.y;
    const bool pad = padded[p] != 0;
    scattered[p * N + map[i]] = values[p * N + i];
    if (i < items_per_bundle) {
        u64 lower = 0, higher = 0;
        if (pad) {
            lower = higher = ~0ull; // Block::all_one_block
        } else {
            const size_t base = p * (size_t)N + (size_t)i * felts_per_item;
            vec_to_std_block([&](u32 j) { return values[base + j]; }, felts_per_item, t, lower, higher);
        }
        blocks[(p * items_per_bundle + i) * 2] = lower;
        blocks[(p * items_per_bundle + i) * 2 + 1] = higher;
    }
}

// ---- sender side of the exchange: ResultPackage::extract (common/apsu/network/result_package.cpp:175-213) ----
// Decryptor::decrypt at the last level (one prime q0): x = c0 + c1*s, m = round(t*x/q0) mod t, and the invariant
// noise budget (bits) of the ciphertext.  prod = iNTT(NTT(c1) * s) is computed by the caller.
// grid (N/256, n_ct).  cts [n][2][N], prod [n][N] -> plain [n][N] (coefficient form), budget[n] (atomicMin).
__global__ void __launch_bounds__(256)
k_decrypt_round(const u64 *__restrict__ cts, const u64 *__restrict__ prod, u64 *__restrict__ plain, int *__restrict__ budget, DMod q0, u64 t, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t k = blockIdx.y;
    const u64 x = add_mod(cts[(k * 2) * N + i], prod[k * N + i], q0.q);
    // t*x = quo*q0 + rem, by Barrett with the two-word ratio (estimate at most 2 short)
    const u64 lo = t * x, hi = mulhi(t, x);
    u64 carry = mulhi(lo, q0.r0);
    u64 t_lo = lo * q0.r1, t_hi = mulhi(lo, q0.r1);
    u64 s1 = t_lo + carry;
    u64 tmp3 = t_hi + (s1 < t_lo);
    u64 u_lo = hi * q0.r0, u_hi = mulhi(hi, q0.r0);
    u64 s2 = s1 + u_lo;
    u64 quo = hi * q0.r1 + tmp3 + u_hi + (s2 < u_lo);
    u64 rem = lo - quo * q0.q;
    while (rem >= q0.q) {
        rem -= q0.q;
        quo++;
    }
    // round to nearest: floor((t*x + q0/2) / q0)
    u64 m = quo + (rem >= q0.q - (q0.q >> 1) ? 1 : 0);
    m = m >= t ? m - t : m;
    plain[k * N + i] = m;
    const u64 dist = rem > q0.q - rem ? q0.q - rem : rem;
    const int b = (int)(q0.sh + 1) - (dist ? 64 - __clzll((long long)dist) : 0) - 1;
    atomicMin(&budget[k], b);
}
// tmp[k][i] *= s[i]  (NTT domain, modulo q0).  grid (N/256, n_ct)
__global__ void __launch_bounds__(256) k_mul_secret(u64 *__restrict__ tmp, const u64 *__restrict__ s, DMod q0, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t k = blockIdx.y;
    tmp[k * N + i] = mul_mod(tmp[k * N + i], s[i], q0);
}
// BatchEncoder::decode gather (values[p][i] = ntt(plain)[p][map[i]]) and the items' blocks (sender_ddh.cpp:588-594)
// grid (N/256, n_ct)
__global__ void __launch_bounds__(256)
k_decode_gather(const u64 *__restrict__ ntt_plain, u64 *__restrict__ values, u64 *__restrict__ blocks, const u32 *__restrict__ map, u64 t, u32 felts_per_item,
                u32 items_per_bundle, int N)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t p = blockIdx.y;
    values[p * N + i] = ntt_plain[p * N + map[i]];
    if (blocks && i < items_per_bundle) {
        u64 lower, higher;
        const size_t base = p * (size_t)N;
        vec_to_std_block([&](u32 j) { return ntt_plain[base + map[i * felts_per_item + j]]; }, felts_per_item, t, lower, higher);
        blocks[(p * items_per_bundle + i) * 2] = lower;
        blocks[(p * items_per_bundle + i) * 2 + 1] = higher;
    }
}

// out[p] = in[rows[p]] (rows of N words).  grid (N/256, n_rows)
__global__ void k_gather_rows(const u64 *__restrict__ in, u64 *__restrict__ out, const u32 *__restrict__ rows, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t p = blockIdx.y;
    out[p * N + i] = in[(size_t)rows[p] * N + i];
}

// BatchEncoder::encode scatter: out[p][map[i]] = values[p][i]
__global__ void k_slot_scatter(const u64 *__restrict__ values, u64 *__restrict__ out, const u32 *__restrict__ map, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t p = blockIdx.y;
    out[p * N + map[i]] = values[p * N + i];
}

} // namespace apsu_b200
