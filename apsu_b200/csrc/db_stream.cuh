// K1 — the DB stream: out_g = sum_j power_j ⊙ plaintext_{g,j} over NTT-form plaintexts resident in HBM
// (BatchedPlaintextPolyn::eval / eval_patstock inner loops, receiver/apsu/bin_bundle.cpp:142-149, 251-265,
// 280-294, 328-337).  This is the HBM-bound kernel of the path.  Measured facts that shape it (tools/lab,
// profiles/lab_r01_*.log, B200):
//
//  * a pure read of the DB reaches 7.2 TB/s, and a bulk-copy ring with the MACs removed 6.9-7.0 TB/s: the
//    copy pipeline is not the limit, the integer work is;
//  * IMAD.WIDE.U32 (32x32+64) issues at 32 lanes/clk/SM (quarter rate).  A 64x64 product from four of them
//    caps the kernel at 9.3 TB/s with a perfectly busy pipe and reached 3.8-4.6 TB/s in practice.  The MAC is
//    therefore done Karatsuba-style with THREE products per word and component: operands are split at bit
//    s = ceil(bits/2) into (lo, hi); lanes ll += wl*pl, hh += wh*ph, kk += (wl+wh)*(pl+ph) are plain 64-bit
//    sums and  sum w*p = ll + (kk - ll - hh)*2^s + hh*2^2s  is rebuilt and Barrett-reduced once per output
//    (canonical residue => identical to the reference's multiply_plain + add_inplace term by term);
//  * every product must be ONE instruction: `mad.lo.cc.u32` + `madc.hi.u32` on the two halves of a lane is
//    what ptxas turns into a single IMAD.WIDE.U32 with the 64-bit addend (`mad.wide.u32` and plain C are
//    split into IMAD.WIDE + IADD3 + IADD3.X);
//  * operands reach the SM through TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a shared-memory ring
//    signalled by mbarriers, issued by one producer warp.  The DB and the powers are stored TILE-MAJOR
//    ([128-column tile][term][128 words], words already split) so that the four terms of a stage are ONE 4 KB
//    copy per job and one 8 KB copy for the powers: 5 bulk copies per 24 KB stage instead of 24 (row-major
//    1 KB copies cost 15 %);
//  * plaintext tiles carry an L2 evict_first policy (each byte is read exactly once per query), the
//    ciphertext powers evict_last (re-read by every group of the same bundle index);
//  * one CTA = four accumulation jobs sharing the power words; 2 CTAs/SM with 4 x 24 KB stages each.
#pragma once
#include "device_ctx.hpp"
#include "eval_kernels.cuh"

namespace apsu_b200 {

constexpr int kKtCols = 128;   // coefficients per tile
constexpr int kKtTS = 4;       // terms per ring stage
constexpr int kKtG = 4;        // jobs per group (share every ciphertext-power load)
constexpr int kKtThreads = kKtCols + 32; // 4 consumer warps + one producer warp
constexpr int kKtStageWords = kKtTS * (2 + kKtG) * kKtCols; // [P: TS*2 tiles][W job 0: TS tiles] .. [W job G-1]
constexpr size_t kt_smem_bytes(int stages) { return (size_t)stages * kKtStageWords * 8 + 2 * stages * 8 + 16; }
// launch shape used by the engine: 4 stages x 24 KB, 2 CTAs/SM (tools/lab: 6.4 TB/s; 3 stages x 3 CTAs 6.2)
constexpr int kKtStages = 4;
constexpr int kKtCtasPerSm = 2;
// Work items are ordered (tile block, group, tile in block): the CTAs running at any moment work on the same
// kKtTileBlock tiles across many groups, so the power tiles of those columns (pstride*2 KB each, shared by every
// group of a bundle index) stay in L2 however long the sums are.  With tiles fastest over ALL tiles the reuse
// distance was a whole group sweep: fine for 16M-4096 (88 KB per tile, 50 MB per sweep) but not for 256M-4096
// (620 KB per tile, 357 MB per sweep), where the powers were re-read from HBM for every group.
constexpr u32 kKtTileBlock = 16; // divides L*N/128 for every N >= 2048
__device__ __forceinline__ void kt_item(u32 item, u32 n_groups, u32 &group, u32 &tile)
{
    const u32 per_block = n_groups * kKtTileBlock, tb = item / per_block, rem = item - tb * per_block;
    group = rem / kKtTileBlock;
    tile = tb * kKtTileBlock + (rem - group * kKtTileBlock);
}

// One group = up to kKtG accumulation jobs over the same ciphertext powers (power j+1 multiplies term j).
// job k: out_k[c][l][n] = sum_{j < nterms_k} power_j[c][l][n] * w_k[j][l][n]   (mod q_l)
struct KtGroup {
    const u64 *w[kKtG];  // tile-major split plaintexts: word (tile, term, col) at w[k] + ((size_t)tile*wstride[k] + term)*128 + col
    u32 wstride[kKtG];   // plaintext rows per tile in the buffer job k reads
    u32 nterms[kKtG];
    u32 out_idx[kKtG];   // arena index of [2][L][N]
    u32 p_idx;           // arena index of the tile-major split power table: (tile, term, comp, col) at
                         // A + p_idx*N + (((size_t)tile*pstride + term)*2 + comp)*128 + col
    u32 pstride;         // terms per tile in the power table
    u32 njobs;
    u32 max_terms;
    u32 ragged;          // jobs of different length (or fewer than kKtG jobs)
};

__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "LAB_DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ u64 l2_policy_evict_first()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 l2_policy_evict_last()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar, u64 policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
                 : "memory");
}

// Karatsuba lanes of one accumulator
struct AccK {
    u64 ll, kk, hh;
};
// lane += x*y as ONE IMAD.WIDE.U32 with the 64-bit addend (see the header comment)
__device__ __forceinline__ void madw(u64 &lane, u32 x, u32 y)
{
    u32 lo = (u32)lane, hi = (u32)(lane >> 32);
    asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t"
        "madc.hi.u32 %1, %2, %3, %1;"
        : "+r"(lo), "+r"(hi)
        : "r"(x), "r"(y));
    lane = ((u64)hi << 32) | lo;
}
__device__ __forceinline__ void mac_k(AccK &a, u32 wl, u32 wh, u32 ws, u32 pl, u32 ph, u32 ps)
{
    madw(a.ll, wl, pl);
    madw(a.kk, ws, ps);
    madw(a.hh, wh, ph);
}
// value of the lanes modulo q (canonical)
__device__ __forceinline__ u64 reduce_k(const AccK &a, const DMod &m, int s)
{
    const u64 mid = a.kk - a.ll - a.hh;
    u64 lo = a.ll, hi = 0;
    u64 t = mid << s;
    lo += t;
    hi += (mid >> (64 - s)) + (lo < t);
    t = a.hh << (2 * s);
    lo += t;
    hi += (a.hh >> (64 - 2 * s)) + (lo < t);
    // the fold period keeps the sum below 2^(62-b) products of b-bit residues, i.e. below 2^(64+sh): one-word Barrett
    return barrett_prod(lo, hi, m);
}
// replaces the lanes by the ones of the reduced value (keeps the 64-bit lanes from overflowing on long sums)
__device__ __forceinline__ void fold_k(AccK &a, const DMod &m, int s)
{
    const u64 r = reduce_k(a, m, s);
    const u64 lo = r & ((1ull << s) - 1), hi = r >> s;
    a.ll = lo;
    a.hh = 0;
    a.kk = lo + hi; // mid = kk - ll - hh = hi
}

// one ring stage in registers
struct KtRegs {
    u64 p[kKtTS][2], w[kKtG][kKtTS];
};
__device__ __forceinline__ void kt_load(KtRegs &r, const u64 *sb)
{
#pragma unroll
    for (int h = 0; h < kKtTS; h++) {
        r.p[h][0] = sb[(h * 2) * kKtCols];
        r.p[h][1] = sb[(h * 2 + 1) * kKtCols];
#pragma unroll
        for (int k = 0; k < kKtG; k++) r.w[k][h] = sb[(kKtTS * 2 + k * kKtTS + h) * kKtCols];
    }
}
// MASKED: term t0+h of job k only counts while t0+h < nt[k] (ragged groups and the last, partial stage); the
// shared-memory words of absent terms are stale and are multiplied by zero
// lo + hi as a THREE-input add with a run-time zero: ptxas turns a plain two-input 32-bit add into IMAD.IADD to
// "balance" the pipes, but the FMA-heavy pipe (IMAD.WIDE) is exactly what bounds this kernel; IADD3 runs on the ALU
__device__ __forceinline__ u32 add3(u32 a, u32 b, u32 zero) { return a + b + zero; }

template <bool MASKED>
__device__ __forceinline__ void kt_mac(const KtRegs &r, AccK (&acc)[kKtG][2], const u32 (&nt)[kKtG], u32 t0, u32 zero)
{
#pragma unroll
    for (int h = 0; h < kKtTS; h++) {
        const u32 p0l = (u32)r.p[h][0], p0h = (u32)(r.p[h][0] >> 32), p1l = (u32)r.p[h][1], p1h = (u32)(r.p[h][1] >> 32);
        const u32 p0s = add3(p0l, p0h, zero), p1s = add3(p1l, p1h, zero);
#pragma unroll
        for (int k = 0; k < kKtG; k++) {
            u64 w = r.w[k][h];
            if (MASKED) w = (t0 + h < nt[k]) ? w : 0ull;
            const u32 wl = (u32)w, wh = (u32)(w >> 32), ws = add3(wl, wh, zero);
            mac_k(acc[k][0], wl, wh, ws, p0l, p0h, p0s);
            mac_k(acc[k][1], wl, wh, ws, p1l, p1h, p1s);
        }
    }
}

// Persistent kernel: grid = SMs x CTAs/SM, block 160 (4 consumer warps + 1 producer warp).  Work item =
// (group, 128-coefficient tile) in the order of kt_item; a CTA walks items blockIdx.x, +gridDim.x, ... and the
// ring never drains between items (the producer prefetches the next item while the consumers reduce and
// store the current one).
// split = bit position the operands were split at; fold_stages = ring stages between lane folds (host:
// largest count for which the kk lane cannot overflow, from the bit size of the largest prime); zero = 0 (see add3).
template <int STAGES, int MINB>
__global__ void __launch_bounds__(kKtThreads, MINB)
k_db_mac_kt(u64 *A, const KtGroup *__restrict__ groups, u32 n_groups, LevelConsts c, int N, int split, u32 fold_stages, u32 zero)
{
    pdl_enter();
    extern __shared__ __align__(128) u64 smem[];
    u64 *ring = smem;                              // [stage][kKtStageWords]
    u64 *full = smem + STAGES * kKtStageWords;     // [stage]
    u64 *empty = full + STAGES;                    // [stage]

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kKtCols / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const u32 n_tiles = (u32)c.L * N / kKtCols;
    const u32 n_items = n_groups * n_tiles;
    const size_t LN = (size_t)c.L * N;
    u32 it = 0; // stage counter, runs across work items identically in the producer and the consumers

    if (tid >= kKtCols) {
        // ---------------- producer warp ----------------
        if (tid != kKtCols) return;
        const u64 pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
        for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
            u32 tile, gi;
            kt_item(item, n_groups, gi, tile);
            const KtGroup *g = &groups[gi];
            const u32 max_terms = g->max_terms;
            u32 nt[kKtG];
            const u64 *wk[kKtG];
#pragma unroll
            for (int k = 0; k < kKtG; k++) {
                nt[k] = g->nterms[k];
                wk[k] = g->w[k] + (size_t)tile * g->wstride[k] * kKtCols;
            }
            const u64 *pk = A + (size_t)g->p_idx * N + (size_t)tile * g->pstride * 2 * kKtCols;
            const u32 nst = (max_terms + kKtTS - 1) / kKtTS;
            for (u32 st = 0; st < nst; st++, it++) {
                const int s = it % STAGES;
                const u32 use = it / STAGES;
                if (use) mbar_wait(&empty[s], (use - 1) & 1);
                const u32 t0 = st * kKtTS;
                const u32 ntp = min((u32)kKtTS, max_terms - t0);
                u32 nk[kKtG], rows = 2 * ntp;
#pragma unroll
                for (int k = 0; k < kKtG; k++) {
                    nk[k] = nt[k] > t0 ? min((u32)kKtTS, nt[k] - t0) : 0u;
                    rows += nk[k];
                }
                mbar_expect_tx(&full[s], rows * kKtCols * 8);
                u64 *sb = ring + (size_t)s * kKtStageWords;
                bulk_g2s(sb, pk + (size_t)t0 * 2 * kKtCols, ntp * 2 * kKtCols * 8, &full[s], pol_keep);
#pragma unroll
                for (int k = 0; k < kKtG; k++)
                    if (nk[k]) bulk_g2s(sb + (kKtTS * 2 + k * kKtTS) * kKtCols, wk[k] + (size_t)t0 * kKtCols, nk[k] * kKtCols * 8, &full[s], pol_stream);
            }
        }
        return;
    }

    // ---------------- consumer warps: one (prime, coefficient) column per thread ----------------
    const bool lane0 = (tid & 31) == 0;
    for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
        u32 tile, gi;
        kt_item(item, n_groups, gi, tile);
        const KtGroup *g = &groups[gi];
        const u32 max_terms = __ldg(&g->max_terms);
        const bool ragged = __ldg(&g->ragged) != 0;
        u32 nt[kKtG];
#pragma unroll
        for (int k = 0; k < kKtG; k++) nt[k] = __ldg(&g->nterms[k]);
        const u32 col0 = tile * kKtCols;
        const DMod m = c.q[col0 / N];
        AccK acc[kKtG][2];
#pragma unroll
        for (int k = 0; k < kKtG; k++) acc[k][0] = acc[k][1] = AccK{ 0, 0, 0 };
        const u32 nst = (max_terms + kKtTS - 1) / kKtTS;
        const u32 nst_plain = ragged ? 0u : max_terms / kKtTS; // stages whose terms all exist for all jobs
        // stages run in blocks of at most fold_stages; the lanes are folded between blocks (never inside the
        // hot loop: for the reference's parameter sets a whole inner polynomial fits one block)
        for (u32 b0 = 0; b0 < nst; b0 += fold_stages) {
            const u32 b1 = min(nst, b0 + fold_stages), p1 = min(b1, max(nst_plain, b0));
            // unrolled by two: ptxas leaves the accumulators of the fused IMAD.WIDE chain in different
            // registers than they started in, and the moves that repair that are per loop trip, not per stage
#pragma unroll 2
            for (u32 st = b0; st < p1; st++, it++) {
                const int s = it % STAGES;
                mbar_wait(&full[s], (it / STAGES) & 1);
                KtRegs r;
                kt_load(r, ring + (size_t)s * kKtStageWords + tid);
                __syncwarp();
                if (lane0) mbar_arrive(&empty[s]); // the stage is in registers: hand it back before the MACs
                kt_mac<false>(r, acc, nt, 0, zero);
            }
#pragma unroll 1
            for (u32 st = p1; st < b1; st++, it++) {
                const int s = it % STAGES;
                mbar_wait(&full[s], (it / STAGES) & 1);
                KtRegs r;
                kt_load(r, ring + (size_t)s * kKtStageWords + tid);
                __syncwarp();
                if (lane0) mbar_arrive(&empty[s]);
                kt_mac<true>(r, acc, nt, st * kKtTS, zero);
            }
            if (b1 < nst) {
#pragma unroll
                for (int k = 0; k < kKtG; k++) {
                    fold_k(acc[k][0], m, split);
                    fold_k(acc[k][1], m, split);
                }
            }
        }
        const u32 njobs = __ldg(&g->njobs);
        const u32 col = col0 + tid;
#pragma unroll
        for (int k = 0; k < kKtG; k++) {
            if ((u32)k < njobs) {
                u64 *o = A + (size_t)__ldg(&g->out_idx[k]) * N + col;
                o[0] = reduce_k(acc[k][0], m, split);
                o[LN] = reduce_k(acc[k][1], m, split);
            }
        }
    }
}

// ---- layout kernels ----
// standard [rows][L*N] -> tile-major split [L*N/128][rows][128].  grid (L*N/128, ceil(rows/8)), block 128 x 8 rows
__global__ void __launch_bounds__(128) k_pack_tile(const u64 *__restrict__ src, u64 *__restrict__ dst, u32 rows, u32 LN, int split)
{
    const u32 tile = blockIdx.x, c = threadIdx.x;
    const u32 r0 = blockIdx.y * 8;
#pragma unroll
    for (u32 r = r0; r < r0 + 8; r++)
        if (r < rows) dst[((size_t)tile * rows + r) * kKtCols + c] = split_word(src[(size_t)r * LN + (size_t)tile * kKtCols + c], split);
}
// ciphertext powers from the arena into the tile-major split table one bundle index streams against:
// dst[((tile*T + t)*2 + comp)*128 + c] = split(A[(src[t*2+comp] + l)*N + n]),  l*N + n = tile*128 + c.
// src[t*2+comp] = arena index of prime 0 of component comp of term t (its L primes are consecutive polynomials).
// grid (L*N/128, ceil(T*2 / kPackPer), tables), block 128; table z: sources src[z*T*2 ..], destination arena index
// dst_idx[z].  Every thread moves kPackPer words, all loads issued before the first store: with one word per thread the
// launch was 67 584 CTAs of ten instructions each and bound by the CTA launch rate (38 us, 72 us with the two
// griddepcontrol instructions of pdl_enter in front), not by the 140 MB it moves.
constexpr u32 kPackPer = 8;
__global__ void __launch_bounds__(128) k_pack_powers(u64 *A, const u32 *__restrict__ src, const u32 *__restrict__ dst_idx, u32 T, int N, int split)
{
    pdl_enter();
    const u32 tile = blockIdx.x, tc0 = blockIdx.y * kPackPer, z = blockIdx.z, c = threadIdx.x;
    const size_t col = (size_t)tile * kKtCols + c; // l*N + n: consecutive primes are consecutive polynomials
    u64 *dst = A + (size_t)dst_idx[z] * N;
    u64 w[kPackPer];
#pragma unroll
    for (u32 k = 0; k < kPackPer; k++)
        if (tc0 + k < T * 2) w[k] = A[(size_t)src[(size_t)z * T * 2 + tc0 + k] * N + col];
#pragma unroll
    for (u32 k = 0; k < kPackPer; k++)
        if (tc0 + k < T * 2) dst[((size_t)tile * T * 2 + tc0 + k) * kKtCols + c] = split_word(w[k], split);
}

} // namespace apsu_b200
