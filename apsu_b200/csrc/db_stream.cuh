// K1 — the DB stream: out_g = sum_j power_j ⊙ plaintext_{g,j} over NTT-form plaintexts resident in HBM
// (BatchedPlaintextPolyn::eval / eval_patstock inner loops, receiver/apsu/bin_bundle.cpp:142-149, 251-265,
// 280-294).  This is the HBM-bound kernel of the path; everything about it is shaped by that:
//
//  * operands reach the SM through TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a 4-stage
//    shared-memory ring signalled by mbarriers, issued by one producer warp — the amount of HBM traffic
//    in flight is set by the ring depth, not by registers or occupancy;
//  * plaintext tiles carry an L2 evict_first policy (each byte is read exactly once per query), the
//    ciphertext powers evict_last (re-read by every group of the same bundle index);
//  * one CTA = G accumulation jobs sharing the two power words of each term;
//  * B200 has no 64x64 multiplier: a 128-bit multiply-accumulate done with mul.lo/mul.hi costs ~15
//    instructions and would make the kernel issue-bound at about the HBM rate.  Both operands are split
//    at bit 30 instead and the four 32x32->64 partial products are summed into three 64-bit lanes
//    (weights 2^0, 2^30, 2^60) with IMAD.WIDE and no carry chains between lanes; lanes are renormalised
//    every `norm_period` terms and the residue is produced once per output with one Barrett reduction.
//    Canonical outputs => identical to multiply_plain + add_inplace term by term;
//  * the DB stores every NTT-form plaintext word already split ("packed": low 30 bits in the low half,
//    the rest in the high half of the 64-bit word), so the split costs no instructions in the stream.
#pragma once
#include "device_ctx.hpp"
#include "eval_kernels.cuh"

namespace apsu_b200 {

constexpr int kStreamCols = 128;                    // coefficients per CTA tile
constexpr int kStreamStages = 4;
constexpr int kStreamConsumerWarps = kStreamCols / 32;
constexpr int kStreamThreads = kStreamCols + 32;    // consumers + one producer warp
constexpr int kStreamTileBytes = kStreamCols * 8;


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

// three 64-bit lanes of weight 2^0, 2^30, 2^60
struct Acc3 {
    u64 ll, mid, hh;
};
__device__ __forceinline__ void mac3(Acc3 &a, u32 wl, u32 wh, u32 pl, u32 ph)
{
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.ll) : "r"(wl), "r"(pl));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(wl), "r"(ph));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(wh), "r"(pl));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.hh) : "r"(wh), "r"(ph));
}
__device__ __forceinline__ void normalize3(Acc3 &a)
{
    a.mid += a.ll >> 30;
    a.ll &= 0x3FFFFFFFull;
    a.hh += a.mid >> 30;
    a.mid &= 0x3FFFFFFFull;
}
// value of the three lanes modulo q (lanes normalised, total < 2^128)
__device__ __forceinline__ u64 reduce3(const Acc3 &a, const DMod &m)
{
    u64 lo = a.ll, hi = 0;
    u64 t = a.mid << 30;
    lo += t;
    hi += (a.mid >> 34) + (lo < t);
    t = a.hh << 60;
    lo += t;
    hi += (a.hh >> 4) + (lo < t);
    return barrett128(lo, hi, m);
}

// a stage holds TWO consecutive terms: [term a: p0 | p1 | w_0..w_{G-1}] [term b: same]
template <int G>
struct StreamCfg {
    static constexpr int term_words = (2 + G) * kStreamCols;
    static constexpr int stage_words = 2 * term_words;
    static constexpr size_t smem_bytes = (size_t)kStreamStages * stage_words * 8 + 2 * kStreamStages * 8 + 16;
};

// two terms into one accumulator.  The MACs of a lane are adjacent so that ptxas folds each pair of
// products into one IADD3 / IADD3.X (three-input adds with two carries): 4 add instructions per MAC
// instead of 6 (the compiler never keeps the 64-bit addend inside IMAD.WIDE on sm_100a).
__device__ __forceinline__ void mac3x2(Acc3 &a, u32 wla, u32 wha, u32 pla, u32 pha, u32 wlb, u32 whb, u32 plb, u32 phb)
{
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.ll) : "r"(wla), "r"(pla));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.ll) : "r"(wlb), "r"(plb));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(wla), "r"(pha));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(wha), "r"(pla));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(wlb), "r"(phb));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(whb), "r"(plb));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.hh) : "r"(wha), "r"(pha));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.hh) : "r"(whb), "r"(phb));
}

// consumer side of one work item: all term pairs of G jobs on one 128-coefficient tile
template <int G, bool RAGGED>
__device__ __forceinline__ void stream_consume(const MacGroup *__restrict__ g, int job0, u32 max_terms, u32 &it, u64 *ring, u64 *full,
                                               u64 *empty, Acc3 (&acc)[G][2], const DMod &m, u32 norm_pairs, u32 reduce_period)
{
    using Cfg = StreamCfg<G>;
    const int tid = threadIdx.x;
    const u32 npairs = (max_terms + 1) / 2;
    u32 since_norm = 0, since_reduce = 0;
    u32 nt[G];
    if (RAGGED) {
#pragma unroll
        for (int k = 0; k < G; k++) nt[k] = __ldg(&g->nterms[job0 + k]);
    }
    for (u32 jp = 0; jp < npairs; jp++, it++) {
        const int s = it % kStreamStages;
        mbar_wait(&full[s], (it / kStreamStages) & 1);
        const u64 *sa = ring + (size_t)s * Cfg::stage_words + tid;
        const u64 *sb = sa + Cfg::term_words;
        const bool two = 2 * jp + 1 < max_terms; // uniform; an odd tail multiplies the (stale) b operands by zero
        // copy the stage into registers and hand it back to the producer at once: the ring's job is to keep
        // HBM requests in flight, so a stage must not stay occupied while the MACs run
        const u64 p0a = sa[0], p1a = sa[kStreamCols], p0b = sb[0], p1b = sb[kStreamCols];
        u64 wa[G], wb[G];
#pragma unroll
        for (int k = 0; k < G; k++) {
            wa[k] = sa[(2 + k) * kStreamCols]; // packed: halves are the limbs
            wb[k] = sb[(2 + k) * kStreamCols];
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        const u32 p0al = (u32)p0a & 0x3FFFFFFFu, p0ah = (u32)(p0a >> 30), p1al = (u32)p1a & 0x3FFFFFFFu, p1ah = (u32)(p1a >> 30);
        const u32 p0bl = (u32)p0b & 0x3FFFFFFFu, p0bh = (u32)(p0b >> 30), p1bl = (u32)p1b & 0x3FFFFFFFu, p1bh = (u32)(p1b >> 30);
#pragma unroll
        for (int k = 0; k < G; k++) {
            u64 a = wa[k], b = wb[k];
            if (RAGGED) {
                a = (2 * jp < nt[k]) ? a : 0ull; // stale tile of a finished job
                b = (2 * jp + 1 < nt[k]) ? b : 0ull;
            } else if (!two) {
                b = 0ull;
            }
            const u32 wal = (u32)a, wah = (u32)(a >> 32), wbl = (u32)b, wbh = (u32)(b >> 32);
            mac3x2(acc[k][0], wal, wah, p0al, p0ah, wbl, wbh, p0bl, p0bh);
            mac3x2(acc[k][1], wal, wah, p1al, p1ah, wbl, wbh, p1bl, p1bh);
        }
        if (++since_norm == norm_pairs) {
            since_norm = 0;
#pragma unroll
            for (int k = 0; k < G; k++) {
                normalize3(acc[k][0]);
                normalize3(acc[k][1]);
            }
            if (++since_reduce == reduce_period) { // only for primes above 57 bits
                since_reduce = 0;
#pragma unroll
                for (int k = 0; k < G; k++)
                    for (int cc = 0; cc < 2; cc++) {
                        u64 r = reduce3(acc[k][cc], m);
                        acc[k][cc] = Acc3{ r & 0x3FFFFFFFull, r >> 30, 0 };
                    }
            }
        }
    }
}

// Persistent kernel: grid = SMs x resident CTAs, block 160 (4 consumer warps + 1 producer warp).  Work item =
// (group, slice of G jobs, 128-coefficient tile); a CTA walks items blockIdx.x, +gridDim.x, ... with tiles
// fastest, so the CTAs running at any moment read neighbouring memory and the TMA ring never drains between
// items (the producer prefetches the next item while the consumers reduce and store the current one).
// MacGroup::pad_ != 0 marks a ragged group (jobs of different length).
// norm_pairs = term PAIRS between lane renormalisations, reduce_period = renormalisations between full
// reductions; both depend only on the bit size of the largest prime (host computes them).
template <int G>
__global__ void __launch_bounds__(kStreamThreads, (G == 4 ? 4 : 3))
k_db_mac_tma(u64 *A, const MacGroup *__restrict__ groups, u32 n_groups, LevelConsts c, int N, u32 norm_pairs, u32 reduce_period)
{
    using Cfg = StreamCfg<G>;
    extern __shared__ __align__(128) u64 smem[];
    u64 *ring = smem;                                     // [stage][2][2+G][128]
    u64 *full = smem + kStreamStages * Cfg::stage_words;  // [stage]
    u64 *empty = full + kStreamStages;                    // [stage]

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStreamStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kStreamConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    constexpr u32 slices = kMacJobs / G;
    const u32 n_tiles = (u32)c.L * N / kStreamCols;
    const u32 n_items = n_groups * slices * n_tiles;
    const size_t LN = (size_t)c.L * N;
    u32 it = 0; // stage counter, runs across work items identically in the producer and the consumers

    if (tid >= kStreamCols) {
        // ---------------- producer warp ----------------
        if (tid != kStreamCols) return;
        const u64 pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
        for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
            const u32 tile = item % n_tiles, rest = item / n_tiles;
            const MacGroup *g = &groups[rest / slices];
            const int job0 = (int)(rest % slices) * G;
            if (job0 >= (int)g->njobs) continue;
            u32 nt[G], max_terms = 0;
            const u64 *cf[G];
            for (int k = 0; k < G; k++) {
                nt[k] = g->nterms[job0 + k];
                cf[k] = g->coeff[job0 + k];
                max_terms = max(max_terms, nt[k]);
            }
            const bool ragged = g->pad_ != 0;
            const u32 col0 = tile * kStreamCols;
            const u64 *pw = A + (size_t)g->pow_idx * N + col0;
            const size_t tstride = (size_t)g->pow_term_stride * N, cstride = (size_t)g->pow_comp_stride * N;
            const u32 npairs = (max_terms + 1) / 2;
            for (u32 jp = 0; jp < npairs; jp++, it++) {
                const int s = it % kStreamStages;
                const u32 use = it / kStreamStages;
                if (use) mbar_wait(&empty[s], (use - 1) & 1);
                const u32 nt2 = (2 * jp + 1 < max_terms) ? 2u : 1u;
                u32 tiles = 2 * nt2;
                for (u32 h = 0; h < nt2; h++)
                    for (int k = 0; k < G; k++) tiles += (!ragged || 2 * jp + h < nt[k]);
                mbar_expect_tx(&full[s], tiles * kStreamTileBytes);
                for (u32 h = 0; h < nt2; h++) {
                    const u32 j = 2 * jp + h;
                    u64 *st = ring + (size_t)s * Cfg::stage_words + (size_t)h * Cfg::term_words;
                    bulk_g2s(st, pw + j * tstride, kStreamTileBytes, &full[s], pol_keep);
                    bulk_g2s(st + kStreamCols, pw + j * tstride + cstride, kStreamTileBytes, &full[s], pol_keep);
                    for (int k = 0; k < G; k++)
                        if (!ragged || j < nt[k]) bulk_g2s(st + (2 + k) * kStreamCols, cf[k] + j * LN + col0, kStreamTileBytes, &full[s], pol_stream);
                }
            }
        }
        return;
    }

    // ---------------- consumer warps: one (prime, coefficient) column per thread ----------------
    for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
        const u32 tile = item % n_tiles, rest = item / n_tiles;
        const MacGroup *g = &groups[rest / slices];
        const int job0 = (int)(rest % slices) * G;
        const int njobs = (int)__ldg(&g->njobs);
        if (job0 >= njobs) continue;
        u32 max_terms = 0;
#pragma unroll
        for (int k = 0; k < G; k++) max_terms = max(max_terms, __ldg(&g->nterms[job0 + k]));
        const u32 col0 = tile * kStreamCols;
        const DMod m = c.q[col0 / N];
        Acc3 acc[G][2];
#pragma unroll
        for (int k = 0; k < G; k++) acc[k][0] = acc[k][1] = Acc3{ 0, 0, 0 };
        if (__ldg(&g->pad_) != 0)
            stream_consume<G, true>(g, job0, max_terms, it, ring, full, empty, acc, m, norm_pairs, reduce_period);
        else
            stream_consume<G, false>(g, job0, max_terms, it, ring, full, empty, acc, m, norm_pairs, reduce_period);
        const u32 col = col0 + tid;
#pragma unroll
        for (int k = 0; k < G; k++) {
            if (job0 + k < njobs) {
                normalize3(acc[k][0]);
                normalize3(acc[k][1]);
                u64 *o = A + (size_t)__ldg(&g->out_idx[job0 + k]) * N + col;
                o[0] = reduce3(acc[k][0], m);
                o[LN] = reduce3(acc[k][1], m);
            }
        }
    }
}

// in-place conversion of DB plaintext words to / from the packed form
__global__ void k_pack30(u64 *__restrict__ data, size_t count, int unpack)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) data[i] = unpack ? unpack30_word(data[i]) : pack30_word(data[i]);
}

} // namespace apsu_b200
