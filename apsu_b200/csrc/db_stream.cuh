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

template <int G>
struct StreamCfg {
    static constexpr int stage_words = (2 + G) * kStreamCols; // p0 | p1 | w_0..w_{G-1}
    static constexpr size_t smem_bytes = (size_t)kStreamStages * stage_words * 8 + 2 * kStreamStages * 8 + sizeof(MacGroup) + 16;
};

// grid (L*N/128, n_groups, kMacJobs/G), block 160 (4 consumer warps + 1 producer warp).
// A CTA handles jobs [z*G, z*G+G) of its MacGroup.  RAGGED: jobs of the group have different term counts
// (a job that ran out of terms multiplies by zero); otherwise every job has max_terms terms.
// norm_period / reduce_period depend only on the bit size of the largest prime (host computes them).
template <int G, bool RAGGED>
__global__ void __launch_bounds__(kStreamThreads)
k_db_mac_tma(u64 *A, const MacGroup *__restrict__ groups, LevelConsts c, int N, u32 norm_period, u32 reduce_period)
{
    using Cfg = StreamCfg<G>;
    extern __shared__ __align__(128) u64 smem[];
    u64 *ring = smem;                                     // [stage][2+G][128]
    u64 *full = smem + kStreamStages * Cfg::stage_words;  // [stage]
    u64 *empty = full + kStreamStages;                    // [stage]
    MacGroup *g = reinterpret_cast<MacGroup *>(empty + kStreamStages);

    const int tid = threadIdx.x;
    if (tid < (int)(sizeof(MacGroup) / 4)) reinterpret_cast<u32 *>(g)[tid] = reinterpret_cast<const u32 *>(&groups[blockIdx.y])[tid];
    if (tid == 0) {
        for (int s = 0; s < kStreamStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kStreamConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int job0 = blockIdx.z * G;
    if (job0 >= (int)g->njobs) return; // whole CTA: nothing to do for this slice of the group
    u32 max_terms = 0;
#pragma unroll
    for (int k = 0; k < G; k++) max_terms = max(max_terms, g->nterms[job0 + k]);

    const u32 col0 = blockIdx.x * kStreamCols; // l*N + n0 of this tile
    const size_t LN = (size_t)c.L * N;

    if (tid >= kStreamCols) {
        // ---------------- producer warp ----------------
        if (tid == kStreamCols) {
            const u64 pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
            const u64 *pw = A + (size_t)g->pow_idx * N + col0;
            const size_t tstride = (size_t)g->pow_term_stride * N, cstride = (size_t)g->pow_comp_stride * N;
            for (u32 j = 0; j < max_terms; j++) {
                const int s = j % kStreamStages;
                const u32 use = j / kStreamStages;
                if (use) mbar_wait(&empty[s], (use - 1) & 1);
                u32 active = G;
                if (RAGGED) {
                    active = 0;
                    for (int k = 0; k < G; k++) active += (j < g->nterms[job0 + k]);
                }
                mbar_expect_tx(&full[s], (2 + active) * kStreamTileBytes);
                u64 *st = ring + (size_t)s * Cfg::stage_words;
                bulk_g2s(st, pw + j * tstride, kStreamTileBytes, &full[s], pol_keep);
                bulk_g2s(st + kStreamCols, pw + j * tstride + cstride, kStreamTileBytes, &full[s], pol_keep);
                for (int k = 0; k < G; k++)
                    if (!RAGGED || j < g->nterms[job0 + k])
                        bulk_g2s(st + (2 + k) * kStreamCols, g->coeff[job0 + k] + j * LN + col0, kStreamTileBytes, &full[s], pol_stream);
            }
        }
        return;
    }

    // ---------------- consumer warps: one (prime, coefficient) column per thread ----------------
    const DMod m = c.q[col0 / N];
    Acc3 acc[G][2];
#pragma unroll
    for (int k = 0; k < G; k++) acc[k][0] = acc[k][1] = Acc3{ 0, 0, 0 };
    u32 since_norm = 0, since_reduce = 0;

    for (u32 j = 0; j < max_terms; j++) {
        const int s = j % kStreamStages;
        mbar_wait(&full[s], (j / kStreamStages) & 1);
        const u64 *st = ring + (size_t)s * Cfg::stage_words + tid;
        const u64 p0 = st[0], p1 = st[kStreamCols];
        const u32 p0l = (u32)p0 & 0x3FFFFFFFu, p0h = (u32)(p0 >> 30);
        const u32 p1l = (u32)p1 & 0x3FFFFFFFu, p1h = (u32)(p1 >> 30);
#pragma unroll
        for (int k = 0; k < G; k++) {
            u64 w = st[(2 + k) * kStreamCols]; // packed: halves are the two limbs
            if (RAGGED) w = (j < g->nterms[job0 + k]) ? w : 0ull; // stale tile of a finished job
            const u32 wl = (u32)w, wh = (u32)(w >> 32);
            mac3(acc[k][0], wl, wh, p0l, p0h);
            mac3(acc[k][1], wl, wh, p1l, p1h);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        if (++since_norm == norm_period) {
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
    const u32 col = col0 + tid;
#pragma unroll
    for (int k = 0; k < G; k++) {
        if (job0 + k < (int)g->njobs) {
            normalize3(acc[k][0]);
            normalize3(acc[k][1]);
            u64 *o = A + (size_t)g->out_idx[job0 + k] * N + col;
            o[0] = reduce3(acc[k][0], m);
            o[LN] = reduce3(acc[k][1], m);
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
