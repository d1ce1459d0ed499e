// Negacyclic NTT / inverse NTT over one RNS prime, N in {2048, 4096, 8192, ...}: one CTA per
// polynomial, the whole polynomial staged in shared memory, radix-2^R butterflies done in
// registers between shared-memory exchanges (kernels K2/K3 of SURVEY.md §2.2).  The transform is bound by the
// integer-multiply pipe (ncu: fmaheavy 65-74 %, DRAM 8 %), so the butterfly is built for few IMAD.WIDE.
//
// Convention (SEAL 3.7): psi = minimal primitive 2N-th root; forward = Cooley-Tukey,
// natural-order input, bit-reversed output, out[k] = a(psi^(2*bitrev(k)+1)); inverse =
// Gentleman-Sande incl. N^-1; both return canonical residues.
// Twiddle table: tw[i] = psi^bitrev(i) in Shoup form, stage with m groups uses tw[m + group].
#pragma once
#include "device_ctx.hpp"

namespace apsu_b200 {

// How a CTA finds its polynomials: optional gather/scatter index arrays (units of one polynomial).
struct NttSrc {
    const unsigned *src_idx; // null: p
    const unsigned *dst_idx; // null: p
    int reduce_input;        // reduce every input word modulo the target modulus first
};

// Fused prologues: the element-wise kernel that used to produce a transform's input runs inside the transform, on
// the way into shared memory, so its output never makes the round trip through the arena and one launch of the
// (launch-latency-bound) ComputePowers chain disappears per fusion:
//   kNttExtend (forward):  polynomial p = (rp, j), j < L+S.  j < L: plain transform of source polynomial src_idx[rp]+j;
//                          j >= L: BEHZ steps (1)-(2) for auxiliary prime j-L computed from the L source residues
//                          (k_behz_extend, kernels.cuh), then transformed.
//   kNttTensor (inverse):  polynomial p = (o, c, j), c < 3: component c of the tensor product of the extended
//                          operands a_idx[o], b_idx[o] at modulus slot j (k_tensor), then inverse transform.
//   kNttKsMac (inverse):   polynomial p = (o, comp, I): key-switch inner product over the L digits at modulus slot I
//                          (k_ks_mac), then inverse transform.
// value n of the fused input of polynomial p (see NttFuse); A = arena base
template <int MODE>
__device__ __forceinline__ u64 fused_input(const u64 *A, const NttSrc &src, const NttFuse &f, unsigned p, unsigned n, int N, const DMod &m)
{
    if (MODE == kNttExtend) {
        const int L = f.L, LS = f.L + f.S;
        const unsigned rp = p / LS, jb = p % LS - L;
        const LevelConsts &c = *f.lc;
        const u64 *x = A + (size_t)src.src_idx[rp] * N + n;
        u64 tmp[kMaxQ];
        u32 ymt = 0;
#pragma unroll
        for (int i = 0; i < kMaxQ; i++) {
            if (i < L) {
                tmp[i] = mul_shoup(x[(size_t)i * N], c.mtilde_inv_punct_q[i], c.q[i].q);
                ymt += (u32)tmp[i] * c.q_punct_mod_mtilde[i];
            }
        }
        const u32 r = ymt * c.neg_inv_q_mod_mtilde; // arithmetic mod m_tilde = 2^32
        u64 s = 0;
        int pending = 0;
#pragma unroll
        for (int i = 0; i < kMaxQ; i++) {
            if (i < L) {
                s += mul_shoup_lazy3(tmp[i], c.ext_punct_bsk[jb][i].op, c.ext_punct_bsk[jb][i].quot, 0 - m.q);
                if (++pending == 2) s = reduce_8q(s, m.q), pending = 0;
            }
        }
        u64 rr = r;
        if (r >= 0x80000000u) rr += m.q - 0x100000000ull; // centred lift of r
        s += mul_shoup_lazy3(rr, c.ext_q_bsk[jb].op, c.ext_q_bsk[jb].quot, 0 - m.q);
        return reduce_8q(s, m.q);
    } else if (MODE == kNttTensor) {
        const int LS = f.L + f.S;
        const unsigned j = p % LS, oc = p / LS, o = oc / 3, cc = oc % 3;
        const size_t cs = (size_t)LS * N;
        const u64 *a = A + ((size_t)f.a_idx[o] + j) * N + n;
        const u64 *b = A + ((size_t)f.b_idx[o] + j) * N + n;
        if (cc == 0) return mul_mod(a[0], b[0], m);
        if (cc == 2) return mul_mod(a[cs], b[cs], m);
        const u64 a0 = a[0], a1 = a[cs], b0 = b[0], b1 = b[cs];
        Acc128 acc{ 0, 0 };
        mac128(acc, a0, b1);
        mac128(acc, a1, b0);
        return barrett_prod(acc.lo, acc.hi, m); // 2 q^2 < 2^(64+sh)
    } else {
        const int L = f.L, R = f.L + 1;
        const unsigned I = p % R, oc = p / R, o = oc >> 1, comp = oc & 1;
        const unsigned key_index = I == (unsigned)L ? f.K - 1 : I;
        const u64 *dg = A + ((size_t)f.a_idx[o] + I) * N + n;
        Acc128 acc{ 0, 0 };
        for (int J = 0; J < L; J++) mac128(acc, dg[(size_t)J * R * N], f.keys[(((size_t)J * 2 + comp) * f.K + key_index) * N + n]);
        return barrett_prod(acc.lo, acc.hi, m); // <= 5 products of reduced operands
    }
}

// Butterflies with the 3q-lazy Shoup product (modarith.cuh: mul_shoup_lazy3): forward values live in [0, 6q),
// inverse values in [0, 3q); nq = 2^64 - q, q3 = 3q (q < 2^61.4, so 6q fits a word).
// twiddle fetch: from the global table through the read-only path, or from the copy staged in shared memory (TWS)
template <bool TWS>
__device__ __forceinline__ ulonglong2 tw_at(const ulonglong2 *tw, unsigned i)
{
    return TWS ? tw[i] : __ldg(&tw[i]);
}

template <int R, bool TWS = false>
__device__ __forceinline__ void fwd_group(u64 (&x)[1 << R], const ulonglong2 *__restrict__ tw, unsigned hi, int s, u64 nq, u64 q3, unsigned root = 1u)
{
    // root: 1 for a whole transform; C + c for slice c of a transform cut into C contiguous slices after its first
    // log2 C stages (ntt_split_kernel): stage s of the slice is stage s + log2 C of the transform, groups c*2^s ...
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int d = 1 << (R - 1 - r);          // pair distance inside the register group
        const unsigned mbase = (root << (s + r)) + (hi << r);
#pragma unroll
        for (int k = 0; k < (1 << R); k++) {
            if (k & d) continue;
            ulonglong2 w = tw_at<TWS>(tw, mbase + (k >> (R - r)));
            const u64 u = csub(x[k], q3);                              // [0, 3q)
            const u64 v = mul_shoup_lazy3(x[k + d], w.x, w.y, nq);     // [0, 3q)
            x[k] = u + v;                                              // [0, 6q)
            x[k + d] = u + q3 - v;                                     // (0, 6q)
        }
    }
}

template <int R, bool TWS = false>
__device__ __forceinline__ void inv_group(u64 (&x)[1 << R], const ulonglong2 *__restrict__ tw, unsigned hi, int s, u64 nq, u64 q3)
{
#pragma unroll
    for (int r = R - 1; r >= 0; r--) {
        const int d = 1 << (R - 1 - r);
        const unsigned mbase = (1u << (s + r)) + (hi << r);
#pragma unroll
        for (int k = 0; k < (1 << R); k++) {
            if (k & d) continue;
            ulonglong2 w = tw_at<TWS>(tw, mbase + (k >> (R - r)));
            const u64 u = x[k], v = x[k + d];                          // [0, 3q)
            x[k] = csub(u + v, q3);                                    // [0, 3q)
            x[k + d] = mul_shoup_lazy3(u + q3 - v, w.x, w.y, nq);      // [0, 3q)
        }
    }
}

// stages R-1 .. 1 of the first (last executed) inverse group, i.e. inv_group without its stage-0 butterflies
template <int R>
__device__ __forceinline__ void inv_group_upper(u64 (&x)[1 << R], const ulonglong2 *__restrict__ tw, u64 nq, u64 q3)
{
#pragma unroll
    for (int r = R - 1; r >= 1; r--) {
        const int d = 1 << (R - 1 - r);
        const unsigned mbase = 1u << r;
#pragma unroll
        for (int k = 0; k < (1 << R); k++) {
            if (k & d) continue;
            ulonglong2 w = __ldg(&tw[mbase + (k >> (R - r))]);
            const u64 u = x[k], v = x[k + d];
            x[k] = csub(u + v, q3);
            x[k + d] = mul_shoup_lazy3(u + q3 - v, w.x, w.y, nq);
        }
    }
}

// lab switches (tools/lab/ntt_variants.sh): unroll factor of the shared-memory pass loop, CTAs per SM of the N = 8192
// throughput shape
#ifndef APSU_NTT_UNROLL
#define APSU_NTT_UNROLL 1
#endif
#ifndef APSU_NTT_MINB13
#define APSU_NTT_MINB13 3
#endif
constexpr int kNttUnroll = APSU_NTT_UNROLL;
// Shared-memory layout: one pad word after every 16 coefficients.  The butterflies of the late passes
// touch coefficients at strides 1..8; with the pad a half-warp's 16 eight-byte accesses fall into 16
// distinct bank pairs instead of colliding 8-way (measured: 54% of LSU wavefronts were bank conflicts).
__device__ __forceinline__ unsigned pad_idx(unsigned i) { return i + (i >> 4); }

// one pass over stages [s, s+R) on the shared-memory polynomial
template <int LOGN, int R, bool FWD, bool TWS = false>
__device__ __forceinline__ void smem_pass(u64 *sm, const ulonglong2 *__restrict__ tw, int s, u64 nq, u64 q3, unsigned root = 1u)
{
    constexpr int N = 1 << LOGN;
    const int log_stride = LOGN - s - R;
    const unsigned stride = 1u << log_stride;
#pragma unroll kNttUnroll
    for (unsigned g = threadIdx.x; g < (N >> R); g += blockDim.x) {
        unsigned lo = g & (stride - 1), hi = g >> log_stride;
        unsigned base = (hi << (LOGN - s)) + lo;
        u64 x[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); k++) x[k] = sm[pad_idx(base + k * stride)];
        if (FWD)
            fwd_group<R, TWS>(x, tw, hi, s, nq, q3, root);
        else
            inv_group<R, TWS>(x, tw, hi, s, nq, q3);
#pragma unroll
        for (int k = 0; k < (1 << R); k++) sm[pad_idx(base + k * stride)] = x[k];
    }
}

// middle passes (shared memory -> shared memory), COUNT passes of three stages starting at stage S;
// the inverse transform runs the same passes in the opposite order
template <int LOGN, bool FWD, int S, int COUNT, bool TWS = false>
struct MidRunner {
    __device__ static __forceinline__ void run(u64 *sm, const ulonglong2 *tw, u64 nq, u64 q3, unsigned root = 1u)
    {
        if (FWD) {
            smem_pass<LOGN, 3, true, TWS>(sm, tw, S, nq, q3, root);
            __syncthreads();
            MidRunner<LOGN, FWD, S + 3, COUNT - 1, TWS>::run(sm, tw, nq, q3, root);
        } else {
            MidRunner<LOGN, FWD, S + 3, COUNT - 1, TWS>::run(sm, tw, nq, q3);
            smem_pass<LOGN, 3, false, TWS>(sm, tw, S, nq, q3);
            __syncthreads();
        }
    }
};
template <int LOGN, bool FWD, int S, bool TWS>
struct MidRunner<LOGN, FWD, S, 0, TWS> {
    __device__ static __forceinline__ void run(u64 *, const ulonglong2 *, u64, u64, unsigned = 1u) {}
};

// One CTA per polynomial.  Stages are grouped as [RF | 3 | 3 | ... | 3] with RF = 2..4; the pass that touches
// global memory on the way in and the one on the way out do their butterflies straight from / to global
// memory (coalesced), so a transform costs (number of passes - 1) shared-memory round trips.
// Launch shape: DIV = coefficients per thread.  Throughput shape (batches of more than a wave): DIV = 32, i.e. N/32
// threads per CTA (each thread does four radix-8 groups per pass) and as many CTAs per SM as shared memory allows
// (3 at N = 8192): 80 registers per thread without spills and three independent CTAs to fill each other's barrier
// stalls; measured against N/16 threads x 2 CTAs (64 registers, spills): 79 -> 75 ns/poly.  Latency shapes: a batch
// that does not fill the GPU (one bundle index on one of 8 GPUs: 24..300 polynomials) takes one CTA's latency
// whatever its size (20 us at DIV = 32), so batches of at most one CTA per SM use DIV = 8 (N/8 threads, 16 us).
constexpr int ntt_min_blocks(int logn, int div)
{
    if (logn == 13 && div == 32) return APSU_NTT_MINB13;
    // CTAs per SM: shared memory allows 3 << (13 - logn) (1 at logn = 14); at most 1024 resident threads so that
    // every shape has at least 64 registers per thread
    const int by_smem = logn >= 14 ? 1 : 3 << (13 - (logn >= 14 ? 13 : logn));
    const int threads = (1 << logn) / div, by_threads = threads >= 1024 ? 1 : 1024 / threads;
    return by_smem < by_threads ? by_smem : by_threads;
}
// in and out may alias (in-place transforms, gather/scatter inside one arena): each CTA reads its whole polynomial
// before its first store and no CTA's destination is another CTA's source, so neither pointer is __restrict__.
// TWS (latency shape only): the modulus' whole twiddle table (N entries, 16 bytes each) is copied into shared memory
// with cp.async while the polynomial is being loaded, and every pass but the one on stages [0, RF) (15 entries, the
// same for all threads) takes its twiddles from there.  A transform alone on its SM has nobody to hide the L2 latency
// of its twiddle loads behind (ncu, 168 polynomials: long-scoreboard is the top stall, issue slots 33 % busy); with
// one CTA per SM the 128 KB at N = 8192 fit next to the polynomial.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
template <int LOGN, bool FWD, int DIV, int MODE = kNttPlain, bool TWS = false>
__global__ void __launch_bounds__((1 << LOGN) / DIV, ntt_min_blocks(LOGN, DIV)) ntt_kernel(const u64 *in, u64 *out, NttArgs a, NttSrc src, NttFuse fuse = NttFuse())
{
    pdl_enter();
    static_assert(!TWS || MODE == kNttPlain, "staged twiddles are implemented for the plain transform");
    constexpr int N = 1 << LOGN;
    constexpr int RF = (LOGN % 3 == 0) ? 3 : (LOGN % 3 == 1 ? 4 : 2); // 13 = 4+3+3+3: one shared-memory round trip less than 1+3+3+3+3
    constexpr int MID = (LOGN - RF - 3) / 3;
    extern __shared__ u64 sm[];
    const unsigned p = blockIdx.x;
    const int slot = p % a.pattern_len;
    const DMod m = a.mod[slot];
    const u64 q = m.q, nq = 0 - m.q, q3 = 3 * m.q;
    const ulonglong2 *tw = a.tw + ((size_t)a.table[slot] * 2 + (FWD ? 0 : 1)) * N;
    ulonglong2 *tws = reinterpret_cast<ulonglong2 *>(sm + N + (N >> 4)); // TWS: behind the padded polynomial
    if (TWS) {
        for (unsigned k = threadIdx.x; k < (unsigned)N; k += blockDim.x) cp_async16(tws + k, tw + k);
    }
    u64 *op = out + (size_t)(src.dst_idx ? src.dst_idx[p] : p) * N;
    const bool reduce = src.reduce_input != 0;
    // fused prologue: the input is computed element by element into shared memory and the first pass runs from there
    // (CTA-uniform choice: in kNttExtend the polynomials of the q primes take the plain path from their source)
    bool from_smem = false;
    const u64 *ip;
    if (MODE == kNttExtend) {
        const unsigned LS = fuse.L + fuse.S, rp = p / LS, j = p % LS;
        ip = in + ((size_t)src.src_idx[rp] + j) * N; // only used when j < L
        from_smem = j >= (unsigned)fuse.L;
    } else {
        ip = in + (size_t)((MODE == kNttPlain && src.src_idx) ? src.src_idx[p] : p) * N;
        from_smem = MODE != kNttPlain;
    }
    if (MODE != kNttPlain && from_smem) {
        // four elements per trip, all their loads issued before the first product: a lone CTA has nobody to hide a
        // load round trip per element behind (N / blockDim is a multiple of 4 in every launch shape)
        constexpr unsigned U = 4;
        for (unsigned n0 = threadIdx.x; n0 < (unsigned)N; n0 += blockDim.x * U) {
            u64 v[U];
#pragma unroll
            for (unsigned u = 0; u < U; u++) v[u] = fused_input<MODE>(in, src, fuse, p, n0 + u * blockDim.x, N, m);
#pragma unroll
            for (unsigned u = 0; u < U; u++) sm[pad_idx(n0 + u * blockDim.x)] = v[u];
        }
        __syncthreads();
    }

    if (FWD) {
        // stages [0, RF): global -> registers -> shared.  The trip count is a compile-time constant and the loads of
        // kBatch iterations are issued together: with a run-time loop every iteration waited a full memory latency
        // on its own two loads (long-scoreboard was the top stall of the forward transform)
        if (MODE != kNttPlain && from_smem) {
            smem_pass<LOGN, RF, true>(sm, tw, 0, nq, q3);
        } else {
            constexpr unsigned stride = N >> RF;
            constexpr unsigned kThreads = N / DIV, kIters = (stride + kThreads - 1) / kThreads, kBatch = (RF == 1 && kIters % 4 == 0) ? 4 : 1;
            static_assert(kIters % kBatch == 0, "first-pass batching");
#pragma unroll 1
            for (unsigned it = 0; it < kIters; it += kBatch) {
                if (stride < kThreads && threadIdx.x >= stride) break; // more threads than groups (latency shape)
                u64 x[kBatch][1 << RF];
#pragma unroll
                for (unsigned b = 0; b < kBatch; b++) {
                    const unsigned g = threadIdx.x + (it + b) * kThreads;
#pragma unroll
                    for (int k = 0; k < (1 << RF); k++) x[b][k] = ip[g + k * stride];
                }
#pragma unroll
                for (unsigned b = 0; b < kBatch; b++) {
                    const unsigned g = threadIdx.x + (it + b) * kThreads;
                    if (reduce) {
#pragma unroll
                        for (int k = 0; k < (1 << RF); k++) x[b][k] = barrett64(x[b][k], m);
                    }
                    fwd_group<RF>(x[b], tw, 0, 0, nq, q3);
#pragma unroll
                    for (int k = 0; k < (1 << RF); k++) sm[pad_idx(g + k * stride)] = x[b][k];
                }
            }
        }
        if (TWS) cp_async_wait_all();
        __syncthreads();
        MidRunner<LOGN, true, RF, MID, TWS>::run(sm, TWS ? tws : tw, nq, q3);
        // stages [LOGN-3, LOGN): shared -> registers -> global, fully reduced
        for (unsigned g = threadIdx.x; g < (N >> 3); g += blockDim.x) {
            u64 x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = sm[pad_idx(8 * g + k)];
            fwd_group<3, TWS>(x, TWS ? tws : tw, g, LOGN - 3, nq, q3);
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = csub(csub(csub(x[k], q3), 2 * q), q); // [0, 6q) -> [0, q)
            ulonglong2 *o2 = reinterpret_cast<ulonglong2 *>(op + 8 * g);
#pragma unroll
            for (int k = 0; k < 4; k++) o2[k] = make_ulonglong2(x[2 * k], x[2 * k + 1]);
        }
    } else {
        // stages [LOGN-3, LOGN) first: global -> registers -> shared (fused prologue: shared -> shared)
        if (MODE != kNttPlain) smem_pass<LOGN, 3, false>(sm, tw, LOGN - 3, nq, q3);
        for (unsigned g = threadIdx.x; MODE == kNttPlain && g < (N >> 3); g += blockDim.x) {
            u64 x[8];
            const ulonglong2 *i2 = reinterpret_cast<const ulonglong2 *>(ip + 8 * g);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                ulonglong2 v = i2[k];
                x[2 * k] = reduce ? barrett64(v.x, m) : v.x;
                x[2 * k + 1] = reduce ? barrett64(v.y, m) : v.y;
            }
            if (TWS) { // the table must have landed: N/8 groups = threads in this shape, so this runs once per thread
                cp_async_wait_all();
                __syncthreads();
            }
            inv_group<3, TWS>(x, TWS ? tws : tw, g, LOGN - 3, nq, q3);
#pragma unroll
            for (int k = 0; k < 8; k++) sm[pad_idx(8 * g + k)] = x[k];
        }
        __syncthreads();
        MidRunner<LOGN, false, RF, MID, TWS>::run(sm, TWS ? tws : tw, nq, q3);
        // stages [0, RF) last: shared -> registers -> global.  The N^-1 scaling rides on the very last stage (one
        // twiddle, psi^-(N/2)): x[k] = (u+v)*N^-1, x[k+d] = (u-v)*(w*N^-1), two lazy products instead of one lazy and
        // two exact ones
        const DShoup inv_n = a.inv_n[slot], inv_n_w = a.inv_n_w[slot];
        constexpr unsigned stride = N >> RF;
        for (unsigned g = threadIdx.x; g < stride; g += blockDim.x) {
            u64 x[1 << RF];
#pragma unroll
            for (int k = 0; k < (1 << RF); k++) x[k] = sm[pad_idx(g + k * stride)];
            inv_group_upper<RF>(x, tw, nq, q3); // stages RF-1 .. 1
            constexpr int d = 1 << (RF - 1);
#pragma unroll
            for (int k = 0; k < d; k++) {
                const u64 u = x[k], v = x[k + d];
                const u64 lo = mul_shoup_lazy3(u + v, inv_n.op, inv_n.quot, nq);
                const u64 hi = mul_shoup_lazy3(u + q3 - v, inv_n_w.op, inv_n_w.quot, nq);
                op[g + k * stride] = csub(csub(lo, 2 * q), q);
                op[g + (k + d) * stride] = csub(csub(hi, 2 * q), q);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Split transforms: C = 2^LC CTAs (one thread-block cluster) per polynomial, for batches that leave SMs idle.
//
// A lone transform is bound by ONE SM's multiplier pipe and its own dependent chains (16 us at N = 8192 however few
// polynomials the launch has), and the ComputePowers chain of one bundle index is ~30 such launches in a row.  Here
// every CTA owns a slice of M = N / C coefficients on which all but log2 C stages are an independent M-point transform:
//   forward (strides N/2 ... 1): after the first LC stages slice c = positions [c*M, (c+1)*M) is on its own.  A CTA
//     computes ITS outputs of those LC stages straight from the C inputs each depends on (C = 4: three products per
//     coefficient instead of one per two coefficients and stage -- every CTA reads the whole polynomial, redundant
//     work 8 % at C = 2 and 31 % at C = 4, spread over C SMs), then runs the M-point passes with the twiddles of
//     its sub-tree: stage s of the slice uses tw[(C + c) * 2^s + group].
//   inverse (strides 1 ... N/2): after the first LC stages the positions congruent to c mod C are on their own, and
//     stage s of that strided slice uses tw[2^s + group] of the SAME table -- the M-point inverse code unchanged.
//     Outputs go to positions m*C + c (8-byte stores at stride C: acceptable for a launch that does not fill the GPU).
// The N^-1 scaling and the reduction to canonical residues are as in ntt_kernel, so results are bit-identical.
// in and out may alias (in-place): a CTA must not store before its cluster peers have read the polynomial, hence one
// split cluster barrier (arrive after the loads were consumed, wait before the first global store).
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// output c of the first LC forward stages at slice offset j; x = the C inputs at j + i*M (canonical), result in [0, 6q)
template <int LC>
__device__ __forceinline__ u64 fwd_split_head(const u64 (&x)[1 << LC], const ulonglong2 *__restrict__ tw, unsigned c, u64 nq, u64 q3)
{
    const ulonglong2 w1 = __ldg(&tw[1]);
    if (LC == 1) {
        const u64 t = mul_shoup_lazy3(x[1], w1.x, w1.y, nq);
        return c ? x[0] + q3 - t : x[0] + t;
    } else {
        const u64 t2 = mul_shoup_lazy3(x[2], w1.x, w1.y, nq), t3 = mul_shoup_lazy3(x[3], w1.x, w1.y, nq);
        const u64 A = (c & 2) ? x[0] + q3 - t2 : x[0] + t2; // [0, 4q)
        const u64 B = (c & 2) ? x[1] + q3 - t3 : x[1] + t3;
        const ulonglong2 w2 = __ldg(&tw[2 + (c >> 1)]);
        const u64 u = csub(A, q3), v = mul_shoup_lazy3(B, w2.x, w2.y, nq);
        return (c & 1) ? u + q3 - v : u + v;
    }
}
// output c of the first LC inverse stages for the C consecutive inputs x of block mm (positions mm*C ...), in [0, 3q)
template <int LC, int LOGN>
__device__ __forceinline__ u64 inv_split_head(const u64 (&x)[1 << LC], const ulonglong2 *__restrict__ tw, unsigned mm, unsigned c, u64 nq, u64 q3)
{
    constexpr unsigned H = 1u << (LOGN - 1);
    if (LC == 1) {
        if (!c) return csub(x[0] + x[1], q3);
        const ulonglong2 w = __ldg(&tw[H + mm]);
        return mul_shoup_lazy3(x[0] + q3 - x[1], w.x, w.y, nq);
    } else {
        u64 A, B;
        if (c & 1) {
            const ulonglong2 wa = __ldg(&tw[H + 2 * mm]), wb = __ldg(&tw[H + 2 * mm + 1]);
            A = mul_shoup_lazy3(x[0] + q3 - x[1], wa.x, wa.y, nq);
            B = mul_shoup_lazy3(x[2] + q3 - x[3], wb.x, wb.y, nq);
        } else {
            A = csub(x[0] + x[1], q3);
            B = csub(x[2] + x[3], q3);
        }
        if (!(c & 2)) return csub(A + B, q3);
        const ulonglong2 w = __ldg(&tw[(H >> 1) + mm]);
        return mul_shoup_lazy3(A + q3 - B, w.x, w.y, nq);
    }
}

template <int LOGN, int LC, bool FWD, int DIV>
__global__ void __cluster_dims__(1 << LC, 1, 1) __launch_bounds__((1 << (LOGN - LC)) / DIV) ntt_split_kernel(const u64 *in, u64 *out, NttArgs a, NttSrc src)
{
    pdl_enter();
    constexpr int LOGM = LOGN - LC, M = 1 << LOGM, C = 1 << LC;
    constexpr int RF = (LOGM % 3 == 0) ? 3 : (LOGM % 3 == 1 ? 4 : 2);
    constexpr int MID = (LOGM - RF - 3) / 3;
    static_assert(LOGM >= RF + 3, "slice too small");
    extern __shared__ u64 sm[];
    const unsigned p = blockIdx.x >> LC, c = blockIdx.x & (C - 1);
    const int slot = p % a.pattern_len;
    const DMod m = a.mod[slot];
    const u64 q = m.q, nq = 0 - m.q, q3 = 3 * m.q;
    const ulonglong2 *tw = a.tw + ((size_t)a.table[slot] * 2 + (FWD ? 0 : 1)) * (size_t)(1 << LOGN);
    const u64 *ip = in + (size_t)(src.src_idx ? src.src_idx[p] : p) * (1 << LOGN);
    u64 *op = out + (size_t)(src.dst_idx ? src.dst_idx[p] : p) * (1 << LOGN);
    const bool reduce = src.reduce_input != 0;

    if (FWD) {
        // head stages + stages [0, RF) of the slice: global -> registers -> shared
        constexpr unsigned stride = M >> RF;
        for (unsigned g = threadIdx.x; g < stride; g += blockDim.x) {
            u64 raw[1 << RF][C];
#pragma unroll
            for (int k = 0; k < (1 << RF); k++)
#pragma unroll
                for (int i = 0; i < C; i++) raw[k][i] = ip[g + k * stride + i * M];
            u64 x[1 << RF];
#pragma unroll
            for (int k = 0; k < (1 << RF); k++) {
                if (reduce) {
#pragma unroll
                    for (int i = 0; i < C; i++) raw[k][i] = barrett64(raw[k][i], m);
                }
                x[k] = fwd_split_head<LC>(raw[k], tw, c, nq, q3);
            }
            fwd_group<RF>(x, tw, 0, 0, nq, q3, C + c);
#pragma unroll
            for (int k = 0; k < (1 << RF); k++) sm[pad_idx(g + k * stride)] = x[k];
        }
        __syncthreads();
        cluster_arrive();
        MidRunner<LOGM, true, RF, MID>::run(sm, tw, nq, q3, C + c);
        cluster_wait();
        u64 *oc = op + (size_t)c * M;
        for (unsigned g = threadIdx.x; g < (M >> 3); g += blockDim.x) {
            u64 x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = sm[pad_idx(8 * g + k)];
            fwd_group<3>(x, tw, g, LOGM - 3, nq, q3, C + c);
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = csub(csub(csub(x[k], q3), 2 * q), q);
            ulonglong2 *o2 = reinterpret_cast<ulonglong2 *>(oc + 8 * g);
#pragma unroll
            for (int k = 0; k < 4; k++) o2[k] = make_ulonglong2(x[2 * k], x[2 * k + 1]);
        }
    } else {
        // head stages + stages [LOGM-3, LOGM) of the strided slice: 8*C consecutive inputs per thread
        for (unsigned g = threadIdx.x; g < (M >> 3); g += blockDim.x) {
            u64 raw[8][C];
            const ulonglong2 *i2 = reinterpret_cast<const ulonglong2 *>(ip + (size_t)8 * C * g);
#pragma unroll
            for (int k = 0; k < 8; k++)
#pragma unroll
                for (int i = 0; i < C; i += 2) {
                    const ulonglong2 v = i2[(k * C + i) >> 1];
                    raw[k][i] = v.x;
                    raw[k][i + 1] = v.y;
                }
            u64 x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (reduce) {
#pragma unroll
                    for (int i = 0; i < C; i++) raw[k][i] = barrett64(raw[k][i], m);
                }
                x[k] = inv_split_head<LC, LOGN>(raw[k], tw, 8 * g + k, c, nq, q3);
            }
            inv_group<3>(x, tw, g, LOGM - 3, nq, q3);
#pragma unroll
            for (int k = 0; k < 8; k++) sm[pad_idx(8 * g + k)] = x[k];
        }
        __syncthreads();
        cluster_arrive();
        MidRunner<LOGM, false, RF, MID>::run(sm, tw, nq, q3);
        cluster_wait();
        const DShoup inv_n = a.inv_n[slot], inv_n_w = a.inv_n_w[slot];
        constexpr unsigned stride = M >> RF;
        for (unsigned g = threadIdx.x; g < stride; g += blockDim.x) {
            u64 x[1 << RF];
#pragma unroll
            for (int k = 0; k < (1 << RF); k++) x[k] = sm[pad_idx(g + k * stride)];
            inv_group_upper<RF>(x, tw, nq, q3);
            constexpr int d = 1 << (RF - 1);
#pragma unroll
            for (int k = 0; k < d; k++) {
                const u64 u = x[k], v = x[k + d];
                const u64 lo = mul_shoup_lazy3(u + v, inv_n.op, inv_n.quot, nq);
                const u64 hi = mul_shoup_lazy3(u + q3 - v, inv_n_w.op, inv_n_w.quot, nq);
                op[(size_t)(g + k * stride) * C + c] = csub(csub(lo, 2 * q), q);
                op[(size_t)(g + (k + d) * stride) * C + c] = csub(csub(hi, 2 * q), q);
            }
        }
    }
}

} // namespace apsu_b200
