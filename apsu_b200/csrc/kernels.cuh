// Element-wise RNS kernels of the path (K4-K9 of SURVEY.md §2.2): BEHZ base extension, tensor
// product, scale-and-round back to base q, key-switch inner product and mod-down, mod-switch,
// plaintext addition, final bit clearing.  One thread per coefficient; all polynomials live in one
// device arena and are addressed by 32-bit polynomial indices (units of N words), so that a whole
// batch of independent ciphertext operations is one launch.
//
// Every kernel returns canonical residues and follows SEAL 3.7's formulas —
// intermediate laziness is free, the reduced outputs are what must match bit-for-bit.
#pragma once
#include "device_ctx.hpp"

namespace apsu_b200 {



// Base conversions multiply residues by CONSTANTS, so every product is a lazy Shoup product (mul_shoup_lazy3: 4 IMAD.WIDE
// + 1 IMAD.HI, result in [0, 3m)) and a sum of them is reduced once.  Sums are kept below 8m < 2^64 (m < 2^61): a
// canonical value plus at most two lazy terms between reductions.
struct LazySum {
    u64 s = 0;
    int pending = 0;
    __device__ __forceinline__ void add(u64 x, const DShoup &w, u64 m)
    {
        s += mul_shoup_lazy3(x, w.op, w.quot, 0 - m);
        if (++pending == 2) {
            s = reduce_8q(s, m);
            pending = 0;
        }
    }
    __device__ __forceinline__ u64 get(u64 m) const { return pending ? reduce_8q(s, m) : s; }
};

// ---- BEHZ steps (1)-(2): base q -> base Bsk, Montgomery-reduced (fastbconv_m_tilde + sm_mrq) ----
// grid (N/256, n_polys).  src[rp] -> first prime of an RNS polynomial [L][N]; dst[rp] -> [S][N].
// x'_j = (y_j + q*r) * m_tilde^-1 with y = FastBConv(x*m_tilde; q -> Bsk): the m_tilde^-1 is folded into the
// conversion constants (ext_punct_bsk, ext_q_bsk), everything is arithmetic modulo Bsk_j and the output canonical.
// TL, TS: compile-time |q| and |Bsk| (0: read them from the constants) — the specialised instances are fully
// unrolled with no predicated-off work; the engine dispatches on (L, S).
template <int TL, int TS>
__global__ void __launch_bounds__(kEwThreads)
k_behz_extend(u64 *A, const u32 *__restrict__ src, const u32 *__restrict__ dst, LevelConsts c, int N)
{
    pdl_enter();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 rp = blockIdx.y;
    const u64 *x = A + (size_t)src[rp] * N + n;
    u64 *o = A + (size_t)dst[rp] * N + n;
    const int L = TL ? TL : c.L, S = TS ? TS : c.S;
    constexpr int ML = TL ? TL : kMaxQ, MS = TS ? TS : kMaxBsk;
    u64 tmp[ML];
    u32 ymt = 0;
#pragma unroll
    for (int i = 0; i < ML; i++) {
        if (i < L) {
            tmp[i] = mul_shoup(x[(size_t)i * N], c.mtilde_inv_punct_q[i], c.q[i].q);
            ymt += (u32)tmp[i] * c.q_punct_mod_mtilde[i];
        }
    }
    const u32 r = ymt * c.neg_inv_q_mod_mtilde; // arithmetic mod m_tilde = 2^32
#pragma unroll
    for (int j = 0; j < MS; j++) {
        if (j >= S) break;
        const u64 m = c.bsk[j].q;
        LazySum acc;
#pragma unroll
        for (int i = 0; i < ML; i++)
            if (i < L) acc.add(tmp[i], c.ext_punct_bsk[j][i], m);
        u64 rr = r;
        if (r >= 0x80000000u) rr += m - 0x100000000ull; // centred lift of r
        acc.add(rr, c.ext_q_bsk[j], m);
        o[(size_t)j * N] = acc.get(m);
    }
}

// ---- BEHZ step (4): size-2 x size-2 tensor product in NTT form over the extended base q ∪ Bsk ----
// grid (N/256, L+S, n_ops).  Extended ciphertext = [2][L+S][N]; output [3][L+S][N].
__global__ void __launch_bounds__(kEwThreads)
k_tensor(u64 *A, const u32 *__restrict__ a_idx, const u32 *__restrict__ b_idx, const u32 *__restrict__ d_idx, LevelConsts c, int N)
{
    pdl_enter();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, LS = c.L + c.S;
    const u32 o = blockIdx.z;
    const DMod m = j < c.L ? c.q[j] : c.bsk[j - c.L];
    const u64 *a = A + ((size_t)a_idx[o] + j) * N + n;
    const u64 *b = A + ((size_t)b_idx[o] + j) * N + n;
    u64 *d = A + ((size_t)d_idx[o] + j) * N + n;
    const size_t cs = (size_t)LS * N; // component stride
    u64 a0 = a[0], a1 = a[cs], b0 = b[0], b1 = b[cs];
    // three full products instead of four: a0*b1 + a1*b0 = (a0+a1)*(b0+b1) - a0*b0 - a1*b1 on the unreduced 128-bit
    // products (operands < 2^61, so the sums fit a word and nothing wraps); each of the three values is below
    // 2^(64+sh) and is reduced with the one-word Barrett
    const u64 p0l = a0 * b0, p0h = mulhi(a0, b0), p2l = a1 * b1, p2h = mulhi(a1, b1);
    const u64 sa = a0 + a1, sb = b0 + b1;
    u64 ml = sa * sb, mh = mulhi(sa, sb);
    asm("sub.cc.u64 %0, %0, %2;\n\tsubc.u64 %1, %1, %3;" : "+l"(ml), "+l"(mh) : "l"(p0l), "l"(p0h));
    asm("sub.cc.u64 %0, %0, %2;\n\tsubc.u64 %1, %1, %3;" : "+l"(ml), "+l"(mh) : "l"(p2l), "l"(p2h));
    d[0] = barrett_prod(p0l, p0h, m);
    d[cs] = barrett_prod(ml, mh, m);
    d[2 * cs] = barrett_prod(p2l, p2h, m);
}

// ---- BEHZ steps (6)-(8): multiply by t, fast floor (divide by q), Shenoy-Kumaresan back to q ----
// grid (N/256, n_polys).  src[rp] -> [L+S][N] coefficient form; dst[rp] -> [L][N].
template <int TL, int TS>
__global__ void __launch_bounds__(kEwThreads)
k_behz_scale_down(u64 *A, const u32 *__restrict__ src, const u32 *__restrict__ dst, LevelConsts c, int N)
{
    pdl_enter();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 rp = blockIdx.y;
    const u64 *d = A + (size_t)src[rp] * N + n;
    u64 *o = A + (size_t)dst[rp] * N + n;
    const int L = TL ? TL : c.L, S = TS ? TS : c.S, nb = S - 1;
    constexpr int kMaxQ = TL ? TL : apsu_b200::kMaxQ, kMaxBsk = TS ? TS : apsu_b200::kMaxBsk; // loop bounds of this instance
    u64 tmp[kMaxQ], f[kMaxBsk];
#pragma unroll
    for (int i = 0; i < kMaxQ; i++)
        if (i < L) tmp[i] = mul_shoup(d[(size_t)i * N], c.t_inv_punct_q[i], c.q[i].q);
    // f_j = (t*d_j - FastBConv(t*d_q; q -> Bsk_j)) * q^-1, the q^-1 folded into both constants; for the primes of B the
    // constants also carry (B/B_j)^-1, so f[j] (j < nb) already is the Shenoy-Kumaresan digit g_j = f_j * (B/B_j)^-1
#pragma unroll
    for (int j = 0; j < kMaxBsk; j++) {
        if (j < S) {
            const u64 m = c.bsk[j].q;
            LazySum acc;
            acc.add(d[(size_t)(L + j) * N], c.floor_t_bsk[j], m);
#pragma unroll
            for (int i = 0; i < kMaxQ; i++)
                if (i < L) acc.add(tmp[i], c.floor_punct_bsk[j][i], m);
            f[j] = acc.get(m);
        }
    }
    // Shenoy-Kumaresan
    const u64(&g)[kMaxBsk] = f;
    const u64 msk = c.bsk[S - 1].q;
    LazySum aacc;
#pragma unroll
    for (int k = 0; k < kMaxBsk; k++) {
        if (k < nb) aacc.add(g[k], c.B_punct_mod_msk[k], msk);
    }
    const u64 alpha = sub_mod(aacc.get(msk), f[S - 1], msk); // B^-1 is inside B_punct_mod_msk and the m_sk floor constants
    const bool neg = alpha > (msk >> 1);
    const u64 corr = neg ? msk - alpha : alpha;
#pragma unroll
    for (int i = 0; i < kMaxQ; i++) {
        if (i >= L) break;
        const u64 m = c.q[i].q;
        LazySum acc;
#pragma unroll
        for (int k = 0; k < kMaxBsk; k++)
            if (k < nb) acc.add(g[k], c.B_punct_mod_q[i][k], m);
        acc.add(corr, neg ? c.B_mod_q[i] : c.neg_B_mod_q[i], m);
        o[(size_t)i * N] = acc.get(m);
    }
}

// ---- key switching: inner product of the NTT'd digits with the relinearisation keys ----
// grid (N/256, R, n_ops).  digits[o] -> [L][R][N] (digit J, modulus slot I); keys [K-1][2][K][N];
// out[o] -> [2][R][N].
__global__ void __launch_bounds__(kEwThreads)
k_ks_mac(u64 *A, const u32 *__restrict__ dig_idx, const u32 *__restrict__ out_idx, const u64 *__restrict__ keys, KeySwitchConsts c, int N)
{
    pdl_enter();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int I = blockIdx.y, R = c.L + 1;
    const u32 o = blockIdx.z;
    const int key_index = I == c.L ? c.K - 1 : I;
    const u64 *dg = A + ((size_t)dig_idx[o] + I) * N + n;
    // both components in one thread (they share the digits: one read of them instead of two, half the CTAs), and all
    // 3L loads first, then the products: ncu (profiles/ncu_r02_keyswitch_summary.json) showed the rolled loop waiting on
    // one load pair at a time (long-scoreboard 6.6 of 9.6 stall cycles per issue, FMA pipe 26 % busy)
    u64 dv[kMaxQ], k0[kMaxQ], k1[kMaxQ];
#pragma unroll
    for (int J = 0; J < kMaxQ; J++)
        if (J < c.L) {
            dv[J] = dg[(size_t)J * R * N];
            k0[J] = keys[(((size_t)J * 2 + 0) * c.K + key_index) * N + n];
            k1[J] = keys[(((size_t)J * 2 + 1) * c.K + key_index) * N + n];
        }
    Acc128 a0{ 0, 0 }, a1{ 0, 0 };
#pragma unroll
    for (int J = 0; J < kMaxQ; J++)
        if (J < c.L) {
            mac128(a0, dv[J], k0[J]);
            mac128(a1, dv[J], k1[J]);
        }
    u64 *out = A + ((size_t)out_idx[o] + I) * N + n;
    out[0] = barrett_prod(a0.lo, a0.hi, c.key_mod[I]); // <= 5 products of reduced operands
    out[(size_t)R * N] = barrett_prod(a1.lo, a1.hi, c.key_mod[I]);
}

// ---- key switching: divide by the special prime with rounding and add to (c0, c1) ----
// grid (N/256, 2, n_ops).  acc[o] -> [2][R][N] coefficient form; ct[o] -> [>=2][L][N] (c0,c1 read);
// dst[o] -> [2][L][N].
// PEERS: the PowersDag of a bundle index is split over several GPUs (collective C2) and this launch
// produces this rank's share of a DAG level: every output word is ALSO stored into the same place of the peers'
// arenas through NVLink peer memory (plain coalesced 8-byte stores, 256 B per warp), so the exchange of the level
// overlaps the arithmetic that produces it and no separate collective runs; k_xgpu_barrier closes the level.
constexpr int kMaxPeers = 7;
struct PeerArenas {
    u64 *base[kMaxPeers];
    int n;
};
template <bool PEERS>
__global__ void __launch_bounds__(kEwThreads)
k_ks_moddown(u64 *A, const u32 *__restrict__ acc_idx, const u32 *__restrict__ ct_idx, const u32 *__restrict__ dst_idx, KeySwitchConsts c, int N, PeerArenas peers)
{
    pdl_enter();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 comp = blockIdx.y, o = blockIdx.z;
    const int L = c.L, R = L + 1;
    const u64 *ac = A + ((size_t)acc_idx[o] + (size_t)comp * R) * N + n;
    const u64 *ct = A + ((size_t)ct_idx[o] + (size_t)comp * L) * N + n;
    const size_t dst_off = ((size_t)dst_idx[o] + (size_t)comp * L) * N + n;
    const u64 P = c.key_mod[L].q;
    u64 u = add_mod(ac[(size_t)L * N], c.half_P, P);
    for (int i = 0; i < L; i++) {
        const DMod m = c.key_mod[i];
        u64 delta = sub_mod(reduce_known(u, m, c.P_kind[i]), c.half_P_mod[i], m.q);
        u64 v = mul_shoup(sub_mod(ac[(size_t)i * N], delta, m.q), c.inv_P[i], m.q);
        const u64 r = add_mod(ct[(size_t)i * N], v, m.q);
        A[dst_off + (size_t)i * N] = r;
        if (PEERS) {
#pragma unroll
            for (int k = 0; k < kMaxPeers; k++)
                if (k < peers.n) peers.base[k][dst_off + (size_t)i * N] = r;
        }
    }
}

// Barrier between the GPUs that split a PowersDag, on the stream: every rank bumps its own epoch, publishes it into its
// slot of every peer's flag array (release, system scope: the peer-memory stores of the preceding kernels are ordered
// before it) and waits until every peer has published at least the same epoch (acquire).  One warp.  A peer that never
// arrives (it failed) must not hang this GPU: the wait gives up after `timeout_cycles` and raises flags.error.
struct PeerFlags {
    u32 *peer[kMaxPeers + 1]; // flag array of rank k of the group (entry `me` unused)
    u32 *mine;                // this rank's flag array [kMaxPeers + 1]
    u32 *epoch;               // this rank's barrier counter
    int *error;               // set when a wait timed out
    int n, me;
};
__global__ void k_xgpu_barrier(PeerFlags f, long long timeout_cycles)
{
    __shared__ u32 e_s;
    const int l = threadIdx.x;
    if (l == 0) {
        e_s = *f.epoch + 1;
        *f.epoch = e_s;
    }
    __syncthreads();
    const u32 e = e_s;
    __threadfence_system();
    if (l < f.n && l != f.me) {
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.peer[l] + f.me), "r"(e) : "memory");
        const long long t0 = clock64();
        u32 v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.mine + l) : "memory");
            if ((int)(v - e) >= 0) break;
            if (clock64() - t0 > timeout_cycles) {
                atomicExch(f.error, 1);
                break;
            }
        }
    }
    __syncthreads();
    __threadfence_system();
}

// ---- mod_switch_to_next (divide_and_round_q_last): [L][N] -> [L-1][N] ----
// grid (N/256, n_polys)
__global__ void __launch_bounds__(kEwThreads)
k_mod_switch_next(u64 *A, const u32 *__restrict__ src, const u32 *__restrict__ dst, LevelConsts c, int N)
{
    pdl_enter();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 rp = blockIdx.y;
    const u64 *x = A + (size_t)src[rp] * N + n;
    u64 *o = A + (size_t)dst[rp] * N + n;
    const int L = c.L;
    const u64 ql = c.q[L - 1].q;
    const u64 a = add_mod(x[(size_t)(L - 1) * N], ql >> 1, ql);
    for (int i = 0; i + 1 < L; i++) {
        const DMod m = c.q[i];
        u64 tmp = sub_mod(reduce_known(a, m, c.last_kind[i]), c.half_mod[i], m.q);
        o[(size_t)i * N] = mul_shoup(sub_mod(x[(size_t)i * N], tmp, m.q), c.inv_qlast[i], m.q);
    }
}

// ---- sum of `count` RNS polynomials: dst = sum_k A[src[rp*count_stride + k]] (mod q_j) ----
// grid (N/256, L, n_out).  lists: first[rp], n_terms[rp] into `terms`.
__global__ void __launch_bounds__(kEwThreads)
k_sum_polys(u64 *A, const u32 *__restrict__ terms, const u32 *__restrict__ first, const u32 *__restrict__ n_terms,
            const u32 *__restrict__ dst, LevelConsts c, int N)
{
    pdl_enter();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const u32 rp = blockIdx.z;
    const u64 q = c.q[j].q;
    u64 s = 0;
    const u32 f = first[rp], cnt = n_terms[rp];
    for (u32 k = 0; k < cnt; k++) s = add_mod(s, A[((size_t)terms[f + k] + j) * N + n], q);
    A[((size_t)dst[rp] + j) * N + n] = s;
}

// ---- Query validation: seal::is_valid_for -> is_data_valid_for (receiver/apsu/query.cpp:45-66) ----
// every residue of polynomial p must be below modulus p % nmods.  grid (N/256, n_polys); *bad is set on a violation.
struct RangeMods {
    u64 q[8];
    int n;
};
__global__ void __launch_bounds__(kEwThreads) k_check_range(const u64 *__restrict__ base, RangeMods mods, int N, int *__restrict__ bad)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 p = blockIdx.y;
    if (base[(size_t)p * N + n] >= mods.q[p % mods.n]) atomicExch(bad, 1);
}

} // namespace apsu_b200
