// Device-side view of the SEAL context the path needs (CryptoContext / SEALContext / RNSTool):
// modulus table, NTT twiddles, per-level BEHZ and scaling constants.  Built once per context by
// DeviceContext::build (context.cu).  Reference: common/apsu/crypto_context.h:28-125; the constants
// mirror SEAL 3.7's ContextData/RNSTool.
#pragma once
#include "modarith.cuh"
#include <cstdint>
#include <vector>

namespace apsu_b200 {

constexpr int kMaxQ = 5;   // data-level primes (K-1 <= 5 supported; reference parameter files use <= 4)
constexpr int kMaxBsk = 7; // |B| + 1 (|B| = L or L+1)
constexpr int kMaxKey = 6; // K
constexpr int kMaxPattern = 12; // |q| + |Bsk| of the largest level
constexpr int kEwThreads = 256; // element-wise kernels: one thread per coefficient

// NTT launch descriptor: polynomial p uses modulus slot p % pattern_len
struct NttArgs {
    const ulonglong2 *tw; // [modulus][fwd|inv][N] Shoup pairs, bit-reversed order
    DMod mod[kMaxPattern];
    DShoup inv_n[kMaxPattern];
    DShoup inv_n_w[kMaxPattern]; // N^-1 times the twiddle of the last inverse stage (psi^-(N/2)): the scaling rides on that stage
    int table[kMaxPattern]; // modulus-table index (for the twiddle offset)
    int pattern_len;
};

// per-level constants (level = number of data primes L)
struct LevelConsts {
    int L, S; // |q|, |Bsk| (= |B|+1); B = Bsk[0..S-2], m_sk = Bsk[S-1]
    DMod q[kMaxQ];
    DMod bsk[kMaxBsk];
    // plaintext scaling (add_plain): floor(q/t) mod q_j, q mod t
    u64 coeff_div_plain[kMaxQ];
    u64 q_mod_t;
    // mod_switch_to_next: (q_{L-1})^-1 mod q_j (Shoup), (q_{L-1}>>1) mod q_j
    DShoup inv_qlast[kMaxQ];
    u64 half_mod[kMaxQ];
    u32 last_kind[kMaxQ];               // reduce_known kind of a residue of q_{L-1} modulo q_j
    // BEHZ step 1: x_i * (m_tilde * (q/q_i)^-1) mod q_i
    DShoup mtilde_inv_punct_q[kMaxQ];
    // steps 1-2 fused: x'_j = sum_i tmp_i * ext_punct_bsk[j][i] + r * ext_q_bsk[j]  (mod Bsk_j) with
    // ext_punct_bsk = (q/q_i) * m_tilde^-1, ext_q_bsk = q * m_tilde^-1; (q/q_i) mod 2^32 for the m_tilde residue
    DShoup ext_punct_bsk[kMaxBsk][kMaxQ];
    DShoup ext_q_bsk[kMaxBsk];
    u32 q_punct_mod_mtilde[kMaxQ];
    u32 neg_inv_q_mod_mtilde;
    // steps 6-8
    DShoup t_mod_q[kMaxQ];              // t mod q_i   (multiply by plain modulus)
    DShoup inv_punct_q[kMaxQ];          // (q/q_i)^-1 mod q_i
    DShoup t_inv_punct_q[kMaxQ];        // t * (q/q_i)^-1 mod q_i
    // step 7 fused: f_j = d_j * floor_t_bsk[j] + sum_i tmp_i * floor_punct_bsk[j][i]  (mod Bsk_j) with
    // floor_t_bsk = t * q^-1, floor_punct_bsk = -(q/q_i) * q^-1 — both times (B/B_j)^-1 for the primes of B (j < |B|),
    // which makes f_j the Shenoy-Kumaresan digit of step 8, and times B^-1 for m_sk (inv_punct_B / inv_B_mod_msk: reference only)
    DShoup floor_t_bsk[kMaxBsk];
    DShoup floor_punct_bsk[kMaxBsk][kMaxQ];
    DShoup inv_punct_B[kMaxBsk];           // (B/B_i)^-1 mod B_i
    DShoup B_punct_mod_q[kMaxQ][kMaxBsk];  // (B/B_i) mod q_j
    DShoup B_punct_mod_msk[kMaxBsk];       // (B/B_i) * B^-1 mod m_sk (the floor constants of m_sk carry the same B^-1)
    DShoup inv_B_mod_msk;               // B^-1 mod m_sk
    DShoup B_mod_q[kMaxQ];              // B mod q_j
    DShoup neg_B_mod_q[kMaxQ];          // -B mod q_j
};

// key-switching constants at level L (decomposition over q_0..q_{L-1}, special prime P = q_{K-1})
struct KeySwitchConsts {
    int L, K;
    DMod key_mod[kMaxKey];  // slot I<L -> q_I ; slot L -> P
    DShoup inv_P[kMaxQ];    // P^-1 mod q_i
    u64 half_P;             // P >> 1
    u64 half_P_mod[kMaxQ];  // (P>>1) mod q_i
    u32 P_kind[kMaxQ];      // reduce_known kind of a residue of P modulo q_i
};

// fused prologues of the transforms (ntt.cuh)
enum { kNttPlain = 0, kNttExtend = 1, kNttTensor = 2, kNttKsMac = 3 };
struct NttFuse {
    const LevelConsts *lc; // kNttExtend: constants of the level (device copy)
    const unsigned *a_idx; // kNttTensor: operand a per op;  kNttKsMac: digit block per op ([L][R] polynomials)
    const unsigned *b_idx; // kNttTensor: operand b per op
    const u64 *keys;       // kNttKsMac: relinearisation keys [K-1][2][K][N]
    int L, S, K;           // |q|, |Bsk| (kNttExtend / kNttTensor); |q|, -, key primes (kNttKsMac)
};


} // namespace apsu_b200
