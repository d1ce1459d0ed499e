// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the Microsoft SEAL 3.7 algorithms that APSU's receiver-side query
// evaluation calls.  SEAL is a third-party dependency of the reference
// (cmake/APSUConfig.cmake.in:44 pins "3.7") and is NOT vendored under /root/reference nor
// installed in this image, so the algorithms are restated from SEAL's published design
// (BFV with BEHZ RNS multiplication, Harvey NTT, RNS key switching with one special prime).
//
// PARITY UNPINNED: the reference tree holds no ciphertext-level golden vectors
// (SURVEY.md §4, §8c).  The oracle is pinned only by (i) the derived-constant fixtures of
// SURVEY.md Appendix B/D, (ii) an independent Python big-integer model (tests/bigint_model.py)
// and (iii) algebraic properties (decrypt(eval) - mask == P(x)).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// use anything in this directory.
//
// Every function returns canonical residues in [0, m); SEAL's internal lazy ranges are not
// observable at the API boundary, so the mathematics is what is restated.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>
#include <array>
#include <algorithm>
#include <random>

namespace orc {

using u64 = std::uint64_t;
using u128 = unsigned __int128;

// ----------------------------------------------------------------------------------------------
// Modulus + Barrett arithmetic (SEAL: seal/modulus.h, seal/util/uintarithsmallmod.h)
// ----------------------------------------------------------------------------------------------
struct Modulus {
    u64 value = 0;
    u64 ratio[2] = { 0, 0 }; // floor(2^128 / value), low and high word
    int bits = 0;
    Modulus() = default;
    explicit Modulus(u64 v) { set(v); }
    void set(u64 v)
    {
        value = v;
        bits = 0;
        for (u64 x = v; x; x >>= 1) bits++;
        // floor(2^128 / v) via two-step long division
        u128 num_hi = (u128)1 << 64; // 2^128 = num_hi * 2^64
        u64 q_hi = (u64)(num_hi / v);
        u128 rem = num_hi % v;
        u64 q_lo = (u64)((rem << 64) / v);
        ratio[0] = q_lo;
        ratio[1] = q_hi;
    }
};

// SEAL barrett_reduce_128: x (128-bit) mod m
static inline u64 barrett_reduce_128(u128 x, const Modulus &m)
{
    u64 x0 = (u64)x, x1 = (u64)(x >> 64);
    u64 carry = (u64)(((u128)x0 * m.ratio[0]) >> 64);
    u128 t = (u128)x0 * m.ratio[1] + carry;
    u64 tmp1 = (u64)t, tmp3 = (u64)(t >> 64);
    u128 t2 = (u128)x1 * m.ratio[0] + tmp1;
    carry = (u64)(t2 >> 64);
    u64 qhat = x1 * m.ratio[1] + tmp3 + carry;
    u64 r = x0 - qhat * m.value;
    while (r >= m.value) r -= m.value;
    return r;
}
static inline u64 barrett_reduce_64(u64 x, const Modulus &m)
{
    u64 qhat = (u64)(((u128)x * m.ratio[1]) >> 64);
    u64 r = x - qhat * m.value;
    while (r >= m.value) r -= m.value;
    return r;
}
static inline u64 mul_mod(u64 a, u64 b, const Modulus &m) { return barrett_reduce_128((u128)a * b, m); }
static inline u64 add_mod(u64 a, u64 b, const Modulus &m)
{
    u64 s = a + b;
    return s >= m.value ? s - m.value : s;
}
static inline u64 sub_mod(u64 a, u64 b, const Modulus &m) { return a >= b ? a - b : a + m.value - b; }
static inline u64 neg_mod(u64 a, const Modulus &m) { return a ? m.value - a : 0; }
static inline u64 pow_mod(u64 a, u64 e, const Modulus &m)
{
    u64 r = 1 % m.value;
    a = barrett_reduce_64(a, m);
    while (e) {
        if (e & 1) r = mul_mod(r, a, m);
        a = mul_mod(a, a, m);
        e >>= 1;
    }
    return r;
}
// inverse by extended Euclid (modulus may be composite, e.g. m_tilde = 2^32)
static inline bool try_inv_mod(u64 a, u64 m, u64 &out)
{
    __int128 r0 = m, r1 = a % m, s0 = 0, s1 = 1;
    while (r1) {
        __int128 q = r0 / r1;
        __int128 t = r0 - q * r1;
        r0 = r1;
        r1 = t;
        t = s0 - q * s1;
        s0 = s1;
        s1 = t;
    }
    if (r0 != 1) return false;
    if (s0 < 0) s0 += m;
    out = (u64)s0;
    return true;
}
static inline u64 inv_mod(u64 a, const Modulus &m)
{
    u64 r;
    if (!try_inv_mod(a, m.value, r)) throw std::logic_error("value not invertible");
    return r;
}

// Shoup-form constant operand (SEAL MultiplyUIntModOperand)
struct ShoupOp {
    u64 op = 0, quot = 0;
    void set(u64 v, const Modulus &m)
    {
        op = v;
        quot = (u64)((((u128)v) << 64) / m.value);
    }
};
static inline u64 mul_shoup(u64 x, const ShoupOp &y, const Modulus &m)
{
    u64 hi = (u64)(((u128)x * y.quot) >> 64);
    u64 r = y.op * x - hi * m.value;
    return r >= m.value ? r - m.value : r;
}

// ----------------------------------------------------------------------------------------------
// Primes (SEAL: util/numth.cpp get_primes / is_prime; modulus.cpp CoeffModulus::Create,
// PlainModulus::Batching) — SURVEY.md A.1
// ----------------------------------------------------------------------------------------------
static inline bool is_prime(u64 n)
{
    if (n < 2) return false;
    static const u64 small[] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37 };
    for (u64 p : small) {
        if (n == p) return true;
        if (n % p == 0) return false;
    }
    u64 d = n - 1;
    int r = 0;
    while (!(d & 1)) {
        d >>= 1;
        r++;
    }
    Modulus m(n);
    for (u64 a : small) { // deterministic for all 64-bit n
        u64 x = pow_mod(a, d, m);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < r; i++) {
            x = mul_mod(x, x, m);
            if (x == n - 1) {
                comp = false;
                break;
            }
        }
        if (comp) return false;
    }
    return true;
}

// primes = 1 (mod factor), `bits` bits, descending from the top of the range
static inline std::vector<u64> get_primes(u64 factor, int bits, size_t count)
{
    std::vector<u64> out;
    u64 value = (((u64)1 << bits) - 1) / factor * factor + 1;
    u64 lower = (u64)1 << (bits - 1);
    while (count > 0 && value > lower) {
        if (is_prime(value)) {
            out.push_back(value);
            count--;
        }
        value -= factor;
    }
    if (count > 0) throw std::logic_error("failed to find enough qualifying primes");
    return out;
}

// CoeffModulus::Create(N, bit_sizes): one descending list per distinct size, consumed from the back
static inline std::vector<u64> coeff_modulus_create(size_t N, const std::vector<int> &bit_sizes)
{
    std::vector<u64> result;
    std::vector<std::pair<int, std::vector<u64>>> tables;
    for (int b : bit_sizes) {
        bool seen = false;
        for (auto &t : tables) seen |= (t.first == b);
        if (seen) continue;
        size_t cnt = (size_t)std::count(bit_sizes.begin(), bit_sizes.end(), b);
        tables.emplace_back(b, get_primes(2 * (u64)N, b, cnt));
    }
    for (int b : bit_sizes) {
        for (auto &t : tables) {
            if (t.first == b) {
                result.push_back(t.second.back());
                t.second.pop_back();
            }
        }
    }
    return result;
}
static inline u64 plain_modulus_batching(size_t N, int bits) { return get_primes(2 * (u64)N, bits, 1)[0]; }

// minimal primitive `degree`-th root of unity (degree a power of two) — SURVEY.md A.3
static inline u64 minimal_primitive_root(u64 degree, const Modulus &m)
{
    u64 group = m.value - 1;
    if (group % degree) throw std::logic_error("modulus is not 1 mod degree");
    u64 quot = group / degree;
    u64 root = 0;
    for (u64 x = 2;; x++) {
        u64 g = pow_mod(x, quot, m);
        if (pow_mod(g, degree >> 1, m) == m.value - 1) { // primitive: g^(degree/2) == -1
            root = g;
            break;
        }
    }
    u64 gsq = mul_mod(root, root, m), cur = root, best = root;
    for (u64 i = 0; i < degree / 2; i++) { // all odd powers
        if (cur < best) best = cur;
        cur = mul_mod(cur, gsq, m);
    }
    return best;
}

static inline int log2_exact(size_t n)
{
    int l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}
static inline u64 reverse_bits(u64 x, int bits)
{
    u64 r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

// ----------------------------------------------------------------------------------------------
// Negacyclic NTT (SEAL: util/ntt.cpp, util/dwthandler.h) — SURVEY.md A.3
//   forward : natural order in, bit-reversed order out, out[k] = a(psi^(2*bitrev(k)+1))
//   inverse : exact inverse incl. N^-1 scaling, fully reduced
// ----------------------------------------------------------------------------------------------
struct NTTTables {
    size_t N = 0;
    int logN = 0;
    Modulus mod;
    u64 root = 0;
    std::vector<ShoupOp> w;    // w[bitrev(i)]  = psi^i
    std::vector<ShoupOp> winv; // winv[bitrev(i)] = psi^-i
    ShoupOp inv_n;
    void init(size_t n, u64 modulus)
    {
        N = n;
        logN = log2_exact(n);
        mod.set(modulus);
        root = minimal_primitive_root(2 * (u64)n, mod);
        u64 iroot = inv_mod(root, mod);
        w.assign(n, ShoupOp());
        winv.assign(n, ShoupOp());
        u64 p = 1, ip = 1;
        for (size_t i = 0; i < n; i++) {
            w[reverse_bits(i, logN)].set(p, mod);
            winv[reverse_bits(i, logN)].set(ip, mod);
            p = mul_mod(p, root, mod);
            ip = mul_mod(ip, iroot, mod);
        }
        inv_n.set(inv_mod((u64)n % modulus, mod), mod);
    }
    void forward(u64 *a) const
    {
        const u64 q = mod.value;
        size_t t = N;
        for (size_t m = 1; m < N; m <<= 1) {
            t >>= 1;
            for (size_t i = 0; i < m; i++) {
                const ShoupOp &W = w[m + i];
                u64 *x = a + 2 * i * t, *y = x + t;
                for (size_t j = 0; j < t; j++) {
                    u64 u = x[j], v = mul_shoup(y[j], W, mod);
                    u64 s = u + v;
                    x[j] = s >= q ? s - q : s;
                    y[j] = u >= v ? u - v : u + q - v;
                }
            }
        }
    }
    void inverse(u64 *a) const
    {
        const u64 q = mod.value;
        size_t t = 1;
        for (size_t m = N; m > 1; m >>= 1) {
            size_t h = m >> 1;
            for (size_t i = 0; i < h; i++) {
                const ShoupOp &W = winv[h + i];
                u64 *x = a + 2 * i * t, *y = x + t;
                for (size_t j = 0; j < t; j++) {
                    u64 u = x[j], v = y[j];
                    u64 s = u + v;
                    x[j] = s >= q ? s - q : s;
                    u64 d = u >= v ? u - v : u + q - v;
                    y[j] = mul_shoup(d, W, mod);
                }
            }
            t <<= 1;
        }
        for (size_t j = 0; j < N; j++) a[j] = mul_shoup(a[j], inv_n, mod);
    }
};

// ----------------------------------------------------------------------------------------------
// RNS base + fast base conversion (SEAL: util/rns.cpp RNSBase / BaseConverter) — SURVEY.md A.6
// ----------------------------------------------------------------------------------------------
struct RNSBase {
    std::vector<Modulus> base;
    std::vector<u64> inv_punct; // (P/p_i)^-1 mod p_i
    size_t size() const { return base.size(); }
    void init(const std::vector<u64> &primes)
    {
        base.clear();
        for (u64 p : primes) base.emplace_back(p);
        inv_punct.resize(primes.size());
        for (size_t i = 0; i < primes.size(); i++) {
            u64 prod = 1;
            for (size_t k = 0; k < primes.size(); k++)
                if (k != i) prod = mul_mod(prod, barrett_reduce_64(primes[k], base[i]), base[i]);
            inv_punct[i] = inv_mod(prod, base[i]);
        }
    }
    // (P / p_i) mod m
    u64 punct_mod(size_t i, const Modulus &m) const
    {
        u64 prod = 1 % m.value;
        for (size_t k = 0; k < base.size(); k++)
            if (k != i) prod = mul_mod(prod, barrett_reduce_64(base[k].value, m), m);
        return prod;
    }
    u64 prod_mod(const Modulus &m) const
    {
        u64 prod = 1 % m.value;
        for (size_t k = 0; k < base.size(); k++) prod = mul_mod(prod, barrett_reduce_64(base[k].value, m), m);
        return prod;
    }
};

struct BaseConverter {
    const RNSBase *ibase = nullptr;
    std::vector<Modulus> obase;
    std::vector<std::vector<u64>> matrix; // matrix[j][i] = (P/p_i) mod obase_j
    void init(const RNSBase &in, const std::vector<Modulus> &out)
    {
        ibase = &in;
        obase = out;
        matrix.assign(out.size(), std::vector<u64>(in.size()));
        for (size_t j = 0; j < out.size(); j++)
            for (size_t i = 0; i < in.size(); i++) matrix[j][i] = in.punct_mod(i, out[j]);
    }
    // in: [ibase][N] ; out: [obase][N]
    void fast_convert_array(const u64 *in, u64 *out, size_t N) const
    {
        size_t is = ibase->size(), os = obase.size();
        std::vector<u64> tmp(is * N);
        for (size_t i = 0; i < is; i++)
            for (size_t n = 0; n < N; n++) tmp[i * N + n] = mul_mod(in[i * N + n], ibase->inv_punct[i], ibase->base[i]);
        for (size_t j = 0; j < os; j++)
            for (size_t n = 0; n < N; n++) {
                // SEAL dot_product_mod: 128-bit lazy sum (ibase sizes are tiny), single reduction
                u128 acc = 0;
                for (size_t i = 0; i < is; i++) {
                    u128 p = (u128)tmp[i * N + n] * matrix[j][i];
                    // keep the running sum reduced enough never to overflow: fold when high bit set
                    acc += p;
                    if (acc >> 127) acc = barrett_reduce_128(acc, obase[j]);
                }
                out[j * N + n] = barrett_reduce_128(acc, obase[j]);
            }
    }
};

// ----------------------------------------------------------------------------------------------
// Context: key level (all K primes) and the data levels (first L primes), each with the RNS-tool
// constants BFV multiply / mod-switch need.  (SEAL: context.cpp, util/rns.cpp RNSTool) — A.2, A.6
// ----------------------------------------------------------------------------------------------
constexpr int kInternalModBits = 61;

struct Level {
    size_t L = 0; // number of coeff-modulus primes at this level
    RNSBase q;
    std::vector<const NTTTables *> ntt; // borrowed, one per prime
    // plaintext scaling (add_plain) constants
    u64 q_mod_t = 0;
    std::vector<u64> coeff_div_plain; // floor(q/t) mod q_j
    std::vector<u64> upper_half_increment; // q_j - t   (fast plain lift)
    // mod_switch_to_next: inverse of the last prime modulo the others
    std::vector<u64> inv_q_last_mod_q;
    // BEHZ
    RNSBase B;
    std::vector<Modulus> Bsk; // B ∪ {m_sk}
    Modulus m_sk, m_tilde, gamma;
    std::vector<NTTTables> Bsk_ntt;
    BaseConverter q_to_Bsk, q_to_mtilde, B_to_q, B_to_msk;
    std::vector<u64> m_tilde_mod_q;       // m_tilde mod q_i
    std::vector<u64> prod_q_mod_Bsk;      // q mod Bsk_j
    std::vector<u64> inv_prod_q_mod_Bsk;  // q^-1 mod Bsk_j
    std::vector<u64> inv_m_tilde_mod_Bsk; // m_tilde^-1 mod Bsk_j
    u64 neg_inv_prod_q_mod_m_tilde = 0;
    u64 inv_prod_B_mod_m_sk = 0;
    std::vector<u64> prod_B_mod_q; // B mod q_j
};

// big unsigned as little-endian u64 words; only what floor(q/t) mod q_j and q mod t need
struct BigUInt {
    std::vector<u64> w;
    static BigUInt product(const std::vector<u64> &f)
    {
        BigUInt r;
        r.w = { 1 };
        for (u64 x : f) {
            u64 carry = 0;
            for (auto &d : r.w) {
                u128 p = (u128)d * x + carry;
                d = (u64)p;
                carry = (u64)(p >> 64);
            }
            if (carry) r.w.push_back(carry);
        }
        return r;
    }
    u64 mod_small(u64 m) const
    {
        u128 r = 0;
        for (size_t i = w.size(); i-- > 0;) r = ((r << 64) | w[i]) % m;
        return (u64)r;
    }
    BigUInt div_small(u64 m) const
    {
        BigUInt q;
        q.w.assign(w.size(), 0);
        u128 r = 0;
        for (size_t i = w.size(); i-- > 0;) {
            u128 cur = (r << 64) | w[i];
            q.w[i] = (u64)(cur / m);
            r = cur % m;
        }
        while (q.w.size() > 1 && q.w.back() == 0) q.w.pop_back();
        return q;
    }
    int bit_count() const
    {
        int b = 0;
        for (u64 x = w.back(); x; x >>= 1) b++;
        return (int)(64 * (w.size() - 1)) + b;
    }
};

struct Context {
    size_t N = 0;
    int logN = 0;
    Modulus t;
    std::vector<u64> primes; // key-level coefficient modulus, K primes
    size_t K = 0;
    std::vector<NTTTables> ntt; // per key-level prime
    NTTTables plain_ntt;        // modulus t (batching)
    std::vector<size_t> slot_map; // BatchEncoder matrix_reps_index_map
    // levels[L] for L = 1..first_L (index 0 unused); key_inv_P_mod_q for key switching
    std::vector<Level> levels;
    size_t first_L = 0;
    std::vector<u64> inv_P_mod_q; // (special prime)^-1 mod q_i, i < K-1
    bool using_keyswitching() const { return K > 1; }

    void init(size_t n, u64 plain_modulus, const std::vector<u64> &coeff_primes)
    {
        N = n;
        logN = log2_exact(n);
        t.set(plain_modulus);
        primes = coeff_primes;
        K = primes.size();
        ntt.resize(K);
        for (size_t i = 0; i < K; i++) ntt[i].init(N, primes[i]);
        plain_ntt.init(N, plain_modulus);
        // BatchEncoder index map — SURVEY.md A.4
        slot_map.resize(N);
        {
            size_t row = N >> 1, m = N << 1;
            u64 pos = 1;
            for (size_t i = 0; i < row; i++) {
                u64 i1 = (pos - 1) >> 1, i2 = (m - pos - 1) >> 1;
                slot_map[i] = (size_t)reverse_bits(i1, logN);
                slot_map[row | i] = (size_t)reverse_bits(i2, logN);
                pos = (pos * 3) & (m - 1);
            }
        }
        first_L = (K > 1) ? K - 1 : 1;
        levels.assign(first_L + 1, Level());
        for (size_t L = 1; L <= first_L; L++) init_level(levels[L], L);
        inv_P_mod_q.clear();
        if (K > 1) {
            for (size_t i = 0; i + 1 < K; i++) {
                Modulus qi(primes[i]);
                inv_P_mod_q.push_back(inv_mod(barrett_reduce_64(primes[K - 1], qi), qi));
            }
        }
    }

    // number of primes at the level with the given SEAL chain index, clamped to the first data
    // level (common/apsu/util/utils.cpp:179-189). chain_index c (data level) has c+1 primes.
    size_t level_for_chain_idx(size_t chain_idx) const { return std::min(first_L, chain_idx + 1); }

private:
    void init_level(Level &lv, size_t L)
    {
        lv.L = L;
        std::vector<u64> qp(primes.begin(), primes.begin() + L);
        lv.q.init(qp);
        lv.ntt.clear();
        for (size_t i = 0; i < L; i++) lv.ntt.push_back(&ntt[i]);
        BigUInt qprod = BigUInt::product(qp);
        lv.q_mod_t = qprod.mod_small(t.value);
        BigUInt q_div_t = qprod.div_small(t.value);
        lv.coeff_div_plain.resize(L);
        lv.upper_half_increment.resize(L);
        for (size_t i = 0; i < L; i++) {
            lv.coeff_div_plain[i] = q_div_t.mod_small(qp[i]);
            if (qp[i] <= t.value) throw std::logic_error("fast plain lift requires t < q_i");
            lv.upper_half_increment[i] = qp[i] - t.value;
        }
        lv.inv_q_last_mod_q.clear();
        for (size_t i = 0; i + 1 < L; i++)
            lv.inv_q_last_mod_q.push_back(inv_mod(barrett_reduce_64(qp[L - 1], lv.q.base[i]), lv.q.base[i]));

        // ---- BEHZ auxiliary bases (RNSTool::initialize) ----
        size_t B_size = L;
        int total_bits = qprod.bit_count();
        if (32 + t.bits + total_bits >= kInternalModBits * (int)L + kInternalModBits) B_size++;
        std::vector<u64> aux = get_primes(2 * (u64)N, kInternalModBits, B_size + 2);
        lv.m_sk.set(aux[0]);
        lv.gamma.set(aux[1]);
        std::vector<u64> Bp(aux.begin() + 2, aux.begin() + 2 + B_size);
        lv.B.init(Bp);
        lv.m_tilde.set((u64)1 << 32);
        lv.Bsk = lv.B.base;
        lv.Bsk.push_back(lv.m_sk);
        lv.Bsk_ntt.resize(lv.Bsk.size());
        for (size_t j = 0; j < lv.Bsk.size(); j++) lv.Bsk_ntt[j].init(N, lv.Bsk[j].value);
        lv.q_to_Bsk.init(lv.q, lv.Bsk);
        lv.q_to_mtilde.init(lv.q, { lv.m_tilde });
        lv.B_to_q.init(lv.B, lv.q.base);
        lv.B_to_msk.init(lv.B, { lv.m_sk });
        lv.m_tilde_mod_q.resize(L);
        for (size_t i = 0; i < L; i++) lv.m_tilde_mod_q[i] = barrett_reduce_64(lv.m_tilde.value, lv.q.base[i]);
        size_t S = lv.Bsk.size();
        lv.prod_q_mod_Bsk.resize(S);
        lv.inv_prod_q_mod_Bsk.resize(S);
        lv.inv_m_tilde_mod_Bsk.resize(S);
        for (size_t j = 0; j < S; j++) {
            lv.prod_q_mod_Bsk[j] = lv.q.prod_mod(lv.Bsk[j]);
            lv.inv_prod_q_mod_Bsk[j] = inv_mod(lv.prod_q_mod_Bsk[j], lv.Bsk[j]);
            lv.inv_m_tilde_mod_Bsk[j] = inv_mod(barrett_reduce_64(lv.m_tilde.value, lv.Bsk[j]), lv.Bsk[j]);
        }
        u64 q_mod_mt = lv.q.prod_mod(lv.m_tilde);
        lv.neg_inv_prod_q_mod_m_tilde = neg_mod(inv_mod(q_mod_mt, lv.m_tilde), lv.m_tilde);
        lv.inv_prod_B_mod_m_sk = inv_mod(lv.B.prod_mod(lv.m_sk), lv.m_sk);
        lv.prod_B_mod_q.resize(L);
        for (size_t i = 0; i < L; i++) lv.prod_B_mod_q[i] = lv.B.prod_mod(lv.q.base[i]);
    }
};

// ----------------------------------------------------------------------------------------------
// Plain data carriers.  Layouts are the ones at the reference seam:
//   Ciphertext data = u64[size][L][N];  NTT-form Plaintext = u64[L][N]; coefficient-form = u64[N]
// ----------------------------------------------------------------------------------------------
struct Ciphertext {
    size_t size = 0, L = 0;
    bool ntt = false;
    std::vector<u64> d;
    void resize(size_t N, size_t sz, size_t l)
    {
        size = sz;
        L = l;
        d.assign(sz * l * N, 0);
    }
    u64 *poly(size_t N, size_t c) { return d.data() + c * L * N; }
    const u64 *poly(size_t N, size_t c) const { return d.data() + c * L * N; }
    bool empty() const { return size == 0; }
};

struct Plaintext {
    size_t L = 0; // 0 => coefficient form mod t (N words); else NTT form at that level (L*N words)
    std::vector<u64> d;
};

// ----------------------------------------------------------------------------------------------
// BatchEncoder (SEAL batchencoder.cpp) — A.4
// ----------------------------------------------------------------------------------------------
static inline void batch_encode(const Context &ctx, const u64 *values, size_t count, u64 *out)
{
    for (size_t i = 0; i < ctx.N; i++) out[ctx.slot_map[i]] = (i < count) ? values[i] : 0;
    ctx.plain_ntt.inverse(out);
}
static inline void batch_decode(const Context &ctx, const u64 *plain, u64 *values)
{
    std::vector<u64> tmp(plain, plain + ctx.N);
    ctx.plain_ntt.forward(tmp.data());
    for (size_t i = 0; i < ctx.N; i++) values[i] = tmp[ctx.slot_map[i]];
}

// ----------------------------------------------------------------------------------------------
// Evaluator (SEAL evaluator.cpp, util/rns.cpp, util/scalingvariant.cpp) — A.5, A.6, A.7
// ----------------------------------------------------------------------------------------------
struct Evaluator {
    const Context &ctx;
    explicit Evaluator(const Context &c) : ctx(c) {}
    size_t N() const { return ctx.N; }

    // transform_to_ntt_inplace(Plaintext, parms_id): lift mod t -> RNS (fast plain lift), NTT per prime
    void plain_to_ntt(const u64 *coeff_form, size_t L, u64 *out) const
    {
        const Level &lv = ctx.levels[L];
        u64 thr = (ctx.t.value + 1) >> 1;
        for (size_t j = 0; j < L; j++) {
            u64 *o = out + j * N();
            for (size_t n = 0; n < N(); n++) o[n] = coeff_form[n] >= thr ? coeff_form[n] + lv.upper_half_increment[j] : coeff_form[n];
            lv.ntt[j]->forward(o);
        }
    }
    void to_ntt(Ciphertext &c) const
    {
        if (c.ntt) throw std::invalid_argument("already NTT form");
        for (size_t k = 0; k < c.size; k++)
            for (size_t j = 0; j < c.L; j++) ctx.ntt[j].forward(c.poly(N(), k) + j * N());
        c.ntt = true;
    }
    void from_ntt(Ciphertext &c) const
    {
        if (!c.ntt) throw std::invalid_argument("not NTT form");
        for (size_t k = 0; k < c.size; k++)
            for (size_t j = 0; j < c.L; j++) ctx.ntt[j].inverse(c.poly(N(), k) + j * N());
        c.ntt = false;
    }
    // multiply_plain, both in NTT form at the same level
    void multiply_plain_ntt(const Ciphertext &c, const u64 *plain_ntt, Ciphertext &out) const
    {
        if (!c.ntt) throw std::invalid_argument("ciphertext must be NTT form");
        out.resize(N(), c.size, c.L);
        out.ntt = true;
        for (size_t k = 0; k < c.size; k++)
            for (size_t j = 0; j < c.L; j++) {
                const Modulus &m = ctx.levels[c.L].q.base[j];
                const u64 *a = c.poly(N(), k) + j * N();
                const u64 *p = plain_ntt + j * N();
                u64 *o = out.poly(N(), k) + j * N();
                for (size_t n = 0; n < N(); n++) o[n] = mul_mod(a[n], p[n], m);
            }
    }
    // multiply_plain, both in coefficient form (multiply_plain_normal generic path)
    void multiply_plain_normal(const Ciphertext &c, const u64 *plain_coeff, Ciphertext &out) const
    {
        if (c.ntt) throw std::invalid_argument("ciphertext must be coefficient form");
        std::vector<u64> pn(c.L * N());
        plain_to_ntt(plain_coeff, c.L, pn.data());
        out = c;
        for (size_t k = 0; k < c.size; k++)
            for (size_t j = 0; j < c.L; j++) {
                const Modulus &m = ctx.levels[c.L].q.base[j];
                u64 *o = out.poly(N(), k) + j * N();
                ctx.ntt[j].forward(o);
                for (size_t n = 0; n < N(); n++) o[n] = mul_mod(o[n], pn[j * N() + n], m);
                ctx.ntt[j].inverse(o);
            }
    }
    void add_inplace(Ciphertext &a, const Ciphertext &b) const
    {
        if (a.L != b.L || a.ntt != b.ntt) throw std::invalid_argument("add: parameter mismatch");
        size_t mx = std::max(a.size, b.size), mn = std::min(a.size, b.size);
        if (a.size < mx) {
            a.d.resize(mx * a.L * N(), 0);
            a.size = mx;
        }
        for (size_t k = 0; k < mn; k++)
            for (size_t j = 0; j < a.L; j++) {
                const Modulus &m = ctx.levels[a.L].q.base[j];
                u64 *x = a.poly(N(), k) + j * N();
                const u64 *y = b.poly(N(), k) + j * N();
                for (size_t n = 0; n < N(); n++) x[n] = add_mod(x[n], y[n], m);
            }
        for (size_t k = mn; k < b.size; k++) // a was smaller: copy the tail of b
            std::memcpy(a.poly(N(), k), b.poly(N(), k), a.L * N() * sizeof(u64));
    }
    // add_plain_inplace (BFV): c0 += round(q*m/t)   (multiply_add_plain_with_scaling_variant)
    void add_plain_inplace(Ciphertext &c, const u64 *plain_coeff) const
    {
        if (c.ntt) throw std::invalid_argument("add_plain: ciphertext must be coefficient form");
        const Level &lv = ctx.levels[c.L];
        u64 thr = (ctx.t.value + 1) >> 1;
        for (size_t n = 0; n < N(); n++) {
            u128 num = (u128)plain_coeff[n] * lv.q_mod_t + thr;
            u64 fix = (u64)(num / ctx.t.value);
            for (size_t j = 0; j < c.L; j++) {
                const Modulus &m = lv.q.base[j];
                u64 scaled = add_mod(mul_mod(plain_coeff[n], lv.coeff_div_plain[j], m), barrett_reduce_64(fix, m), m);
                u64 *x = c.poly(N(), 0) + j * N();
                x[n] = add_mod(x[n], scaled, m);
            }
        }
    }
    // mod_switch_to_next_inplace (BFV): divide_and_round_q_last per polynomial
    void mod_switch_to_next(Ciphertext &c) const
    {
        if (c.ntt) throw std::invalid_argument("mod_switch: BFV ciphertext must be coefficient form");
        if (c.L < 2) throw std::invalid_argument("mod_switch: already at last level");
        const Level &lv = ctx.levels[c.L];
        size_t L = c.L;
        Ciphertext out;
        out.resize(N(), c.size, L - 1);
        const Modulus &ql = lv.q.base[L - 1];
        u64 half = ql.value >> 1;
        for (size_t k = 0; k < c.size; k++) {
            const u64 *last = c.poly(N(), k) + (L - 1) * N();
            for (size_t j = 0; j + 1 < L; j++) {
                const Modulus &m = lv.q.base[j];
                u64 half_mod = barrett_reduce_64(half, m);
                const u64 *x = c.poly(N(), k) + j * N();
                u64 *o = out.poly(N(), k) + j * N();
                for (size_t n = 0; n < N(); n++) {
                    u64 a = add_mod(last[n], half, ql);
                    u64 tmp = sub_mod(barrett_reduce_64(a, m), half_mod, m);
                    o[n] = mul_mod(sub_mod(x[n], tmp, m), lv.inv_q_last_mod_q[j], m);
                }
            }
        }
        c = std::move(out);
    }
    void mod_switch_to(Ciphertext &c, size_t L) const
    {
        if (c.L < L) throw std::invalid_argument("mod_switch_to: cannot switch to higher level");
        while (c.L > L) mod_switch_to_next(c);
    }

    // ---- BEHZ helpers on one polynomial ----
    // steps (1)-(2): x (base q) -> x' (base Bsk)
    void behz_extend(const Level &lv, const u64 *x, u64 *out_Bsk) const
    {
        size_t L = lv.L, S = lv.Bsk.size(), n = N();
        std::vector<u64> tmp(L * n), y((S + 1) * n);
        for (size_t i = 0; i < L; i++)
            for (size_t k = 0; k < n; k++) tmp[i * n + k] = mul_mod(x[i * n + k], lv.m_tilde_mod_q[i], lv.q.base[i]);
        lv.q_to_Bsk.fast_convert_array(tmp.data(), y.data(), n);
        lv.q_to_mtilde.fast_convert_array(tmp.data(), y.data() + S * n, n);
        // sm_mrq
        const u64 *ymt = y.data() + S * n;
        u64 mt = lv.m_tilde.value, mt_half = mt >> 1;
        for (size_t j = 0; j < S; j++) {
            const Modulus &m = lv.Bsk[j];
            for (size_t k = 0; k < n; k++) {
                u64 r = mul_mod(ymt[k], lv.neg_inv_prod_q_mod_m_tilde, lv.m_tilde);
                if (r >= mt_half) r += m.value - mt;
                u64 v = add_mod(mul_mod(r, lv.prod_q_mod_Bsk[j], m), y[j * n + k], m);
                out_Bsk[j * n + k] = mul_mod(v, lv.inv_m_tilde_mod_Bsk[j], m);
            }
        }
    }
    // steps (6)-(8) on one output polynomial: dq (base q, coefficient form), dB (base Bsk) -> out (base q)
    void behz_scale_down(const Level &lv, const u64 *dq, const u64 *dB, u64 *out) const
    {
        size_t L = lv.L, S = lv.Bsk.size(), n = N(), Bs = lv.B.size();
        u64 t = ctx.t.value;
        std::vector<u64> tq(L * n), tB(S * n), fl(S * n);
        for (size_t i = 0; i < L; i++)
            for (size_t k = 0; k < n; k++) tq[i * n + k] = mul_mod(dq[i * n + k], barrett_reduce_64(t, lv.q.base[i]), lv.q.base[i]);
        for (size_t j = 0; j < S; j++)
            for (size_t k = 0; k < n; k++) tB[j * n + k] = mul_mod(dB[j * n + k], barrett_reduce_64(t, lv.Bsk[j]), lv.Bsk[j]);
        // fast_floor
        lv.q_to_Bsk.fast_convert_array(tq.data(), fl.data(), n);
        for (size_t j = 0; j < S; j++) {
            const Modulus &m = lv.Bsk[j];
            for (size_t k = 0; k < n; k++) fl[j * n + k] = mul_mod(sub_mod(tB[j * n + k], fl[j * n + k], m), lv.inv_prod_q_mod_Bsk[j], m);
        }
        // fastbconv_sk
        std::vector<u64> alpha(n);
        lv.B_to_q.fast_convert_array(fl.data(), out, n);
        lv.B_to_msk.fast_convert_array(fl.data(), alpha.data(), n);
        const u64 *fsk = fl.data() + Bs * n;
        u64 msk_half = lv.m_sk.value >> 1;
        for (size_t k = 0; k < n; k++) alpha[k] = mul_mod(sub_mod(alpha[k], fsk[k], lv.m_sk), lv.inv_prod_B_mod_m_sk, lv.m_sk);
        for (size_t i = 0; i < L; i++) {
            const Modulus &m = lv.q.base[i];
            u64 pB = lv.prod_B_mod_q[i], npB = neg_mod(pB, m);
            for (size_t k = 0; k < n; k++) {
                u64 a = alpha[k];
                if (a > msk_half)
                    out[i * n + k] = add_mod(mul_mod(barrett_reduce_64(lv.m_sk.value - a, m), pB, m), out[i * n + k], m);
                else
                    out[i * n + k] = add_mod(mul_mod(barrett_reduce_64(a, m), npB, m), out[i * n + k], m);
            }
        }
    }

    // Evaluator::multiply (bfv_multiply); square(x) == multiply(x, x) bit-for-bit
    void multiply(const Ciphertext &a, const Ciphertext &b, Ciphertext &out) const
    {
        if (a.ntt || b.ntt) throw std::invalid_argument("multiply: operands cannot be NTT form");
        if (a.L != b.L) throw std::invalid_argument("multiply: level mismatch");
        const Level &lv = ctx.levels[a.L];
        size_t L = a.L, S = lv.Bsk.size(), n = N();
        size_t dsz = a.size + b.size - 1;
        auto prep = [&](const Ciphertext &c, std::vector<u64> &cq, std::vector<u64> &cB) {
            cq.assign(c.d.begin(), c.d.end());
            cB.assign(c.size * S * n, 0);
            for (size_t k = 0; k < c.size; k++) {
                behz_extend(lv, c.poly(n, k), cB.data() + k * S * n);
                for (size_t j = 0; j < L; j++) lv.ntt[j]->forward(cq.data() + (k * L + j) * n);
                for (size_t j = 0; j < S; j++) lv.Bsk_ntt[j].forward(cB.data() + (k * S + j) * n);
            }
        };
        std::vector<u64> aq, aB, bq, bB;
        prep(a, aq, aB);
        prep(b, bq, bB);
        std::vector<u64> dq(dsz * L * n, 0), dB(dsz * S * n, 0);
        for (size_t o = 0; o < dsz; o++) {
            for (size_t i = 0; i < a.size; i++) {
                if (o < i || o - i >= b.size) continue;
                size_t k2 = o - i;
                for (size_t j = 0; j < L; j++) {
                    const Modulus &m = lv.q.base[j];
                    const u64 *x = aq.data() + (i * L + j) * n, *y = bq.data() + (k2 * L + j) * n;
                    u64 *d = dq.data() + (o * L + j) * n;
                    for (size_t k = 0; k < n; k++) d[k] = add_mod(d[k], mul_mod(x[k], y[k], m), m);
                }
                for (size_t j = 0; j < S; j++) {
                    const Modulus &m = lv.Bsk[j];
                    const u64 *x = aB.data() + (i * S + j) * n, *y = bB.data() + (k2 * S + j) * n;
                    u64 *d = dB.data() + (o * S + j) * n;
                    for (size_t k = 0; k < n; k++) d[k] = add_mod(d[k], mul_mod(x[k], y[k], m), m);
                }
            }
        }
        Ciphertext res;
        res.resize(n, dsz, L);
        for (size_t o = 0; o < dsz; o++) {
            for (size_t j = 0; j < L; j++) lv.ntt[j]->inverse(dq.data() + (o * L + j) * n);
            for (size_t j = 0; j < S; j++) lv.Bsk_ntt[j].inverse(dB.data() + (o * S + j) * n);
            behz_scale_down(lv, dq.data() + o * L * n, dB.data() + o * S * n, res.poly(n, o));
        }
        out = std::move(res);
    }

    // relinearize_inplace (size 3 -> 2): switch_key_inplace on c2 with relin key index 0.
    // keys: u64[K-1][2][K][N] (NTT form at the key level)
    void relinearize(Ciphertext &c, const u64 *keys) const
    {
        if (c.size != 3) throw std::invalid_argument("relinearize: size must be 3");
        if (c.ntt) throw std::invalid_argument("relinearize: BFV ciphertext must be coefficient form");
        size_t L = c.L, K = ctx.K, n = N();
        const u64 *target = c.poly(n, 2);
        size_t R = L + 1;
        std::vector<u64> prod(2 * R * n, 0); // [c][I][n]
        std::vector<u64> tn(n);
        std::vector<u128> acc(n);
        for (size_t I = 0; I < R; I++) {
            size_t key_index = (I == L) ? K - 1 : I;
            const Modulus km(ctx.primes[key_index]);
            for (size_t cc = 0; cc < 2; cc++) {
                std::fill(acc.begin(), acc.end(), (u128)0);
                for (size_t J = 0; J < L; J++) {
                    for (size_t k = 0; k < n; k++) tn[k] = barrett_reduce_64(target[J * n + k], km);
                    ctx.ntt[key_index].forward(tn.data());
                    const u64 *key = keys + ((J * 2 + cc) * K + key_index) * n;
                    for (size_t k = 0; k < n; k++) {
                        acc[k] += (u128)tn[k] * key[k];
                        if (acc[k] >> 127) acc[k] = barrett_reduce_128(acc[k], km);
                    }
                }
                u64 *p = prod.data() + (cc * R + I) * n;
                for (size_t k = 0; k < n; k++) p[k] = barrett_reduce_128(acc[k], km);
            }
        }
        const Modulus P(ctx.primes[K - 1]);
        u64 half = P.value >> 1;
        for (size_t cc = 0; cc < 2; cc++) {
            u64 *last = prod.data() + (cc * R + L) * n;
            ctx.ntt[K - 1].inverse(last);
            for (size_t k = 0; k < n; k++) last[k] = add_mod(last[k], half, P);
            for (size_t i = 0; i < L; i++) {
                const Modulus qi(ctx.primes[i]);
                u64 half_mod = barrett_reduce_64(half, qi);
                u64 *pi = prod.data() + (cc * R + i) * n;
                ctx.ntt[i].inverse(pi);
                u64 *dst = c.poly(n, cc) + i * n;
                for (size_t k = 0; k < n; k++) {
                    u64 delta = sub_mod(barrett_reduce_64(last[k], qi), half_mod, qi);
                    u64 v = mul_mod(sub_mod(pi[k], delta, qi), ctx.inv_P_mod_q[i], qi);
                    dst[k] = add_mod(dst[k], v, qi);
                }
            }
        }
        c.d.resize(2 * L * n);
        c.size = 2;
    }
};

// ----------------------------------------------------------------------------------------------
// Harness-only: key generation, symmetric encryption, decryption (the *sender's* job in APSU,
// sender/apsu/sender_ddh.cpp:127-141, plaintext_powers.cpp:33-49, result_package.cpp:175-213).
// Any valid BFV key/ciphertext works — parity is about evaluation — so a seeded splitmix64 PRNG
// replaces SEAL's blake2xb sampler.
// ----------------------------------------------------------------------------------------------
struct SplitMix64 {
    u64 s;
    explicit SplitMix64(u64 seed) : s(seed) {}
    u64 next()
    {
        u64 z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    u64 below(u64 bound) { return (u64)(((u128)next() * bound) >> 64); }
};

struct KeyMaterial {
    std::vector<int8_t> s;      // ternary secret, coefficient form
    std::vector<u64> s_ntt;     // [K][N]
    std::vector<u64> relin;     // [K-1][2][K][N]
};

static inline int sample_cbd_error(SplitMix64 &rng)
{
    // centered binomial, variance 21/2 ~ sigma 3.24 (close to SEAL's 3.2 clipped normal)
    u64 r = rng.next();
    int a = __builtin_popcountll(r & 0x1FFFFF), b = __builtin_popcountll((r >> 21) & 0x1FFFFF);
    return a - b;
}

static inline void keygen(const Context &ctx, u64 seed, KeyMaterial &km)
{
    size_t N = ctx.N, K = ctx.K;
    SplitMix64 rng(seed);
    km.s.resize(N);
    for (size_t i = 0; i < N; i++) km.s[i] = (int8_t)((int)rng.below(3) - 1);
    km.s_ntt.assign(K * N, 0);
    for (size_t j = 0; j < K; j++) {
        u64 q = ctx.primes[j];
        for (size_t i = 0; i < N; i++) km.s_ntt[j * N + i] = km.s[i] < 0 ? q - 1 : (u64)km.s[i];
        ctx.ntt[j].forward(km.s_ntt.data() + j * N);
    }
    km.relin.clear();
    if (K < 2) return;
    km.relin.assign((K - 1) * 2 * K * N, 0);
    std::vector<int> e(N);
    std::vector<u64> en(N);
    for (size_t J = 0; J + 1 < K; J++) {
        for (size_t i = 0; i < N; i++) e[i] = sample_cbd_error(rng);
        for (size_t I = 0; I < K; I++) {
            const Modulus m(ctx.primes[I]);
            u64 *c0 = km.relin.data() + ((J * 2 + 0) * K + I) * N;
            u64 *c1 = km.relin.data() + ((J * 2 + 1) * K + I) * N;
            for (size_t i = 0; i < N; i++) {
                c1[i] = rng.below(m.value); // uniform a, already "NTT form"
                en[i] = e[i] < 0 ? m.value - (u64)(-e[i]) : (u64)e[i];
            }
            ctx.ntt[I].forward(en.data());
            const u64 *s = km.s_ntt.data() + I * N;
            u64 factor = barrett_reduce_64(ctx.primes[K - 1], m);
            for (size_t i = 0; i < N; i++) {
                u64 v = neg_mod(add_mod(mul_mod(c1[i], s[i], m), en[i], m), m);
                if (I == J) v = add_mod(v, mul_mod(mul_mod(s[i], s[i], m), factor, m), m);
                c0[i] = v;
            }
        }
    }
}

// encrypt_symmetric of a coefficient-form plaintext at the first data level
static inline void encrypt_symmetric(const Context &ctx, const KeyMaterial &km, const u64 *plain, u64 seed, Ciphertext &out)
{
    size_t N = ctx.N, L = ctx.first_L;
    SplitMix64 rng(seed);
    out.resize(N, 2, L);
    out.ntt = false;
    std::vector<int> e(N);
    for (size_t i = 0; i < N; i++) e[i] = sample_cbd_error(rng);
    std::vector<u64> a(N), en(N);
    for (size_t j = 0; j < L; j++) {
        const Modulus m(ctx.primes[j]);
        for (size_t i = 0; i < N; i++) {
            a[i] = rng.below(m.value);
            en[i] = e[i] < 0 ? m.value - (u64)(-e[i]) : (u64)e[i];
        }
        // c1 = a (sampled in NTT domain, converted back), c0 = -(a*s + e)
        ctx.ntt[j].forward(en.data());
        u64 *c0 = out.poly(N, 0) + j * N, *c1 = out.poly(N, 1) + j * N;
        const u64 *s = km.s_ntt.data() + j * N;
        for (size_t i = 0; i < N; i++) {
            c0[i] = neg_mod(add_mod(mul_mod(a[i], s[i], m), en[i], m), m);
            c1[i] = a[i];
        }
        ctx.ntt[j].inverse(c0);
        ctx.ntt[j].inverse(c1);
    }
    Evaluator(ctx).add_plain_inplace(out, plain);
}

// decrypt a single-prime (last level) ciphertext exactly: m = round(t * [c0 + c1 s (+ c2 s^2)]_q / q) mod t.
// Also returns the invariant-noise budget in bits (min over coefficients), -1 if undecryptable is unknowable here.
static inline int decrypt_last_level(const Context &ctx, const KeyMaterial &km, const Ciphertext &c, u64 *plain_out)
{
    if (c.L != 1) throw std::invalid_argument("decrypt_last_level: ciphertext must have one prime");
    size_t N = ctx.N;
    const Modulus q(ctx.primes[0]);
    std::vector<u64> acc(N, 0), tmp(N);
    const u64 *s = km.s_ntt.data();
    std::vector<u64> spow(s, s + N);
    std::vector<u64> cc;
    if (c.ntt) throw std::invalid_argument("decrypt: coefficient form expected");
    for (size_t k = 0; k < c.size; k++) {
        tmp.assign(c.poly(N, k), c.poly(N, k) + N);
        if (k == 0) {
            acc = tmp;
            continue;
        }
        ctx.ntt[0].forward(tmp.data());
        for (size_t i = 0; i < N; i++) tmp[i] = mul_mod(tmp[i], spow[i], q);
        ctx.ntt[0].inverse(tmp.data());
        for (size_t i = 0; i < N; i++) acc[i] = add_mod(acc[i], tmp[i], q);
        for (size_t i = 0; i < N; i++) spow[i] = mul_mod(spow[i], s[i], q);
    }
    u64 t = ctx.t.value;
    int min_budget = 1 << 20;
    for (size_t i = 0; i < N; i++) {
        u128 num = (u128)acc[i] * t;
        u64 r = (u64)(num % q.value);
        u64 m = (u64)((num + (q.value >> 1)) / q.value) % t;
        plain_out[i] = m;
        // invariant noise = |t*x mod q centred| / q ; budget = -log2(2 * noise)
        u64 dist = r > q.value - r ? q.value - r : r;
        int bits = 0;
        for (u64 x = dist; x; x >>= 1) bits++;
        int budget = q.bits - bits - 1;
        if (budget < min_budget) min_budget = budget;
    }
    return min_budget;
}

} // namespace orc
