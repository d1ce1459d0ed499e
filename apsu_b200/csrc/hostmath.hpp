// Host-side number theory used once at context creation: prime selection, primitive roots, modular
// inverses, products of primes.  Follows the selection rules of SEAL 3.7 that the reference relies
// on through CoeffModulus::Create / PlainModulus::Batching / NTTTables / RNSTool
// (reference call sites: common/apsu/psu_params.cpp:355-365, common/apsu/crypto_context.h:33-40;
// rules in SURVEY.md A.1-A.3, A.6).  Nothing here is performance critical.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace apsu_b200 {
namespace hm {

using u64 = std::uint64_t;
using u128 = unsigned __int128;

inline u64 mulm(u64 a, u64 b, u64 m) { return (u64)((u128)a * b % m); }
inline u64 addm(u64 a, u64 b, u64 m) { return (u64)(((u128)a + b) % m); }
inline u64 subm(u64 a, u64 b, u64 m) { return (a % m + m - b % m) % m; }
inline u64 powm(u64 a, u64 e, u64 m)
{
    u64 r = 1 % m;
    a %= m;
    for (; e; e >>= 1, a = mulm(a, a, m))
        if (e & 1) r = mulm(r, a, m);
    return r;
}
inline int bit_length(u64 v)
{
    int b = 0;
    for (; v; v >>= 1) b++;
    return b;
}
// modular inverse for any modulus (m_tilde = 2^32 is not prime)
inline u64 invm(u64 a, u64 m)
{
    __int128 g0 = m, g1 = a % m, x0 = 0, x1 = 1;
    while (g1) {
        __int128 q = g0 / g1, tmp = g0 - q * g1;
        g0 = g1, g1 = tmp;
        tmp = x0 - q * x1;
        x0 = x1, x1 = tmp;
    }
    if (g0 != 1) throw std::logic_error("modular inverse does not exist");
    return (u64)(x0 < 0 ? x0 + m : x0);
}

inline bool miller_rabin(u64 n)
{
    if (n < 4) return n == 2 || n == 3;
    if (!(n & 1)) return false;
    u64 d = n - 1;
    int s = 0;
    while (!(d & 1)) d >>= 1, s++;
    for (u64 a : { 2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull }) {
        if (a % n == 0) continue;
        u64 x = powm(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (int i = 1; i < s && witness; i++) {
            x = mulm(x, x, n);
            if (x == n - 1) witness = false;
        }
        if (witness) return false;
    }
    return true;
}

// `count` largest primes with exactly `bits` bits that are 1 modulo `factor`, largest first
inline std::vector<u64> primes_below_pow2(u64 factor, int bits, size_t count)
{
    if (bits < 2 || bits > 61) throw std::invalid_argument("bit_size is invalid");
    std::vector<u64> found;
    const u64 floor_ = 1ull << (bits - 1);
    for (u64 cand = ((1ull << bits) - 1) / factor * factor + 1; found.size() < count && cand > floor_; cand -= factor)
        if (miller_rabin(cand)) found.push_back(cand);
    if (found.size() < count) throw std::logic_error("failed to find enough qualifying primes");
    return found;
}

// smallest primitive `order`-th root of unity modulo prime p (order a power of two dividing p-1)
inline u64 min_primitive_root(u64 order, u64 p)
{
    if ((p - 1) % order) throw std::invalid_argument("modulus does not support the transform size");
    u64 g = 0;
    for (u64 base = 2; !g; base++) {
        u64 c = powm(base, (p - 1) / order, p);
        if (powm(c, order / 2, p) == p - 1) g = c;
    }
    // the primitive roots are exactly the odd powers of g
    u64 step = mulm(g, g, p), best = g, cur = g;
    for (u64 k = 1; k < order / 2; k++) {
        cur = mulm(cur, step, p);
        if (cur < best) best = cur;
    }
    return best;
}

inline unsigned bitrev(unsigned x, int bits)
{
    unsigned r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

// product of primes modulo m, optionally skipping one index
inline u64 prod_mod(const std::vector<u64> &primes, u64 m, int skip = -1)
{
    u64 r = 1 % m;
    for (size_t i = 0; i < primes.size(); i++)
        if ((int)i != skip) r = mulm(r, primes[i] % m, m);
    return r;
}

// little-endian multiword product of primes, with the two reductions add_plain needs
struct Wide {
    std::vector<u64> limb{ 1 };
    void mul(u64 f)
    {
        u64 carry = 0;
        for (auto &l : limb) {
            u128 v = (u128)l * f + carry;
            l = (u64)v;
            carry = (u64)(v >> 64);
        }
        if (carry) limb.push_back(carry);
    }
    u64 mod(u64 m) const
    {
        u128 r = 0;
        for (size_t i = limb.size(); i--;) r = ((r << 64) | limb[i]) % m;
        return (u64)r;
    }
    Wide div(u64 d) const
    {
        Wide q;
        q.limb.assign(limb.size(), 0);
        u128 r = 0;
        for (size_t i = limb.size(); i--;) {
            u128 cur = (r << 64) | limb[i];
            q.limb[i] = (u64)(cur / d);
            r = cur % d;
        }
        return q;
    }
    int bits() const
    {
        for (size_t i = limb.size(); i--;)
            if (limb[i]) return (int)(64 * i) + bit_length(limb[i]);
        return 0;
    }
};

} // namespace hm
} // namespace apsu_b200
