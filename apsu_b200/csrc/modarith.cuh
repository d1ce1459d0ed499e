// 64-bit modular arithmetic for sm_100a.  Moduli are < 2^61 (user primes <= 60 bits, BEHZ auxiliary
// primes 61 bits, m_tilde = 2^32), so lazy Harvey ranges [0,4q) fit a u64.  B200 has no native 64x64
// multiplier: mul.hi.u64 / mul.lo.u64 compile to IMAD.WIDE.U32 chains on the FMA pipe, which is the
// integer roofline these kernels run against.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace apsu_b200 {

using u64 = unsigned long long;
using u32 = unsigned int;

// modulus with its Barrett constants: floor(2^128 / q) (two words, for arbitrary 128-bit inputs) and
// mu = floor(2^(64+sh) / q) with sh = bit_length(q) - 1 (one word, for inputs below 2^(64+sh) such as sums of a
// few products of reduced operands: barrett_prod)
struct DMod {
    u64 q;
    u64 r0, r1;
    u64 mu;
    u32 sh, pad_;
};

// constant multiplicand in Shoup form: quot = floor(op * 2^64 / q)
struct DShoup {
    u64 op, quot;
};

// Programmatic dependent launch (launch_pdl, context.hpp): every kernel of the query path starts with this.  The wait
// returns once the preceding kernel of the stream has completed and its writes are visible (immediately when the launch
// carries no programmatic edge); the trigger then lets the NEXT kernel's CTAs be scheduled while this one runs, so its
// launch latency is hidden behind this kernel instead of following it (they block in their own wait until this grid
// is complete, so every data dependency of the launch sequence holds transitively).
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.wait;\n\tgriddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ u64 mulhi(u64 a, u64 b) { return __umul64hi(a, b); }

// x * w mod q for x < 2^64, result in [0, 2q)
__device__ __forceinline__ u64 mul_shoup_lazy(u64 x, u64 w, u64 wq, u64 q) { return w * x - mulhi(x, wq) * q; }
// x * w mod q for ANY x < 2^64 with an approximate quotient: result in [0, 3q).  nq = 2^64 - q, wq = floor(w*2^64/q).
// The quotient estimate drops the low x low partial product (at most 1 short), and the remainder is formed as
// lo64(w*x + qhat*nq): 4 IMAD.WIDE + 1 IMAD.HI + 4 IMAD in SASS instead of the 6 IMAD.WIDE + 4 IMAD + carries of the
// exact form.  The integer-multiply pipe (IMAD.WIDE issues at quarter rate on sm_100a) is what bounds the NTT, so
// the butterflies use this form and keep their values in [0, 6q) (q < 2^61.4).
__device__ __forceinline__ u64 mul_shoup_lazy3(u64 y, u64 w, u64 wq, u64 nq)
{
    const u32 yl = (u32)y, yh = (u32)(y >> 32), wl = (u32)w, wh = (u32)(w >> 32), vl = (u32)wq, vh = (u32)(wq >> 32), nl = (u32)nq, nh = (u32)(nq >> 32);
    u32 q0, q1, r0, r1;
    asm("{\n\t"
        ".reg .u32 a0, a1, s0, s1, c;\n\t"
        "mul.lo.u32 a0, %3, %6;\n\t"          // A = yh*vl
        "mul.hi.u32 a1, %3, %6;\n\t"
        "mad.lo.cc.u32 s0, %2, %7, a0;\n\t"   // S = yl*vh + A, carry c
        "madc.hi.cc.u32 s1, %2, %7, a1;\n\t"
        "addc.u32 c, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, %3, %7, s1;\n\t"   // Q = yh*vh + (S.hi : c)
        "madc.hi.u32 %1, %3, %7, c;\n\t"
        "}"
        : "=&r"(q0), "=&r"(q1)
        : "r"(yl), "r"(yh), "r"(wl), "r"(wh), "r"(vl), "r"(vh));
    asm("{\n\t"
        "mul.lo.u32 %0, %4, %2;\n\t"          // R = wl*yl
        "mul.hi.u32 %1, %4, %2;\n\t"
        "mad.lo.cc.u32 %0, %6, %8, %0;\n\t"   // R += q0*nl
        "madc.hi.u32 %1, %6, %8, %1;\n\t"
        "mad.lo.u32 %1, %4, %3, %1;\n\t"      // hi word += wl*yh + wh*yl + q0*nh + q1*nl
        "mad.lo.u32 %1, %5, %2, %1;\n\t"
        "mad.lo.u32 %1, %6, %9, %1;\n\t"
        "mad.lo.u32 %1, %7, %8, %1;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1)
        : "r"(yl), "r"(yh), "r"(wl), "r"(wh), "r"(q0), "r"(q1), "r"(nl), "r"(nh));
    return ((u64)r1 << 32) | r0;
}
// x >= m ? x - m : x with the selection on the borrow of the subtraction (IADD3, IADD3.X, 2 SEL)
__device__ __forceinline__ u64 csub(u64 x, u64 m)
{
    u32 xl = (u32)x, xh = (u32)(x >> 32), ml = (u32)m, mh = (u32)(m >> 32), dl, dh, b;
    asm("{\n\t"
        "sub.cc.u32 %0, %3, %5;\n\t"
        "subc.cc.u32 %1, %4, %6;\n\t"
        "subc.u32 %2, 0, 0;\n\t" // all ones when x < m
        "}"
        : "=r"(dl), "=r"(dh), "=r"(b)
        : "r"(xl), "r"(xh), "r"(ml), "r"(mh));
    dl = b ? xl : dl;
    dh = b ? xh : dh;
    return ((u64)dh << 32) | dl;
}
// result in [0, q)
__device__ __forceinline__ u64 mul_shoup(u64 x, u64 w, u64 wq, u64 q)
{
    u64 r = mul_shoup_lazy(x, w, wq, q);
    return r >= q ? r - q : r;
}
__device__ __forceinline__ u64 mul_shoup(u64 x, const DShoup &s, u64 q) { return mul_shoup(x, s.op, s.quot, q); }

// (hi:lo) mod q, any 128-bit input, canonical result.  Barrett with ratio floor(2^128/q); the
// quotient estimate is at most 2 short, hence two conditional subtractions.
__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, const DMod &m)
{
    // qhat = floor((hi:lo) * (r1:r0) / 2^128), low carries of lo*r0 dropped
    u64 carry = mulhi(lo, m.r0);
    u64 t_lo = lo * m.r1, t_hi = mulhi(lo, m.r1);
    u64 s1 = t_lo + carry;
    u64 c1 = s1 < t_lo;
    u64 tmp3 = t_hi + c1;
    u64 u_lo = hi * m.r0, u_hi = mulhi(hi, m.r0);
    u64 s2 = s1 + u_lo;
    u64 c2 = s2 < u_lo;
    u64 qhat = hi * m.r1 + tmp3 + u_hi + c2;
    u64 r = lo - qhat * m.q;
    if (r >= m.q) r -= m.q;
    if (r >= m.q) r -= m.q;
    return r;
}
// x mod q for a single word
__device__ __forceinline__ u64 barrett64(u64 x, const DMod &m)
{
    u64 r = x - mulhi(x, m.r1) * m.q;
    if (r >= m.q) r -= m.q;
    if (r >= m.q) r -= m.q;
    return r;
}
// x mod q when the caller knows how large x can be relative to q (decided on the host from the moduli, uniform per
// launch): kind 0: x < q, 1: x < 2q (one conditional subtraction), 2: anything (Barrett).  The mod-switch and mod-down
// kernels reduce a residue of one prime modulo the others; for primes of similar size that is not a multiplication.
__device__ __forceinline__ u64 reduce_known(u64 x, const DMod &m, u32 kind)
{
    if (kind == 0) return x;
    if (kind == 1) return x >= m.q ? x - m.q : x;
    return barrett64(x, m);
}
// (hi:lo) mod q for (hi:lo) < 2^(64+sh), i.e. hi < 2^sh — products of reduced operands and short sums of them.
// One 64x64 high product instead of barrett128's four: xh = (hi:lo) >> sh fits a word, qhat = floor(xh*mu/2^64)
// is at most 2 short of the true quotient (both truncations lose less than 1), hence two conditional subtractions.
// The integer pipe that executes IMAD.WIDE is what bounds the element-wise kernels, so this matters.
__device__ __forceinline__ u64 barrett_prod(u64 lo, u64 hi, const DMod &m)
{
    const u64 xh = (lo >> m.sh) | (hi << (64 - m.sh));
    u64 r = lo - mulhi(xh, m.mu) * m.q;
    if (r >= m.q) r -= m.q;
    if (r >= m.q) r -= m.q;
    return r;
}
__device__ __forceinline__ u64 mul_mod(u64 a, u64 b, const DMod &m) { return barrett_prod(a * b, mulhi(a, b), m); }
// [0, 8q) -> [0, q)
__device__ __forceinline__ u64 csub(u64 x, u64 m);
__device__ __forceinline__ u64 reduce_8q(u64 x, u64 q) { return csub(csub(csub(x, 4 * q), 2 * q), q); }
__device__ __forceinline__ u64 add_mod(u64 a, u64 b, u64 q)
{
    u64 s = a + b;
    return s >= q ? s - q : s;
}
__device__ __forceinline__ u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }

// 128-bit accumulator: acc += a*b  (products < 2^122, callers bound the number of summands)
struct Acc128 {
    u64 lo, hi;
};
__device__ __forceinline__ void mac128(Acc128 &acc, u64 a, u64 b)
{
    u64 pl = a * b, ph = mulhi(a, b);
    asm("add.cc.u64 %0, %0, %2;\n\t"
        "addc.u64 %1, %1, %3;"
        : "+l"(acc.lo), "+l"(acc.hi)
        : "l"(pl), "l"(ph));
}

} // namespace apsu_b200
