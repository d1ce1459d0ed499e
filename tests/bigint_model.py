"""Independent Python big-integer model of the SEAL-level operations on the path,
written from the formulas with plain `%` arithmetic and schoolbook negacyclic products.  It is slow and
only used at toy ring sizes (N = 32..128) to cross-check the C++ oracle bit for bit, and at real sizes
for exact CRT decryption (noise / correctness checks).  TEST INFRASTRUCTURE."""
from __future__ import annotations

from math import prod


def bitrev(x: int, bits: int) -> int:
    return int(format(x, f"0{bits}b")[::-1], 2) if bits else 0


def negacyclic_mul(a, b, q):
    n = len(a)
    out = [0] * n
    for i, ai in enumerate(a):
        if not ai:
            continue
        for j, bj in enumerate(b):
            k = i + j
            if k < n:
                out[k] = (out[k] + ai * bj) % q
            else:
                out[k - n] = (out[k - n] - ai * bj) % q
    return out


def ntt_by_definition(a, q, psi):
    """out[k] = a(psi^(2*bitrev(k)+1)) — the definition of SEAL's forward transform (A.3)."""
    n = len(a)
    bits = n.bit_length() - 1
    out = []
    for k in range(n):
        x = pow(psi, 2 * bitrev(k, bits) + 1, q)
        acc, xp = 0, 1
        for c in a:
            acc = (acc + c * xp) % q
            xp = xp * x % q
        out.append(acc)
    return out


def fastbconv(x, ibase, obase):
    """x[i][n] residues mod ibase[i] -> list per obase modulus (BaseConverter::fast_convert_array, A.6)."""
    P = prod(ibase)
    n = len(x[0])
    tmp = [[x[i][k] * pow(P // p % p, -1, p) % p for k in range(n)] for i, p in enumerate(ibase)]
    return [[sum(tmp[i][k] * (P // p % m) for i, p in enumerate(ibase)) % m for k in range(n)] for m in obase]


def behz_multiply(a, b, q, t, m_sk, B):
    """BFV multiply of ciphertexts a, b (lists of polys, each poly = list per prime of coefficient lists).
    Follows A.6 steps (1)-(8); returns the size-(|a|+|b|-1) ciphertext in base q."""
    mt = 1 << 32
    Bsk = list(B) + [m_sk]
    Q, PB = prod(q), prod(B)
    n = len(a[0][0])

    def extend(poly):
        xt = [[c * mt % p for c in poly[i]] for i, p in enumerate(q)]
        y = fastbconv(xt, q, Bsk + [mt])
        ymt = y[-1]
        r = [v * ((-pow(Q, -1, mt)) % mt) % mt for v in ymt]
        out = []
        for j, m in enumerate(Bsk):
            row = []
            for k in range(n):
                rr = r[k] + (m - mt) if r[k] >= mt // 2 else r[k]
                row.append((y[j][k] + Q * rr) * pow(mt, -1, m) % m)
            out.append(row)
        return out

    aq, bq = a, b
    aB, bB = [extend(p) for p in a], [extend(p) for p in b]
    dsz = len(a) + len(b) - 1
    res = []
    for o in range(dsz):
        dq = [[0] * n for _ in q]
        dB = [[0] * n for _ in Bsk]
        for i in range(len(a)):
            k2 = o - i
            if k2 < 0 or k2 >= len(b):
                continue
            for j, p in enumerate(q):
                pr = negacyclic_mul(aq[i][j], bq[k2][j], p)
                dq[j] = [(x + y) % p for x, y in zip(dq[j], pr)]
            for j, m in enumerate(Bsk):
                pr = negacyclic_mul(aB[i][j], bB[k2][j], m)
                dB[j] = [(x + y) % m for x, y in zip(dB[j], pr)]
        tq = [[c * t % p for c in dq[j]] for j, p in enumerate(q)]
        tB = [[c * t % m for c in dB[j]] for j, m in enumerate(Bsk)]
        conv = fastbconv(tq, q, Bsk)
        f = [[(tB[j][k] - conv[j][k]) * pow(Q, -1, m) % m for k in range(n)] for j, m in enumerate(Bsk)]
        toq = fastbconv(f[:-1], B, q)
        tosk = fastbconv(f[:-1], B, [m_sk])[0]
        alpha = [(tosk[k] - f[-1][k]) * pow(PB, -1, m_sk) % m_sk for k in range(n)]
        out = []
        for j, p in enumerate(q):
            row = []
            for k in range(n):
                if alpha[k] > m_sk // 2:
                    row.append((toq[j][k] + (m_sk - alpha[k]) * (PB % p)) % p)
                else:
                    row.append((toq[j][k] - alpha[k] * (PB % p)) % p)
            out.append(row)
        res.append(out)
    return res


def mod_switch_next(poly, q):
    """divide_and_round_q_last on one RNS polynomial (A.5)."""
    qk = q[-1]
    half = qk >> 1
    last = [(c + half) % qk for c in poly[-1]]
    out = []
    for i, qi in enumerate(q[:-1]):
        inv = pow(qk, -1, qi)
        out.append([((poly[i][k] - (last[k] % qi - half % qi)) * inv) % qi for k in range(len(last))])
    return out


def add_plain(c0, plain, q, t):
    """c0 += round(q*m/t) per coefficient (scaling variant, A.5)."""
    Q = prod(q)
    out = []
    for j, p in enumerate(q):
        row = []
        for k, m in enumerate(plain):
            fix = (m * (Q % t) + ((t + 1) >> 1)) // t
            row.append((c0[j][k] + m * (Q // t % p) + fix) % p)
        out.append(row)
    return out


def relinearize(ct3, keys, primes, L, ntt_fwd, ntt_inv):
    """A.7.  keys[J][c][I] = list of N NTT-form residues at key level; ntt_fwd/ntt_inv(poly, prime_index)."""
    K = len(primes)
    P = primes[-1]
    half = P >> 1
    n = len(ct3[0][0])
    target = ct3[2]
    out = []
    for c in range(2):
        accs = {}
        for I in list(range(L)) + [K - 1]:
            qI = primes[I]
            acc = [0] * n
            for J in range(L):
                d = ntt_fwd([v % qI for v in target[J]], I)
                acc = [(x + y * z) % qI for x, y, z in zip(acc, d, keys[J][c][I])]
            accs[I] = ntt_inv(acc, I)
        u = [(v + half) % P for v in accs[K - 1]]
        comp = []
        for i in range(L):
            qi = primes[i]
            inv = pow(P, -1, qi)
            comp.append([(ct3[c][i][k] + (accs[i][k] - (u[k] % qi - half % qi)) * inv) % qi for k in range(n)])
        out.append(comp)
    return out


def crt_decrypt(ct, secret, q, t):
    """exact decryption with big integers: ct = list of polys (each list per prime of coefficient lists) in
    coefficient form; secret = list of {-1,0,1}.  Returns (plaintext coefficients, noise budget bits)."""
    Q = prod(q)
    n = len(secret)
    # reconstruct each polynomial's coefficients mod Q
    def crt(poly):
        out = []
        for k in range(n):
            v = 0
            for j, p in enumerate(q):
                Mj = Q // p
                v += poly[j][k] * Mj * pow(Mj, -1, p)
            out.append(v % Q)
        return out
    s = [x % Q for x in secret]
    acc = crt(ct[0])
    spow = s
    for poly in ct[1:]:
        acc = [(x + y) % Q for x, y in zip(acc, negacyclic_mul(crt(poly), spow, Q))]
        spow = negacyclic_mul(spow, s, Q)
    plain, worst = [], 0
    for v in acc:
        num = v * t
        m = (num + Q // 2) // Q
        err = abs(num - m * Q)  # |t*v - m*Q| < Q/2 required
        worst = max(worst, err)
        plain.append(m % t)
    budget = (Q // 2).bit_length() - worst.bit_length() if worst else Q.bit_length()
    return plain, budget
