// seal_kat — known-answer generator against the REAL Microsoft SEAL 3.7 (the version APSU pins,
// cmake/APSUConfig.cmake.in:44).  Prints the JSON that tests/golden/seal_kat_expected.json holds for the CPU oracle of
// apsu_b200 (oracle/), so that one run on a machine with SEAL pins — or refutes — the oracle's restatement of SEAL:
//
//     cmake -S tools/seal_kat -B build_kat -DCMAKE_PREFIX_PATH=<SEAL 3.7 install> && cmake --build build_kat
//     ./build_kat/seal_kat <repo>/tests/golden/parameters.json > seal_kat_out.json
//     python tools/seal_kat/compare.py seal_kat_out.json
//
// NOT compiled in the apsu_b200 build container (SEAL is absent there): written against the SEAL 3.7 public API and
// the internal headers it installs (seal/util/rlwe.h, seal/util/ntt.h, seal/util/rns.h, seal/randomgen.h).
// Operands: tools/seal_kat/kat_spec.py (counter-based splitmix64 streams), written straight into
// Ciphertext::data() / Plaintext::data() / RelinKeys — evaluation parity needs deterministic operands, not valid
// encryptions (build SEAL with SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT=OFF, as APSU itself requires, README.md:17).
#include "seal/seal.h"
#include "seal/randomgen.h"
#include "seal/util/ntt.h"
#include "seal/util/rlwe.h"
#include "seal/util/rns.h"
#include <array>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

using namespace seal;
using u64 = std::uint64_t;

// ---------------------------------------------------------------- sha256 (FIPS 180-4), self-contained
struct Sha256 {
    std::uint32_t h[8] = { 0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19 };
    std::vector<unsigned char> buf;
    u64 total = 0;
    static std::uint32_t rotr(std::uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const unsigned char *p)
    {
        static const std::uint32_t k[64] = {
            0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
            0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
            0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
            0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
            0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
            0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2 };
        std::uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = (std::uint32_t)p[4 * i] << 24 | (std::uint32_t)p[4 * i + 1] << 16 | (std::uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            std::uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        std::uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            std::uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g), t1 = hh + S1 + ch + k[i] + w[i];
            std::uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c), t2 = S0 + mj;
            hh = g, g = f, f = e, e = d + t1, d = c, c = b, b = a, a = t1 + t2;
        }
        h[0] += a, h[1] += b, h[2] += c, h[3] += d, h[4] += e, h[5] += f, h[6] += g, h[7] += hh;
    }
    void update(const void *data, std::size_t n)
    {
        const unsigned char *p = static_cast<const unsigned char *>(data);
        total += n;
        buf.insert(buf.end(), p, p + n);
        std::size_t off = 0;
        for (; off + 64 <= buf.size(); off += 64) block(buf.data() + off);
        buf.erase(buf.begin(), buf.begin() + (std::ptrdiff_t)off);
    }
    std::string hex()
    {
        u64 bits = total * 8;
        unsigned char pad[72] = { 0x80 };
        std::size_t padlen = (buf.size() < 56 ? 56 : 120) - buf.size();
        update(pad, padlen);
        unsigned char len[8];
        for (int i = 0; i < 8; i++) len[i] = (unsigned char)(bits >> (56 - 8 * i));
        update(len, 8);
        static const char *d = "0123456789abcdef";
        std::string s;
        for (int i = 0; i < 8; i++)
            for (int j = 3; j >= 0; j--) {
                unsigned char byte = (unsigned char)(h[i] >> (8 * j));
                s += d[byte >> 4];
                s += d[byte & 15];
            }
        return s;
    }
};
static std::string digest_words(const u64 *p, std::size_t n)
{
    Sha256 s; // little-endian words, as numpy '<u8'
    std::vector<unsigned char> b(n * 8);
    for (std::size_t i = 0; i < n; i++)
        for (int k = 0; k < 8; k++) b[8 * i + k] = (unsigned char)(p[i] >> (8 * k));
    s.update(b.data(), b.size());
    return s.hex();
}

// ---------------------------------------------------------------- operands (tools/seal_kat/kat_spec.py)
static const u64 kSeed = 0x5EA15EA15EA15EA1ULL;
static u64 splitmix64_at(u64 seed, u64 k)
{
    u64 z = seed + (k + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void stream(u64 sid, std::size_t n, u64 q, u64 *out)
{
    const u64 seed = kSeed ^ (sid * 0x9E3779B97F4A7C15ULL);
    for (std::size_t k = 0; k < n; k++) out[k] = (u64)(((unsigned __int128)splitmix64_at(seed, k) * q) >> 64);
}
static void rns_poly(u64 sid, const std::vector<Modulus> &primes, std::size_t count, std::size_t N, u64 *out)
{
    for (std::size_t j = 0; j < count; j++) stream(sid * 16 + j, N, primes[j].value(), out + j * N);
}

// ---------------------------------------------------------------- minimal reader of tests/golden/parameters.json
// (the 36 reference parameter files in one object: {"<name>.json": {...}}); only the seal_params of the KAT configs
struct Cfg {
    std::size_t N = 0;
    std::vector<int> bits;
    u64 plain_modulus = 0;
    int plain_bits = 0;
};
static std::string slice_object(const std::string &s, std::size_t at)
{
    std::size_t i = s.find('{', at), depth = 0, j = i;
    for (; j < s.size(); j++) {
        if (s[j] == '{') depth++;
        if (s[j] == '}' && --depth == 0) break;
    }
    return s.substr(i, j - i + 1);
}
static bool find_number(const std::string &s, const std::string &key, u64 &v)
{
    std::size_t k = s.find("\"" + key + "\"");
    if (k == std::string::npos) return false;
    k = s.find(':', k) + 1;
    v = std::stoull(s.substr(k));
    return true;
}
static Cfg read_cfg(const std::string &all, const std::string &name)
{
    std::size_t at = all.find("\"" + name + ".json\"");
    if (at == std::string::npos) throw std::runtime_error("no such parameter set: " + name);
    std::string obj = slice_object(all, at);
    std::string sp = slice_object(obj, obj.find("\"seal_params\""));
    Cfg c;
    u64 v = 0;
    find_number(sp, "poly_modulus_degree", v);
    c.N = (std::size_t)v;
    if (find_number(sp, "plain_modulus_bits", v)) c.plain_bits = (int)v;
    else if (find_number(sp, "plain_modulus", v)) c.plain_modulus = v;
    std::size_t a = sp.find('[', sp.find("\"coeff_modulus_bits\"")), b = sp.find(']', a);
    std::stringstream ss(sp.substr(a + 1, b - a - 1));
    std::string tok;
    while (std::getline(ss, tok, ',')) c.bits.push_back(std::stoi(tok));
    return c;
}

static std::string hex(u64 v)
{
    std::stringstream ss;
    ss << "\"0x" << std::hex << v << "\"";
    return ss.str();
}

static Ciphertext make_ct(const SEALContext &ctx, parms_id_type id, const std::vector<Modulus> &q, std::size_t L, std::size_t N, u64 sid0)
{
    Ciphertext c(ctx, id, 2);
    c.resize(ctx, id, 2);
    c.is_ntt_form() = false;
    for (std::size_t k = 0; k < 2; k++) rns_poly(sid0 + k, q, L, N, c.data(k));
    return c;
}

static void kat_config(const std::string &name, const Cfg &cfg, std::ostream &o)
{
    const std::size_t N = cfg.N;
    EncryptionParameters parms(scheme_type::bfv);
    parms.set_poly_modulus_degree(N);
    parms.set_coeff_modulus(CoeffModulus::Create(N, cfg.bits));
    parms.set_plain_modulus(cfg.plain_bits ? PlainModulus::Batching(N, cfg.plain_bits) : Modulus(cfg.plain_modulus));
    SEALContext ctx(parms, true, sec_level_type::tc128);
    auto key_data = ctx.key_context_data();
    auto first = ctx.first_context_data();
    const auto &key_primes = key_data->parms().coeff_modulus();
    const std::size_t K = key_primes.size(), L = first->parms().coeff_modulus().size();
    const u64 t = parms.plain_modulus().value();
    Evaluator ev(ctx);
    BatchEncoder enc(ctx);
    std::map<std::string, std::string> d; // key -> JSON value

    {
        std::string s = "[";
        for (std::size_t j = 0; j < K; j++) s += (j ? ", " : "") + hex(key_primes[j].value());
        d["coeff_modulus"] = s + "]";
        s = "[";
        for (std::size_t j = 0; j < K; j++) s += (j ? ", " : "") + std::to_string(key_data->small_ntt_tables()[j].get_root());
        d["ntt_roots"] = s + "]";
    }
    d["N"] = std::to_string(N);
    d["plain_modulus"] = std::to_string(t);
    d["plain_ntt_root"] = std::to_string(first->plain_ntt_tables()->get_root());
    {
        auto rt = first->rns_tool();
        std::string s = "{\"base_B\": [";
        // base_Bsk = B followed by m_sk
        for (std::size_t i = 0; i + 1 < rt->base_Bsk()->size(); i++) s += (i ? ", " : "") + hex((*rt->base_Bsk())[i].value());
        s += "], \"gamma\": " + hex(rt->gamma().value()) + ", \"m_sk\": " + hex(rt->m_sk().value()) + "}";
        d["rns_tool_first_level"] = s;
    }
    const auto &q = first->parms().coeff_modulus();
    Ciphertext a = make_ct(ctx, ctx.first_parms_id(), q, L, N, 1), b = make_ct(ctx, ctx.first_parms_id(), q, L, N, 3);
    {
        std::vector<u64> x(a.data(0), a.data(0) + N);
        util::ntt_negacyclic_harvey(x.data(), key_data->small_ntt_tables()[0]);
        d["ntt_forward_q0"] = "\"" + digest_words(x.data(), N) + "\"";
        std::vector<u64> y(a.data(0), a.data(0) + N);
        util::inverse_ntt_negacyclic_harvey(y.data(), key_data->small_ntt_tables()[0]);
        d["ntt_inverse_q0"] = "\"" + digest_words(y.data(), N) + "\"";
    }
    std::vector<u64> values(N);
    stream(900, N, t, values.data());
    Plaintext plain;
    enc.encode(values, plain);
    {
        std::vector<u64> pc(N, 0); // encode yields exactly N coefficients (trailing zeros may be trimmed by coeff_count)
        std::copy_n(plain.data(), plain.coeff_count(), pc.begin());
        d["batch_encode"] = "\"" + digest_words(pc.data(), N) + "\"";
    }
    {
        Plaintext pn = plain;
        ev.transform_to_ntt_inplace(pn, ctx.first_parms_id());
        d["plain_transform_to_ntt"] = "\"" + digest_words(pn.data(), L * N) + "\"";
        Ciphertext c = a;
        ev.add_plain_inplace(c, plain);
        d["add_plain"] = "\"" + digest_words(c.data(), 2 * L * N) + "\"";
        Ciphertext m;
        ev.multiply_plain(a, plain, m);
        d["multiply_plain_coeff_form"] = "\"" + digest_words(m.data(), 2 * L * N) + "\"";
    }
    Ciphertext prod, sq;
    ev.multiply(a, b, prod);
    ev.square(a, sq);
    d["multiply"] = "\"" + digest_words(prod.data(), 3 * L * N) + "\"";
    d["square"] = "\"" + digest_words(sq.data(), 3 * L * N) + "\"";
    if (K > 1) {
        KeyGenerator keygen(ctx);
        RelinKeys rk;
        keygen.create_relin_keys(rk);
        // overwrite the key material with the KAT streams: rk.data()[0][J] = size-2 ciphertext at key level, NTT form
        for (std::size_t J = 0; J + 1 < K; J++)
            for (std::size_t c = 0; c < 2; c++) rns_poly(100 + 2 * J + c, key_primes, K, N, rk.data()[0][J].data().data(c));
        Ciphertext rel = prod;
        ev.relinearize_inplace(rel, rk);
        d["relinearize"] = "\"" + digest_words(rel.data(), 2 * L * N) + "\"";
        std::size_t lvl = L;
        while (lvl > 1) {
            ev.mod_switch_to_next_inplace(rel);
            lvl--;
            d["mod_switch_to_" + std::to_string(lvl) + "_primes"] = "\"" + digest_words(rel.data(), 2 * lvl * N) + "\"";
        }
    }
    {
        prng_seed_type seed;
        unsigned char sb[64];
        for (int i = 0; i < 64; i++) sb[i] = (unsigned char)((37 * i + 11) & 0xFF);
        std::memcpy(seed.data(), sb, 64);
        {
            Blake2xbPRNGFactory f(seed);
            auto prng = f.create();
            std::vector<unsigned char> buf(10000);
            prng->generate(buf.size(), reinterpret_cast<seal_byte *>(buf.data()));
            Sha256 s;
            s.update(buf.data(), buf.size());
            d["blake2xb_prng_first_10000_bytes"] = "\"" + s.hex() + "\"";
        }
        {
            Blake2xbPRNGFactory f(seed);
            std::vector<u64> p(L * N);
            util::sample_poly_uniform(f.create(), first->parms(), p.data());
            d["sample_poly_uniform_first_level"] = "\"" + digest_words(p.data(), p.size()) + "\"";
        }
        {
            Blake2xbPRNGFactory f(seed);
            std::vector<u64> p(K * N);
            util::sample_poly_uniform(f.create(), key_data->parms(), p.data());
            d["sample_poly_uniform_key_level"] = "\"" + digest_words(p.data(), p.size()) + "\"";
        }
    }
    o << "  \"" << name << "\": {\n";
    std::size_t i = 0;
    for (auto &kv : d) o << "   \"" << kv.first << "\": " << kv.second << (++i < d.size() ? ",\n" : "\n");
    o << "  }";
}

int main(int argc, char **argv)
{
    if (argc < 2) {
        std::cerr << "usage: seal_kat <apsu_b200>/tests/golden/parameters.json\n";
        return 2;
    }
    std::ifstream in(argv[1]);
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string all = ss.str();
    const char *names[] = { "256K-512", "1M-1024-cmp", "16M-4096", "256M-4096" };
    std::cout << "{\n \"seal_version\": \"" << SEAL_VERSION << "\",\n \"configs\": {\n";
    for (std::size_t i = 0; i < 4; i++) {
        kat_config(names[i], read_cfg(all, names[i]), std::cout);
        std::cout << (i + 1 < 4 ? ",\n" : "\n");
    }
    std::cout << " }\n}\n";
    return 0;
}
