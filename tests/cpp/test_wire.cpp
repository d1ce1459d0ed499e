// Host-side test of apsu_b200/host/seal_wire.hpp (row f2): SEAL object framing, parms_id, and the FlatBuffers envelopes
// of QueryRequest / ResultPackage, round-tripped through the reader and the writer.  Prints `key=value` lines the
// Python test checks (parms_id against hashlib.blake2b).
#include "../../apsu_b200/host/seal_wire.hpp"
#include <cstdio>

using namespace apsu::wire;

static Ciphertext make_ct(std::uint64_t N, std::uint64_t L, std::uint64_t size, std::uint64_t tag)
{
    Ciphertext c;
    c.parms_id = parms_id(N, { 0xfffffffff70001ULL, 0xfffffffff78001ULL, 0xfffffffffb4001ULL }, 4079617);
    c.size = size;
    c.poly_modulus_degree = N;
    c.coeff_modulus_size = L;
    c.data.resize(size * L * N);
    for (std::size_t i = 0; i < c.data.size(); i++) c.data[i] = tag * 1000003ULL + i * 7919ULL;
    return c;
}
// the seeded form Serializable<Ciphertext>::save emits: c0 + UniformRandomGeneratorInfo
static bytes seeded_blob(const Ciphertext &full, const std::array<std::uint8_t, 64> &seed)
{
    Ciphertext half = full;
    half.data.resize(full.data.size() / 2);
    bytes b = write_ciphertext(half);
    // append the generator info and fix the outer size
    bytes info;
    write_header(info, 16 + 1 + 64);
    info.push_back(1);
    info.insert(info.end(), seed.begin(), seed.end());
    b.insert(b.end(), info.begin(), info.end());
    const std::uint64_t total = b.size();
    for (int k = 0; k < 8; k++) b[8 + k] = (std::uint8_t)(total >> (8 * k));
    return b;
}

#include <fstream>
#include <iterator>
static bytes slurp(const char *path)
{
    std::ifstream f(path, std::ios::binary);
    return bytes(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

// `test_wire dump <file>`: writes the uncompressed reference ciphertext object; `test_wire check <file>...`: every file
// must parse (whatever its compr_mode) to that ciphertext.  No arguments: the self-contained round trips.
int main(int argc, char **argv)
{
    const std::uint64_t N = 64, L = 3;
    if (argc >= 3 && std::string(argv[1]) == "dump") {
        bytes b = write_ciphertext(make_ct(N, L, 2, 1));
        std::ofstream(argv[2], std::ios::binary).write(reinterpret_cast<const char *>(b.data()), (std::streamsize)b.size());
        return 0;
    }
    if (argc >= 3 && std::string(argv[1]) == "check") {
        Ciphertext ref = make_ct(N, L, 2, 1);
        for (int i = 2; i < argc; i++) {
            bytes b = slurp(argv[i]);
            Ciphertext c = read_ciphertext(b.data(), b.size());
            if (c.data != ref.data || c.parms_id != ref.parms_id || c.size != 2) {
                std::printf("mismatch in %s\n", argv[i]);
                return 1;
            }
            std::printf("parsed %s compr_mode=%d\n", argv[i], (int)b[5]);
        }
        return 0;
    }
    auto id = parms_id(8192, { 0xfffffffff70001ULL, 0xfffffffff78001ULL, 0xfffffffffb4001ULL, 0x3ffffffffc001ULL }, 4079617);
    std::printf("parms_id=%016llx,%016llx,%016llx,%016llx\n", (unsigned long long)id[0], (unsigned long long)id[1], (unsigned long long)id[2],
                (unsigned long long)id[3]);
    int ok = 1;
    // expanded ciphertext round trip
    Ciphertext a = make_ct(N, L, 2, 1);
    bytes ab = write_ciphertext(a);
    std::size_t used = 0;
    Ciphertext a2 = read_ciphertext(ab.data(), ab.size(), &used);
    ok &= used == ab.size() && a2.data == a.data && a2.parms_id == a.parms_id && !a2.seeded && a2.size == 2 && a2.coeff_modulus_size == L;
    std::printf("ct_bytes=%zu\n", ab.size()); // 16 + 32 + 1 + 5*8 + 16 + 8 + 2*3*64*8
    // seeded ciphertext
    std::array<std::uint8_t, 64> seed;
    for (int i = 0; i < 64; i++) seed[i] = (std::uint8_t)(3 * i + 1);
    bytes sb = seeded_blob(a, seed);
    Ciphertext s2 = read_ciphertext(sb.data(), sb.size());
    ok &= s2.seeded && s2.seed == seed && s2.data.size() == a.data.size() / 2 && std::equal(s2.data.begin(), s2.data.end(), a.data.begin());
    // QueryRequest envelope: two exponents, two bundle indices, seeded ciphertexts, no keys
    std::vector<std::pair<std::uint32_t, std::vector<bytes>>> parts;
    for (std::uint32_t e : { 1u, 5u }) {
        std::vector<bytes> cts;
        for (std::uint64_t b = 0; b < 2; b++) cts.push_back(seeded_blob(make_ct(N, L, 2, 10 * e + b), seed));
        parts.emplace_back(e, cts);
    }
    bytes qb = write_query_request(0, bytes(), parts);
    QueryRequest q = read_query_request(qb.data(), qb.size());
    ok &= q.parts.size() == 2 && q.parts[0].first == 1 && q.parts[1].first == 5 && !q.has_relin_keys;
    for (std::size_t i = 0; i < 2 && ok; i++)
        for (std::size_t b = 0; b < 2; b++) {
            Ciphertext ref = make_ct(N, L, 2, 10 * q.parts[i].first + b);
            const Ciphertext &g = q.parts[i].second[b];
            ok &= g.seeded && g.data.size() == ref.data.size() / 2 && std::equal(g.data.begin(), g.data.end(), ref.data.begin());
        }
    // ResultPackage envelope
    Ciphertext r = make_ct(N, 1, 2, 77);
    bytes rb = write_result_package(3, 9, write_ciphertext(r));
    ResultPackage rp = read_result_package(rb.data(), rb.size());
    ok &= rp.bundle_idx == 3 && rp.cache_idx == 9 && rp.psu_result.data == r.data && rp.label_byte_count == 0;
    // error paths: truncated buffer, wrong magic, compressed object
    try {
        read_ciphertext(ab.data(), ab.size() - 9);
        ok = 0;
    } catch (const std::runtime_error &) {
    }
    bytes bad = ab;
    bad[0] ^= 1;
    try {
        read_ciphertext(bad.data(), bad.size());
        ok = 0;
    } catch (const std::runtime_error &) {
    }
    bad = ab;
    bad[5] = 2; // zstd
    try {
        read_ciphertext(bad.data(), bad.size());
        ok = 0;
    } catch (const std::runtime_error &) {
    }
    std::printf("ok=%d\n", ok);
    return ok ? 0 : 1;
}
