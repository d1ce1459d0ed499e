// Row f2: the byte formats that surround the accelerated path on the wire, on the HOST, without SEAL
// and without the FlatBuffers library, so that libapsu_b200 can sit directly behind the reference's ZeroMQ flow:
//
//   in:  ReceiverOperation{QueryRequest}   common/apsu/network/rop.fbs, receiver_operation.cpp:187-345
//          relin_keys : seal::Serializable<RelinKeys>::save bytes
//          query[i]   : exponent + one seal::Serializable<Ciphertext>::save blob per bundle index (seeded: c1 is a PRNG
//                       seed; common/apsu/seal_object.h:161-219 expands it on the host in the reference — here the seed
//                       is handed to apsu_b200_query_begin_seeded and expanded on the device)
//   out: ResultPackage                      common/apsu/network/result_package.fbs, result_package.cpp:29-76
//          psu_result : seal::Ciphertext::save bytes of the result ciphertext (last level, size 2)
//
// [SEAL-RECALL — SEAL 3.7 native/src/seal/{serialization.h,ciphertext.cpp,dynarray.h,kswitchkeys.h,publickey.h,
// randomgen.cpp,encryptionparams.cpp}; NOT verifiable in the build container, SEAL is absent.  tools/seal_kat/ is the
// place to pin these against a SEAL install.]
//   SEALHeader (16 bytes, little endian): u16 magic 0xA15E | u8 header_size 0x10 | u8 version_major | u8 version_minor |
//                                         u8 compr_mode (0 none, 1 zlib, 2 zstd) | u16 reserved | u64 size (total, with header)
//   Ciphertext members: parms_id u64[4] | is_ntt_form u8 | size u64 | poly_modulus_degree u64 | coeff_modulus_size u64 |
//                       scale f64 | correction_factor u64 | DynArray (own SEALHeader | u64 count | count words)
//                       seeded form: the DynArray holds HALF the words (c0 only) and is followed by a
//                       UniformRandomGeneratorInfo (own SEALHeader | u8 prng_type (1 = blake2xb) | 64 seed bytes)
//   KSwitchKeys members: parms_id u64[4] | keys_dim1 u64 | per row: keys_dim2 u64 | per key: PublicKey (own SEALHeader |
//                       Ciphertext members)
//   parms_id = BLAKE2b-256 over the u64 words [scheme (bfv = 1), poly_modulus_degree, coeff_modulus..., plain_modulus]
// Compression (SEAL's default when built with zstd — the vcpkg port APSU uses — is compr_mode_type::zstd): an object saved
// compressed is its SEALHeader followed by ONE zstd frame / zlib stream of the uncompressed members (nested objects inside
// are saved uncompressed).  The decoders are bound at run time (dlopen of libzstd.so.1 / libz.so.1), so this header has no
// link-time dependency beyond -ldl; without the library a compressed object raises std::runtime_error.
// FlatBuffers: a read-only table walker and a writer for the one ResultPackage layout, following the FlatBuffers
// binary format (tables with vtables, little-endian uoffset32).
#pragma once
#include <dlfcn.h>
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace apsu {
namespace wire {

using bytes = std::vector<std::uint8_t>;

// ---------------------------------------------------------------------------------------- little-endian cursor
class Reader {
public:
    Reader(const std::uint8_t *p, std::size_t n) : p_(p), n_(n) {}
    std::size_t left() const { return n_ - at_; }
    std::size_t pos() const { return at_; }
    const std::uint8_t *here() const { return p_ + at_; }
    void skip(std::size_t k)
    {
        need(k);
        at_ += k;
    }
    std::uint8_t u8()
    {
        need(1);
        return p_[at_++];
    }
    std::uint16_t u16() { return (std::uint16_t)le(2); }
    std::uint32_t u32() { return (std::uint32_t)le(4); }
    std::uint64_t u64() { return le(8); }
    void words(std::uint64_t *dst, std::size_t count)
    {
        need(count * 8);
        for (std::size_t i = 0; i < count; i++) {
            std::uint64_t v = 0;
            for (int b = 7; b >= 0; b--) v = (v << 8) | p_[at_ + 8 * i + (std::size_t)b];
            dst[i] = v;
        }
        at_ += count * 8;
    }

private:
    void need(std::size_t k) const
    {
        if (k > n_ - at_) throw std::runtime_error("seal_wire: buffer is too short");
    }
    std::uint64_t le(int k)
    {
        need((std::size_t)k);
        std::uint64_t v = 0;
        for (int b = k - 1; b >= 0; b--) v = (v << 8) | p_[at_ + (std::size_t)b];
        at_ += (std::size_t)k;
        return v;
    }
    const std::uint8_t *p_;
    std::size_t n_, at_ = 0;
};
inline void put(bytes &o, std::uint64_t v, int k)
{
    for (int b = 0; b < k; b++) o.push_back((std::uint8_t)(v >> (8 * b)));
}

// ---------------------------------------------------------------------------------------- BLAKE2b-256 (RFC 7693), for parms_id
inline std::array<std::uint64_t, 4> blake2b_256(const std::uint8_t *in, std::size_t inlen)
{
    static const std::uint64_t iv[8] = { 0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                         0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL };
    static const std::uint8_t sg[10][16] = {
        { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 }, { 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3 },
        { 11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4 }, { 7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8 },
        { 9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13 }, { 2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9 },
        { 12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11 }, { 13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10 },
        { 6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5 }, { 10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0 } };
    std::uint64_t h[8];
    for (int i = 0; i < 8; i++) h[i] = iv[i];
    h[0] ^= 0x01010000ULL ^ 32; // digest 32 bytes, no key, fanout 1, depth 1
    auto rotr = [](std::uint64_t x, int n) { return (x >> n) | (x << (64 - n)); };
    auto compress = [&](const std::uint8_t *blk, std::uint64_t t, bool last) {
        std::uint64_t m[16], v[16];
        for (int i = 0; i < 16; i++) {
            m[i] = 0;
            for (int b = 7; b >= 0; b--) m[i] = (m[i] << 8) | blk[8 * i + b];
        }
        for (int i = 0; i < 8; i++) v[i] = h[i], v[8 + i] = iv[i];
        v[12] ^= t;
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, std::uint64_t x, std::uint64_t y) {
            v[a] = v[a] + v[b] + x, v[d] = rotr(v[d] ^ v[a], 32), v[c] = v[c] + v[d], v[b] = rotr(v[b] ^ v[c], 24);
            v[a] = v[a] + v[b] + y, v[d] = rotr(v[d] ^ v[a], 16), v[c] = v[c] + v[d], v[b] = rotr(v[b] ^ v[c], 63);
        };
        for (int r = 0; r < 12; r++) {
            const std::uint8_t *s = sg[r % 10];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]), G(1, 5, 9, 13, m[s[2]], m[s[3]]), G(2, 6, 10, 14, m[s[4]], m[s[5]]), G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]), G(1, 6, 11, 12, m[s[10]], m[s[11]]), G(2, 7, 8, 13, m[s[12]], m[s[13]]), G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    };
    std::uint8_t blk[128];
    std::size_t off = 0;
    std::uint64_t t = 0;
    while (inlen - off > 128) {
        t += 128;
        compress(in + off, t, false);
        off += 128;
    }
    std::memset(blk, 0, 128);
    std::memcpy(blk, in + off, inlen - off);
    t += inlen - off;
    compress(blk, t, true);
    return { h[0], h[1], h[2], h[3] };
}

// parms_id of one level: EncryptionParameters::compute_parms_id (BFV)
inline std::array<std::uint64_t, 4> parms_id(std::uint64_t poly_modulus_degree, const std::vector<std::uint64_t> &coeff_modulus, std::uint64_t plain_modulus)
{
    bytes b;
    put(b, 1, 8); // scheme_type::bfv
    put(b, poly_modulus_degree, 8);
    for (auto q : coeff_modulus) put(b, q, 8);
    put(b, plain_modulus, 8);
    return blake2b_256(b.data(), b.size());
}

// ---------------------------------------------------------------------------------------- SEAL objects
struct SEALHeader {
    std::uint8_t version_major = 3, version_minor = 7, compr_mode = 0;
    std::uint64_t size = 0;
};
// header only; compr_mode is returned, not judged (nested objects must be uncompressed: see read_plain_header)
inline SEALHeader read_any_header(Reader &r)
{
    if (r.u16() != 0xA15E) throw std::runtime_error("seal_wire: bad SEALHeader magic");
    if (r.u8() != 0x10) throw std::runtime_error("seal_wire: bad SEALHeader size");
    SEALHeader h;
    h.version_major = r.u8();
    h.version_minor = r.u8();
    h.compr_mode = r.u8();
    r.u16();
    h.size = r.u64();
    if (h.compr_mode > 2) throw std::runtime_error("seal_wire: unknown compr_mode");
    return h;
}
inline SEALHeader read_header(Reader &r)
{
    SEALHeader h = read_any_header(r);
    if (h.compr_mode != 0) throw std::runtime_error("seal_wire: nested SEAL object is compressed");
    return h;
}

// zstd frame / zlib stream -> bytes, with the system libraries bound at run time
inline bytes inflate(std::uint8_t compr_mode, const std::uint8_t *p, std::size_t n)
{
    if (compr_mode == 2) {
        static void *so = dlopen("libzstd.so.1", RTLD_NOW);
        if (!so) throw std::runtime_error("seal_wire: zstd-compressed SEAL object and libzstd.so.1 is not available");
        using init_t = void *(*)();
        using free_t = std::size_t (*)(void *);
        struct Buf {
            const void *src;
            std::size_t size, pos;
        };
        struct OBuf {
            void *dst;
            std::size_t size, pos;
        };
        using step_t = std::size_t (*)(void *, OBuf *, Buf *);
        using iserr_t = unsigned (*)(std::size_t);
        static auto create = reinterpret_cast<init_t>(dlsym(so, "ZSTD_createDStream"));
        static auto destroy = reinterpret_cast<free_t>(dlsym(so, "ZSTD_freeDStream"));
        static auto step = reinterpret_cast<step_t>(dlsym(so, "ZSTD_decompressStream"));
        static auto iserr = reinterpret_cast<iserr_t>(dlsym(so, "ZSTD_isError"));
        if (!create || !destroy || !step || !iserr) throw std::runtime_error("seal_wire: libzstd lacks the streaming API");
        void *ds = create();
        if (!ds) throw std::runtime_error("seal_wire: ZSTD_createDStream failed");
        bytes out;
        Buf in{ p, n, 0 };
        std::vector<std::uint8_t> chunk(1 << 20);
        for (;;) {
            OBuf ob{ chunk.data(), chunk.size(), 0 };
            const std::size_t rc = step(ds, &ob, &in);
            if (iserr(rc)) {
                destroy(ds);
                throw std::runtime_error("seal_wire: zstd stream is corrupt");
            }
            out.insert(out.end(), chunk.begin(), chunk.begin() + (std::ptrdiff_t)ob.pos);
            if (rc == 0 && in.pos == in.size) break;                 // frame complete, input consumed
            if (in.pos == in.size && ob.pos == 0) {                  // input exhausted inside a frame
                destroy(ds);
                throw std::runtime_error("seal_wire: zstd stream is truncated");
            }
        }
        destroy(ds);
        return out;
    }
    if (compr_mode == 1) {
        static void *so = dlopen("libz.so.1", RTLD_NOW);
        if (!so) throw std::runtime_error("seal_wire: zlib-compressed SEAL object and libz.so.1 is not available");
        using unc_t = int (*)(unsigned char *, unsigned long *, const unsigned char *, unsigned long);
        static auto unc = reinterpret_cast<unc_t>(dlsym(so, "uncompress"));
        if (!unc) throw std::runtime_error("seal_wire: libz lacks uncompress");
        for (std::size_t cap = n * 4 + 4096;; cap *= 2) { // Z_BUF_ERROR (-5): output too small
            bytes out(cap);
            unsigned long len = (unsigned long)cap;
            const int rc = unc(out.data(), &len, p, (unsigned long)n);
            if (rc == 0) {
                out.resize(len);
                return out;
            }
            if (rc != -5 || cap > (std::size_t(1) << 33)) throw std::runtime_error("seal_wire: zlib stream is corrupt");
        }
    }
    return bytes(p, p + n);
}
// Opens a top-level SEAL object: consumes its header and hands back the uncompressed members (a copy when it was compressed)
struct OpenedObject {
    SEALHeader header;
    bytes storage;              // holds the inflated members when the object was compressed
    const std::uint8_t *p = nullptr;
    std::size_t n = 0;
};
inline OpenedObject open_object(const std::uint8_t *buf, std::size_t len)
{
    Reader r(buf, len);
    OpenedObject o;
    o.header = read_any_header(r);
    if (o.header.size > len || o.header.size < 16) throw std::runtime_error("seal_wire: SEAL object is truncated");
    if (o.header.compr_mode) {
        o.storage = inflate(o.header.compr_mode, buf + 16, (std::size_t)o.header.size - 16);
        o.p = o.storage.data();
        o.n = o.storage.size();
    } else {
        o.p = buf + 16;
        o.n = (std::size_t)o.header.size - 16;
    }
    return o;
}
inline void write_header(bytes &o, std::uint64_t total_size, std::uint8_t vmaj = 3, std::uint8_t vmin = 7)
{
    put(o, 0xA15E, 2);
    put(o, 0x10, 1);
    put(o, vmaj, 1);
    put(o, vmin, 1);
    put(o, 0, 1);
    put(o, 0, 2);
    put(o, total_size, 8);
}

struct Ciphertext {
    std::array<std::uint64_t, 4> parms_id{};
    bool is_ntt_form = false;
    std::uint64_t size = 0, poly_modulus_degree = 0, coeff_modulus_size = 0, correction_factor = 1;
    double scale = 1.0;
    std::vector<std::uint64_t> data; // [size][L][N], or [1][L][N] (c0 only) when seeded
    bool seeded = false;
    std::array<std::uint8_t, 64> seed{};
};

// Ciphertext members, after the object's own SEALHeader has been consumed
inline Ciphertext read_ciphertext_members(Reader &r)
{
    Ciphertext c;
    r.words(c.parms_id.data(), 4);
    c.is_ntt_form = r.u8() != 0;
    c.size = r.u64();
    c.poly_modulus_degree = r.u64();
    c.coeff_modulus_size = r.u64();
    std::uint64_t sc = r.u64();
    std::memcpy(&c.scale, &sc, 8);
    c.correction_factor = r.u64();
    if (c.size > 16 || c.poly_modulus_degree > (1u << 17) || c.coeff_modulus_size > 64) throw std::runtime_error("seal_wire: implausible ciphertext metadata");
    const std::uint64_t total = c.size * c.poly_modulus_degree * c.coeff_modulus_size;
    read_header(r); // DynArray
    const std::uint64_t count = r.u64();
    if (count != total && !(c.size == 2 && count == total / 2)) throw std::runtime_error("seal_wire: ciphertext data has the wrong length");
    c.data.resize(count);
    r.words(c.data.data(), count);
    if (count != total) { // seeded: UniformRandomGeneratorInfo follows
        read_header(r);
        if (r.u8() != 1) throw std::runtime_error("seal_wire: seeded ciphertext uses a generator other than blake2xb");
        for (auto &b : c.seed) b = r.u8();
        c.seeded = true;
    }
    return c;
}
inline Ciphertext read_ciphertext(const std::uint8_t *p, std::size_t n, std::size_t *consumed = nullptr)
{
    OpenedObject o = open_object(p, n);
    Reader r(o.p, o.n);
    Ciphertext c = read_ciphertext_members(r);
    if (consumed) *consumed = (std::size_t)o.header.size;
    return c;
}
// Ciphertext::save (compr_mode none) of a fully expanded ciphertext
inline bytes write_ciphertext(const Ciphertext &c)
{
    bytes body;
    for (auto w : c.parms_id) put(body, w, 8);
    put(body, c.is_ntt_form ? 1 : 0, 1);
    put(body, c.size, 8);
    put(body, c.poly_modulus_degree, 8);
    put(body, c.coeff_modulus_size, 8);
    std::uint64_t sc;
    std::memcpy(&sc, &c.scale, 8);
    put(body, sc, 8);
    put(body, c.correction_factor, 8);
    write_header(body, 16 + 8 + c.data.size() * 8);
    put(body, c.data.size(), 8);
    for (auto w : c.data) put(body, w, 8);
    bytes o;
    write_header(o, 16 + body.size());
    o.insert(o.end(), body.begin(), body.end());
    return o;
}

// RelinKeys = KSwitchKeys with one row (key power 2) of K-1 keys; every key a PublicKey wrapping a size-2 ciphertext at
// the key level in NTT form, seeded when it came out of KeyGenerator::create_relin_keys() as a Serializable
struct RelinKeys {
    std::array<std::uint64_t, 4> parms_id{};
    std::vector<Ciphertext> keys;
};
inline RelinKeys read_relin_keys(const std::uint8_t *p, std::size_t n)
{
    OpenedObject o = open_object(p, n);
    Reader r(o.p, o.n);
    RelinKeys k;
    r.words(k.parms_id.data(), 4);
    const std::uint64_t dim1 = r.u64();
    if (dim1 != 1) throw std::runtime_error("seal_wire: relinearisation keys for ciphertexts of size > 3 are not used by APSU");
    const std::uint64_t dim2 = r.u64();
    if (dim2 > 64) throw std::runtime_error("seal_wire: implausible key count");
    for (std::uint64_t j = 0; j < dim2; j++) {
        read_header(r); // PublicKey
        k.keys.push_back(read_ciphertext_members(r));
    }
    return k;
}

// ---------------------------------------------------------------------------------------- FlatBuffers (read-only walker)
class FbTable {
public:
    FbTable() = default;
    FbTable(const std::uint8_t *buf, std::size_t n, std::size_t pos) : b_(buf), n_(n), pos_(pos)
    {
        const std::int32_t so = (std::int32_t)rd32(pos);
        vt_ = (std::size_t)((std::int64_t)pos - so);
        check(vt_ + 4 <= n);
        vt_len_ = rd16(vt_);
        check(vt_ + vt_len_ <= n && vt_len_ >= 4);
    }
    explicit operator bool() const { return b_ != nullptr; }
    bool has(int field) const { return off(field) != 0; }
    std::uint32_t u32(int field, std::uint32_t dflt = 0) const { return off(field) ? rd32(pos_ + off(field)) : dflt; }
    std::uint8_t u8(int field, std::uint8_t dflt = 0) const
    {
        if (!off(field)) return dflt;
        check(pos_ + off(field) < n_);
        return b_[pos_ + off(field)];
    }
    FbTable table(int field) const
    {
        if (!off(field)) return FbTable();
        const std::size_t at = pos_ + off(field);
        return FbTable(b_, n_, at + rd32(at));
    }
    // vector of ubyte: (pointer, length)
    std::pair<const std::uint8_t *, std::size_t> byte_vector(int field) const
    {
        if (!off(field)) return { nullptr, 0 };
        const std::size_t at = pos_ + off(field), v = at + rd32(at);
        const std::size_t len = rd32(v);
        check(v + 4 + len <= n_);
        return { b_ + v + 4, len };
    }
    std::size_t vector_size(int field) const
    {
        if (!off(field)) return 0;
        const std::size_t at = pos_ + off(field);
        return rd32(at + rd32(at));
    }
    FbTable vector_table(int field, std::size_t i) const
    {
        const std::size_t at = pos_ + off(field), v = at + rd32(at);
        check(i < rd32(v));
        const std::size_t e = v + 4 + 4 * i;
        return FbTable(b_, n_, e + rd32(e));
    }

private:
    void check(bool ok) const
    {
        if (!ok) throw std::runtime_error("seal_wire: invalid flatbuffer");
    }
    std::uint16_t rd16(std::size_t at) const
    {
        check(at + 2 <= n_);
        return (std::uint16_t)(b_[at] | (b_[at + 1] << 8));
    }
    std::uint32_t rd32(std::size_t at) const
    {
        check(at + 4 <= n_);
        return (std::uint32_t)b_[at] | ((std::uint32_t)b_[at + 1] << 8) | ((std::uint32_t)b_[at + 2] << 16) | ((std::uint32_t)b_[at + 3] << 24);
    }
    std::uint16_t off(int field) const
    {
        const std::size_t slot = 4 + 2 * (std::size_t)field;
        return slot + 2 <= vt_len_ ? rd16(vt_ + slot) : 0;
    }
    const std::uint8_t *b_ = nullptr;
    std::size_t n_ = 0, pos_ = 0, vt_ = 0;
    std::uint16_t vt_len_ = 0;
};
inline FbTable fb_size_prefixed_root(const std::uint8_t *buf, std::size_t n)
{
    if (n < 8) throw std::runtime_error("seal_wire: invalid flatbuffer");
    Reader r(buf, n);
    const std::uint32_t size = r.u32();
    if ((std::size_t)size + 4 > n) throw std::runtime_error("seal_wire: flatbuffer is truncated");
    const std::uint32_t root = r.u32();
    return FbTable(buf + 4, size, root);
}

// One query ciphertext as the device wants it: c0 words + seed (apsu_b200_query_begin_seeded), or both polynomials
struct QueryRequest {
    std::uint8_t compression_type = 0;
    RelinKeys relin_keys;
    bool has_relin_keys = false;
    std::vector<std::pair<std::uint32_t, std::vector<Ciphertext>>> parts; // (exponent, one ciphertext per bundle index)
};
// ReceiverOperation{request_type = QueryRequest (3), request} — rop.fbs: union Request { ParmsRequest = 1, OPRFRequest = 2,
// QueryRequest = 3, plainResponse = 4 }
inline QueryRequest read_query_request(const std::uint8_t *buf, std::size_t n)
{
    FbTable rop = fb_size_prefixed_root(buf, n);
    if (rop.u8(0) != 3) throw std::runtime_error("unexpected operation type"); // receiver_operation.cpp:265-267
    FbTable req = rop.table(1);
    if (!req) throw std::runtime_error("failed to load ReceiverOperation: invalid buffer");
    QueryRequest q;
    q.compression_type = req.u8(0);
    auto rk = req.byte_vector(1);
    if (rk.first) {
        q.relin_keys = read_relin_keys(rk.first, rk.second);
        q.has_relin_keys = true;
    }
    if (!req.has(2)) throw std::runtime_error("failed to load ReceiverOperation: invalid buffer");
    for (std::size_t i = 0; i < req.vector_size(2); i++) {
        FbTable part = req.vector_table(2, i);
        std::vector<Ciphertext> cts;
        for (std::size_t k = 0; k < part.vector_size(1); k++) {
            auto d = part.vector_table(1, k).byte_vector(0);
            if (!d.first) throw std::runtime_error("failed to load query ciphertext: missing data");
            cts.push_back(read_ciphertext(d.first, d.second));
        }
        for (auto &p : q.parts)
            if (p.first == part.u32(0)) throw std::runtime_error("invalid query data"); // receiver_operation.cpp:314-316
        q.parts.emplace_back(part.u32(0), std::move(cts));
    }
    return q;
}

// ---------------------------------------------------------------------------------------- FlatBuffers (writers)
// Minimal front-to-back writer: every offset is a forward uoffset32 from its own position, vtables precede their
// tables (soffset = table - vtable > 0), everything 4-byte aligned — a valid FlatBuffers buffer, just not byte-identical
// to what FlatBufferBuilder (which builds back to front) would emit.
class FbWriter {
public:
    std::size_t size() const { return o_.size(); }
    void align4()
    {
        while (o_.size() % 4) o_.push_back(0);
    }
    std::size_t u32(std::uint32_t v)
    {
        std::size_t at = o_.size();
        put(o_, v, 4);
        return at;
    }
    void patch_offset(std::size_t slot, std::size_t target)
    {
        const std::uint32_t v = (std::uint32_t)(target - slot);
        for (int b = 0; b < 4; b++) o_[slot + (std::size_t)b] = (std::uint8_t)(v >> (8 * b));
    }
    // table whose fields are all 4 bytes wide: values[i] is written verbatim; fields listed in `absent` get a zero
    // vtable entry (the reader then sees them as not present); returns (table position, slot positions)
    std::pair<std::size_t, std::vector<std::size_t>> table(const std::vector<std::uint32_t> &values, const std::vector<int> &absent = {})
    {
        align4();
        const std::size_t vt = o_.size(), nf = values.size();
        put(o_, 4 + 2 * nf, 2);
        put(o_, 4 + 4 * nf, 2);
        for (std::size_t i = 0; i < nf; i++) {
            bool gone = false;
            for (int a : absent) gone |= (std::size_t)a == i;
            put(o_, gone ? 0 : 4 + 4 * i, 2);
        }
        align4();
        const std::size_t tab = o_.size();
        put(o_, (std::uint32_t)(tab - vt), 4);
        std::vector<std::size_t> slots;
        for (auto v : values) slots.push_back(u32(v));
        return { tab, slots };
    }
    std::size_t byte_vector(const std::uint8_t *p, std::size_t n)
    {
        align4();
        const std::size_t at = u32((std::uint32_t)n);
        o_.insert(o_.end(), p, p + n);
        return at;
    }
    std::size_t offset_vector(std::size_t n, std::vector<std::size_t> &slots)
    {
        align4();
        const std::size_t at = u32((std::uint32_t)n);
        for (std::size_t i = 0; i < n; i++) slots.push_back(u32(0));
        return at;
    }
    bytes finish_size_prefixed(std::size_t root_table)
    {
        // the first 8 bytes were reserved by begin(): size prefix + root uoffset (relative to its own position, byte 4)
        const std::uint32_t total = (std::uint32_t)(o_.size() - 4);
        for (int b = 0; b < 4; b++) o_[(std::size_t)b] = (std::uint8_t)(total >> (8 * b));
        patch_offset(4, root_table);
        return o_;
    }
    void begin()
    {
        o_.clear();
        u32(0);
        u32(0);
    }

private:
    bytes o_;
};

// ResultPackage::save (result_package.cpp:29-76): bundle_idx, cache_idx, psu_result{data}, label_byte_count,
// nonce_byte_count, label_result (empty) — size-prefixed
inline bytes write_result_package(std::uint32_t bundle_idx, std::uint32_t cache_idx, const bytes &psu_ciphertext, std::uint32_t label_byte_count = 0,
                                  std::uint32_t nonce_byte_count = 0)
{
    FbWriter w;
    w.begin();
    auto rp = w.table({ bundle_idx, cache_idx, 0 /*psu_result*/, label_byte_count, nonce_byte_count, 0 /*label_result*/ });
    auto ct = w.table({ 0 /*data*/ });
    w.patch_offset(rp.second[2], ct.first);
    const std::size_t data = w.byte_vector(psu_ciphertext.data(), psu_ciphertext.size());
    w.patch_offset(ct.second[0], data);
    std::vector<std::size_t> none;
    const std::size_t labels = w.offset_vector(0, none);
    w.patch_offset(rp.second[5], labels);
    return w.finish_size_prefixed(rp.first);
}
struct ResultPackage {
    std::uint32_t bundle_idx = 0, cache_idx = 0, label_byte_count = 0, nonce_byte_count = 0;
    Ciphertext psu_result;
};
inline ResultPackage read_result_package(const std::uint8_t *buf, std::size_t n)
{
    FbTable t = fb_size_prefixed_root(buf, n);
    ResultPackage rp;
    rp.bundle_idx = t.u32(0);
    rp.cache_idx = t.u32(1);
    rp.label_byte_count = t.u32(3);
    rp.nonce_byte_count = t.u32(4);
    auto d = t.table(2).byte_vector(0);
    if (!d.first) throw std::runtime_error("failed to load ResultPackage: invalid buffer");
    rp.psu_result = read_ciphertext(d.first, d.second);
    return rp;
}

// Test helper: a ReceiverOperation{QueryRequest} buffer in the same layout (what the sender's
// ReceiverOperationQuery::save emits, receiver_operation.cpp:187-245), from already serialised SEAL blobs
inline bytes write_query_request(std::uint8_t compression_type, const bytes &relin_keys, const std::vector<std::pair<std::uint32_t, std::vector<bytes>>> &parts)
{
    FbWriter w;
    w.begin();
    auto rop = w.table({ 3 /*request_type = QueryRequest (ubyte in a 4-byte slot)*/, 0 /*request*/ });
    auto req = w.table({ compression_type, 0 /*relin_keys*/, 0 /*query*/ }, relin_keys.empty() ? std::vector<int>{ 1 } : std::vector<int>{});
    w.patch_offset(rop.second[1], req.first);
    if (!relin_keys.empty()) w.patch_offset(req.second[1], w.byte_vector(relin_keys.data(), relin_keys.size()));
    std::vector<std::size_t> part_slots;
    w.patch_offset(req.second[2], w.offset_vector(parts.size(), part_slots));
    for (std::size_t i = 0; i < parts.size(); i++) {
        auto part = w.table({ parts[i].first, 0 /*cts*/ });
        w.patch_offset(part_slots[i], part.first);
        std::vector<std::size_t> ct_slots;
        w.patch_offset(part.second[1], w.offset_vector(parts[i].second.size(), ct_slots));
        for (std::size_t k = 0; k < parts[i].second.size(); k++) {
            auto ct = w.table({ 0 });
            w.patch_offset(ct_slots[k], ct.first);
            w.patch_offset(ct.second[0], w.byte_vector(parts[i].second[k].data(), parts[i].second[k].size()));
        }
    }
    return w.finish_size_prefixed(rop.first);
}

} // namespace wire
} // namespace apsu
