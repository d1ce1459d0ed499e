// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points over the restatement, bound from Python with
// ctypes (oracle/oracle.py).  Used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
// legs only; the product library (libapsu_b200.so) never links or loads this.
#include "apsu_restate.hpp"
#include "prng_restate.hpp"
#include <chrono>
#include <memory>
#include <string>

using namespace orc;

namespace {
thread_local std::string g_err;
template <typename F>
int guard(F &&f)
{
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

struct Db {
    const Context *ctx;
    PathParams pp;
    PowersDag pd;
    ReceiverDB db;
};

struct Session { // state of one query: the powers table after ComputePowers + results
    std::vector<CiphertextPowers> all_powers;
    std::vector<QueryResult> results;
    double t_powers_ms = 0, t_eval_ms = 0;
};

Ciphertext wrap_ct(const Context &ctx, const u64 *d, size_t size, size_t L, bool ntt)
{
    Ciphertext c;
    c.resize(ctx.N, size, L);
    c.ntt = ntt;
    std::memcpy(c.d.data(), d, c.d.size() * sizeof(u64));
    return c;
}
} // namespace

extern "C" {

const char *orc_last_error() { return g_err.c_str(); }

// ---- number theory fixtures ----
int orc_get_primes(u64 factor, int bits, size_t count, u64 *out)
{
    return guard([&] {
        auto v = get_primes(factor, bits, count);
        std::copy(v.begin(), v.end(), out);
    });
}
int orc_coeff_modulus_create(size_t N, const int *bits, size_t nbits, u64 *out)
{
    return guard([&] {
        auto v = coeff_modulus_create(N, std::vector<int>(bits, bits + nbits));
        std::copy(v.begin(), v.end(), out);
    });
}
u64 orc_plain_modulus_batching(size_t N, int bits)
{
    u64 r = 0;
    guard([&] { r = plain_modulus_batching(N, bits); });
    return r;
}
u64 orc_minimal_primitive_root(u64 degree, u64 modulus)
{
    u64 r = 0;
    guard([&] { r = minimal_primitive_root(degree, Modulus(modulus)); });
    return r;
}
int orc_ntt_mod(size_t N, u64 modulus, u64 *data, int inverse)
{
    return guard([&] {
        NTTTables t;
        t.init(N, modulus);
        if (inverse)
            t.inverse(data);
        else
            t.forward(data);
    });
}

// ---- context ----
void *orc_ctx_create(size_t N, u64 t, const u64 *primes, size_t K)
{
    Context *c = nullptr;
    int rc = guard([&] {
        c = new Context();
        c->init(N, t, std::vector<u64>(primes, primes + K));
    });
    if (rc) {
        delete c;
        return nullptr;
    }
    return c;
}
void orc_ctx_destroy(void *ctx) { delete (Context *)ctx; }
size_t orc_ctx_first_L(void *ctx) { return ((Context *)ctx)->first_L; }
size_t orc_ctx_level_for_chain_idx(void *ctx, size_t ci) { return ((Context *)ctx)->level_for_chain_idx(ci); }
u64 orc_ctx_root(void *ctx, size_t prime_idx) { return ((Context *)ctx)->ntt[prime_idx].root; }
// aux base of level L: out = [m_sk, gamma, B_0..B_{|B|-1}] ; returns |B|
size_t orc_ctx_aux_base(void *ctx, size_t L, u64 *out)
{
    const Level &lv = ((Context *)ctx)->levels[L];
    out[0] = lv.m_sk.value;
    out[1] = lv.gamma.value;
    for (size_t i = 0; i < lv.B.size(); i++) out[2 + i] = lv.B.base[i].value;
    return lv.B.size();
}

int orc_ntt(void *ctx, size_t prime_idx, u64 *data, int inverse)
{
    return guard([&] {
        const Context &c = *(Context *)ctx;
        if (inverse)
            c.ntt[prime_idx].inverse(data);
        else
            c.ntt[prime_idx].forward(data);
    });
}
int orc_encode(void *ctx, const u64 *values, size_t count, u64 *out)
{
    return guard([&] { batch_encode(*(Context *)ctx, values, count, out); });
}
int orc_decode(void *ctx, const u64 *plain, u64 *values)
{
    return guard([&] { batch_decode(*(Context *)ctx, plain, values); });
}

// ---- evaluator primitives on raw arrays ----
int orc_plain_to_ntt(void *ctx, size_t L, const u64 *plain, u64 *out)
{
    return guard([&] { Evaluator(*(Context *)ctx).plain_to_ntt(plain, L, out); });
}
int orc_multiply(void *ctx, size_t L, const u64 *a, size_t asz, const u64 *b, size_t bsz, u64 *out)
{
    return guard([&] {
        const Context &c = *(Context *)ctx;
        Ciphertext x = wrap_ct(c, a, asz, L, false), y = wrap_ct(c, b, bsz, L, false), r;
        Evaluator(c).multiply(x, y, r);
        std::memcpy(out, r.d.data(), r.d.size() * sizeof(u64));
    });
}
int orc_relinearize(void *ctx, size_t L, const u64 *in3, const u64 *keys, u64 *out2)
{
    return guard([&] {
        const Context &c = *(Context *)ctx;
        Ciphertext x = wrap_ct(c, in3, 3, L, false);
        Evaluator(c).relinearize(x, keys);
        std::memcpy(out2, x.d.data(), x.d.size() * sizeof(u64));
    });
}
int orc_mod_switch_next(void *ctx, size_t L, size_t size, const u64 *in, u64 *out)
{
    return guard([&] {
        const Context &c = *(Context *)ctx;
        Ciphertext x = wrap_ct(c, in, size, L, false);
        Evaluator(c).mod_switch_to_next(x);
        std::memcpy(out, x.d.data(), x.d.size() * sizeof(u64));
    });
}
int orc_add_plain(void *ctx, size_t L, size_t size, u64 *ct, const u64 *plain)
{
    return guard([&] {
        const Context &c = *(Context *)ctx;
        Ciphertext x = wrap_ct(c, ct, size, L, false);
        Evaluator(c).add_plain_inplace(x, plain);
        std::memcpy(ct, x.d.data(), x.d.size() * sizeof(u64));
    });
}
int orc_multiply_plain_normal(void *ctx, size_t L, size_t size, const u64 *ct, const u64 *plain, u64 *out)
{
    return guard([&] {
        const Context &c = *(Context *)ctx;
        Ciphertext x = wrap_ct(c, ct, size, L, false), r;
        Evaluator(c).multiply_plain_normal(x, plain, r);
        std::memcpy(out, r.d.data(), r.d.size() * sizeof(u64));
    });
}

// ---- harness: keys / encrypt / decrypt ----
void *orc_keygen(void *ctx, u64 seed)
{
    KeyMaterial *km = new KeyMaterial();
    if (guard([&] { keygen(*(Context *)ctx, seed, *km); })) {
        delete km;
        return nullptr;
    }
    return km;
}
void orc_km_destroy(void *km) { delete (KeyMaterial *)km; }
size_t orc_km_relin_words(void *km) { return ((KeyMaterial *)km)->relin.size(); }
void orc_km_relin(void *km, u64 *out)
{
    auto &r = ((KeyMaterial *)km)->relin;
    std::copy(r.begin(), r.end(), out);
}
void orc_km_secret(void *km, int8_t *out)
{
    auto &s = ((KeyMaterial *)km)->s;
    std::copy(s.begin(), s.end(), out);
}
int orc_encrypt(void *ctx, void *km, const u64 *plain, u64 seed, u64 *out)
{
    return guard([&] {
        Ciphertext c;
        encrypt_symmetric(*(Context *)ctx, *(KeyMaterial *)km, plain, seed, c);
        std::memcpy(out, c.d.data(), c.d.size() * sizeof(u64));
    });
}
// returns noise budget (bits) or -1000 on error
int orc_decrypt_last(void *ctx, void *km, const u64 *ct, size_t size, u64 *plain_out)
{
    int budget = -1000;
    guard([&] {
        const Context &c = *(Context *)ctx;
        Ciphertext x = wrap_ct(c, ct, size, 1, false);
        budget = decrypt_last_level(c, *(KeyMaterial *)km, x, plain_out);
    });
    return budget;
}

// ---- PowersDag ----
// out arrays sized n_targets: power, depth, p1, p2.  returns number of nodes, or -1
int orc_powers_dag(
    uint32_t ps_low, uint32_t target_degree, const uint32_t *sources, size_t nsrc, uint32_t *power, uint32_t *depth,
    uint32_t *p1, uint32_t *p2)
{
    int n = -1;
    guard([&] {
        PowersDag pd;
        auto tg = create_powers_set(ps_low, target_degree);
        if (!pd.configure(std::set<uint32_t>(sources, sources + nsrc), tg)) throw std::invalid_argument("PowersDag configure failed");
        n = 0;
        for (auto &kv : pd.nodes) {
            power[n] = kv.second.power;
            depth[n] = kv.second.depth;
            p1[n] = kv.second.p1;
            p2[n] = kv.second.p2;
            n++;
        }
    });
    return n;
}

// ---- receiver DB ----
void *orc_db_create(
    void *ctx, uint32_t felts_per_item, uint32_t table_size, uint32_t max_items_per_bin, uint32_t ps_low,
    const uint32_t *query_powers, size_t nqp)
{
    Db *d = nullptr;
    if (guard([&] {
            d = new Db();
            d->ctx = (Context *)ctx;
            d->pp.felts_per_item = felts_per_item;
            d->pp.table_size = table_size;
            d->pp.max_items_per_bin = max_items_per_bin;
            d->pp.ps_low_degree = ps_low;
            d->pp.query_powers = std::set<uint32_t>(query_powers, query_powers + nqp);
            d->pp.query_powers.insert(1);
            d->pp.derive(d->ctx->N);
            if (!d->pd.configure(d->pp.query_powers, create_powers_set(ps_low, max_items_per_bin)))
                throw std::invalid_argument("PowersDag configure failed");
            d->db.bin_bundles.resize(d->pp.bundle_idx_count);
        })) {
        delete d;
        return nullptr;
    }
    return d;
}
void orc_db_destroy(void *db) { delete (Db *)db; }
uint32_t orc_db_bundle_idx_count(void *db) { return ((Db *)db)->pp.bundle_idx_count; }
uint32_t orc_db_bins_per_bundle(void *db) { return ((Db *)db)->pp.bins_per_bundle; }
size_t orc_db_bundle_count(void *db, uint32_t bundle_idx) { return ((Db *)db)->db.bin_bundles[bundle_idx].size(); }

// bins: for each of bins_per_bundle bins, bin_sizes[b] roots taken consecutively from `roots`.
// returns cache_idx or -1
int orc_db_add_bundle_from_bins(void *db, uint32_t bundle_idx, const uint32_t *bin_sizes, const u64 *roots)
{
    int idx = -1;
    guard([&] {
        Db &d = *(Db *)db;
        std::vector<std::vector<u64>> polyns(d.pp.bins_per_bundle);
        size_t off = 0;
        for (uint32_t b = 0; b < d.pp.bins_per_bundle; b++) {
            if (bin_sizes[b] + 1 > d.pp.max_items_per_bin) throw std::invalid_argument("bin exceeds max_items_per_bin - 1 items");
            std::vector<u64> r(roots + off, roots + off + bin_sizes[b]);
            off += bin_sizes[b];
            polyns[b] = polyn_with_roots(r, d.ctx->t);
        }
        BatchedPlaintextPolyn bp;
        bp.build(*d.ctx, polyns, d.pp.ps_low_degree);
        d.db.bin_bundles[bundle_idx].push_back(std::move(bp));
        idx = (int)d.db.bin_bundles[bundle_idx].size() - 1;
    });
    return idx;
}
// synthetic bundle with `ncoeffs` plaintexts whose words are uniform in [0,q_j) / [0,t) (throughput runs,
// SURVEY.md §8d).  Word stream: splitmix64(seed), plaintext-major, prime-major, coefficient-minor,
// value = (next() * modulus) >> 64.
int orc_db_add_bundle_synthetic(void *db, uint32_t bundle_idx, uint32_t ncoeffs, u64 seed)
{
    int idx = -1;
    guard([&] {
        Db &d = *(Db *)db;
        const Context &c = *d.ctx;
        uint32_t ps = d.pp.ps_low_degree;
        size_t plain_L = c.level_for_chain_idx(std::min<size_t>(c.first_L - 1, ps ? 2 : 1));
        SplitMix64 rng(seed);
        BatchedPlaintextPolyn bp;
        for (uint32_t i = 0; i < ncoeffs; i++) {
            Plaintext pt;
            bool to_ntt = (!ps && i != 0) || (ps && (i % (ps + 1)));
            if (to_ntt) {
                pt.L = plain_L;
                pt.d.resize(plain_L * c.N);
                for (size_t j = 0; j < plain_L; j++)
                    for (size_t n = 0; n < c.N; n++) pt.d[j * c.N + n] = rng.below(c.primes[j]);
            } else {
                pt.L = 0;
                pt.d.resize(c.N);
                for (size_t n = 0; n < c.N; n++) pt.d[n] = rng.below(c.t.value);
            }
            bp.batched_coeffs.push_back(std::move(pt));
        }
        d.db.bin_bundles[bundle_idx].push_back(std::move(bp));
        idx = (int)d.db.bin_bundles[bundle_idx].size() - 1;
    });
    return idx;
}
size_t orc_db_bundle_ncoeffs(void *db, uint32_t bundle_idx, uint32_t cache_idx)
{
    return ((Db *)db)->db.bin_bundles[bundle_idx][cache_idx].batched_coeffs.size();
}
// copies plaintext `deg`; returns its level L (0 = coefficient form, N words; else L*N words)
size_t orc_db_bundle_coeff(void *db, uint32_t bundle_idx, uint32_t cache_idx, uint32_t deg, u64 *out)
{
    const Plaintext &p = ((Db *)db)->db.bin_bundles[bundle_idx][cache_idx].batched_coeffs[deg];
    if (out) std::copy(p.d.begin(), p.d.end(), out);
    return p.L;
}

// ---- query ----
// cts: [nsrc][bundle_idx_count][2][first_L][N] coefficient form, src_powers[nsrc]
// masks: dense [alpha_max][bundle_idx_count][N] coefficient-form plaintexts
void *orc_run_query(
    void *db, const uint32_t *src_powers, size_t nsrc, const u64 *cts, const u64 *relin_keys, const u64 *masks,
    size_t threads, int powers_only)
{
    Session *s = nullptr;
    if (guard([&] {
            Db &d = *(Db *)db;
            const Context &c = *d.ctx;
            size_t N = c.N, L = c.first_L;
            uint32_t bic = d.pp.bundle_idx_count;
            {
                std::set<uint32_t> given(src_powers, src_powers + nsrc);
                if (given != d.pp.query_powers) throw std::invalid_argument("query powers do not match the parameters");
            }
            std::vector<std::map<uint32_t, Ciphertext>> query(bic);
            for (size_t k = 0; k < nsrc; k++)
                for (uint32_t b = 0; b < bic; b++)
                    query[b][src_powers[k]] = wrap_ct(c, cts + (k * bic + b) * 2 * L * N, 2, L, false);
            s = new Session();
            if (powers_only) {
                using clk = std::chrono::steady_clock;
                auto t0 = clk::now();
                s->all_powers.resize(bic);
                for (uint32_t b = 0; b < bic; b++) {
                    s->all_powers[b].assign((size_t)d.pp.max_items_per_bin + 1, Ciphertext());
                    for (auto &kv : query[b]) s->all_powers[b][kv.first] = kv.second;
                    if (d.db.bin_bundles[b].empty()) continue;
                    compute_powers(c, d.pd, d.pp.ps_low_degree, relin_keys, s->all_powers[b], threads);
                }
                s->t_powers_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();
            } else {
                s->results = run_query_he(c, d.pp, d.db, d.pd, query, relin_keys, masks, threads, &s->t_powers_ms, &s->t_eval_ms);
            }
        })) {
        delete s;
        return nullptr;
    }
    return s;
}
void orc_session_destroy(void *s) { delete (Session *)s; }
double orc_session_powers_ms(void *s) { return ((Session *)s)->t_powers_ms; }
double orc_session_eval_ms(void *s) { return ((Session *)s)->t_eval_ms; }
size_t orc_session_result_count(void *s) { return ((Session *)s)->results.size(); }
// out: [2][N]
int orc_session_result(void *s, size_t k, uint32_t *bundle_idx, uint32_t *cache_idx, u64 *out)
{
    return guard([&] {
        const QueryResult &r = ((Session *)s)->results.at(k);
        *bundle_idx = r.bundle_idx;
        *cache_idx = r.cache_idx;
        std::copy(r.ct.d.begin(), r.ct.d.end(), out);
    });
}
// returns L of the stored power (0 if absent); *is_ntt set; out may be null to query the shape
size_t orc_session_power(void *s, uint32_t bundle_idx, uint32_t power, int *is_ntt, u64 *out)
{
    Session &se = *(Session *)s;
    if (bundle_idx >= se.all_powers.size() || power >= se.all_powers[bundle_idx].size()) return 0;
    const Ciphertext &c = se.all_powers[bundle_idx][power];
    if (c.empty()) return 0;
    if (is_ntt) *is_ntt = c.ntt ? 1 : 0;
    if (out) std::copy(c.d.begin(), c.d.end(), out);
    return c.L;
}

// evaluate one BinBundle against externally supplied powers is covered by run_query; a timing-only
// helper for the CPU baseline: evaluates the listed (bundle_idx, cache_idx) pairs of a session that
// already holds powers (powers_only=1) and returns elapsed ms.  Results are appended to the session.
double orc_session_eval_subset(
    void *db, void *s, const uint32_t *bundle_idx, const uint32_t *cache_idx, size_t count, const u64 *relin_keys,
    const u64 *masks, size_t threads)
{
    double ms = -1;
    guard([&] {
        Db &d = *(Db *)db;
        Session &se = *(Session *)s;
        const Context &c = *d.ctx;
        uint32_t bic = d.pp.bundle_idx_count;
        size_t base = se.results.size();
        se.results.resize(base + count);
        auto t0 = std::chrono::steady_clock::now();
        parallel_for(count, threads, [&](size_t k) {
            uint32_t b = bundle_idx[k], ci = cache_idx[k];
            const BatchedPlaintextPolyn &bp = d.db.bin_bundles.at(b).at(ci);
            const u64 *mask = masks + ((size_t)b + (size_t)ci * bic) * c.N;
            uint32_t degree = (uint32_t)bp.batched_coeffs.size() - 1;
            bool using_ps = d.pp.ps_low_degree > 1 && d.pp.ps_low_degree < degree;
            QueryResult &r = se.results[base + k];
            r.bundle_idx = b;
            r.cache_idx = ci;
            r.ct = using_ps ? bp_eval_patstock(c, bp, se.all_powers[b], d.pp.ps_low_degree, relin_keys, mask)
                            : bp_eval(c, bp, se.all_powers[b], mask);
        });
        ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    });
    return ms;
}

// ---- SEAL's default random generator and seed expansion (prng_restate.hpp) ----
int orc_blake2b(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen, const uint8_t *key, size_t keylen)
{
    return guard([&] { orc_prng::blake2b(out, outlen, in, inlen, key, keylen); });
}
int orc_blake2xb(uint8_t *out, size_t outlen, const uint8_t *in, size_t inlen, const uint8_t *key, size_t keylen)
{
    return guard([&] { orc_prng::blake2xb(out, outlen, in, inlen, key, keylen); });
}
// first nbytes of the Blake2xbPRNG stream of `seed` (64 bytes)
int orc_prng_bytes(const uint8_t *seed, size_t nbytes, uint8_t *out)
{
    return guard([&] {
        std::array<uint64_t, 8> sd;
        for (int i = 0; i < 8; i++) sd[i] = orc_prng::load64(seed + 8 * i);
        orc_prng::Blake2xbPRNG prng(sd);
        prng.generate(nbytes, out);
    });
}
// sample_poly_uniform with a fresh generator seeded by `seed`: out [L][N]
int orc_sample_poly_uniform(const uint8_t *seed, const u64 *moduli, size_t L, size_t N, u64 *out)
{
    return guard([&] {
        std::array<uint64_t, 8> sd;
        for (int i = 0; i < 8; i++) sd[i] = orc_prng::load64(seed + 8 * i);
        orc_prng::Blake2xbPRNG prng(sd);
        orc_prng::sample_poly_uniform(prng, reinterpret_cast<const uint64_t *>(moduli), L, N, reinterpret_cast<uint64_t *>(out));
    });
}
// mask values of RunQuery (receiver_ddh.cpp:241-262): for every non-padded pack index in ascending order, N draws of
// prng->generate() % t; padded ones are left zero.  values [npack][N]
int orc_mask_values(const uint8_t *seed, const uint8_t *padded, size_t npack, size_t N, u64 t, u64 *values)
{
    return guard([&] {
        std::array<uint64_t, 8> sd;
        for (int i = 0; i < 8; i++) sd[i] = orc_prng::load64(seed + 8 * i);
        orc_prng::Blake2xbPRNG prng(sd);
        for (size_t p = 0; p < npack; p++)
            for (size_t i = 0; i < N; i++) values[p * N + i] = padded[p] ? 0 : (u64)prng.generate() % t;
    });
}

} // extern "C"
