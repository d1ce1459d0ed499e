// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of APSU's receiver-side query evaluation (the hot path of SURVEY.md §8a) on
// top of the SEAL restatement in seal_restate.hpp.  Each function cites the reference lines it
// follows (paths relative to /root/reference).  PARITY UNPINNED — see seal_restate.hpp header.
#pragma once
#include "seal_restate.hpp"
#include <atomic>
#include <map>
#include <set>
#include <thread>
#include <functional>
#include <chrono>

namespace orc {

// ----------------------------------------------------------------------------------------------
// common/apsu/util/utils.cpp:146-177  create_powers_set
// ----------------------------------------------------------------------------------------------
static inline std::set<uint32_t> create_powers_set(uint32_t ps_low_degree, uint32_t target_degree)
{
    if (ps_low_degree > target_degree) throw std::invalid_argument("ps_low_degree cannot be bigger than target_degree");
    if (!target_degree) throw std::invalid_argument("target_degree cannot be zero");
    std::set<uint32_t> r;
    if (ps_low_degree) {
        for (uint32_t p = 1; p <= ps_low_degree; p++) r.insert(p);
        uint32_t first = ps_low_degree + 1, last = (target_degree / first) * first;
        for (uint32_t p = first; p <= last; p += first) r.insert(p);
    } else {
        for (uint32_t p = 1; p <= target_degree; p++) r.insert(p);
    }
    return r;
}

// ----------------------------------------------------------------------------------------------
// common/apsu/powers.cpp:22-107  PowersDag::configure
// ----------------------------------------------------------------------------------------------
struct PowersNode {
    uint32_t power = 0, depth = 0, p1 = 0, p2 = 0;
    bool is_source() const { return !p1 && !p2; }
};
struct PowersDag {
    std::map<uint32_t, PowersNode> nodes;
    std::set<uint32_t> targets;
    uint32_t depth = 0, source_count = 0;
    bool configured = false;
    bool configure(const std::set<uint32_t> &sources, const std::set<uint32_t> &tgts)
    {
        nodes.clear();
        configured = false;
        if (sources.count(0) || !sources.count(1)) return false;
        if (tgts.count(0) || !tgts.count(1)) return false;
        if (!std::includes(tgts.begin(), tgts.end(), sources.begin(), sources.end())) return false;
        for (uint32_t s : sources) nodes[s] = PowersNode{ s, 0, 0, 0 };
        uint32_t curr_depth = 0;
        for (uint32_t cp : tgts) {
            if (sources.count(cp)) continue;
            uint32_t od = cp - 1, o1 = cp - 1, o2 = 1;
            for (uint32_t s1 : tgts) {
                if (s1 >= cp) break;
                uint32_t s2 = cp - s1;
                if (!tgts.count(s2)) continue;
                uint32_t d = std::max(nodes.at(s1).depth, nodes.at(s2).depth) + 1;
                if (d < od) {
                    od = d;
                    o1 = s1;
                    o2 = s2;
                }
            }
            nodes[cp] = PowersNode{ cp, od, o1, o2 };
            curr_depth = std::max(curr_depth, od);
        }
        configured = true;
        targets = tgts;
        depth = curr_depth;
        source_count = (uint32_t)sources.size();
        return true;
    }
};

// ----------------------------------------------------------------------------------------------
// common/apsu/util/interpolate.cpp:27-80  polyn_with_roots
// ----------------------------------------------------------------------------------------------
static inline std::vector<u64> polyn_with_roots(const std::vector<u64> &roots, const Modulus &mod)
{
    std::vector<u64> p;
    p.reserve(roots.size() + 1);
    p.push_back(1);
    for (u64 a : roots) {
        u64 neg_a = neg_mod(a, mod);
        p.push_back(0);
        for (size_t i = p.size() - 1; i > 0; i--) p[i] = add_mod(mul_mod(p[i], neg_a, mod), p[i - 1], mod);
        p[0] = mul_mod(p[0], neg_a, mod);
    }
    return p;
}

// ----------------------------------------------------------------------------------------------
// APSU parameters needed on the path (common/apsu/psu_params.cpp:95-180 derived fields)
// ----------------------------------------------------------------------------------------------
struct PathParams {
    uint32_t felts_per_item = 0, table_size = 0, max_items_per_bin = 0, ps_low_degree = 0;
    std::set<uint32_t> query_powers;
    uint32_t items_per_bundle = 0, bins_per_bundle = 0, bundle_idx_count = 0;
    void derive(size_t N)
    {
        items_per_bundle = (uint32_t)N / felts_per_item;
        bins_per_bundle = items_per_bundle * felts_per_item;
        if (table_size % items_per_bundle) throw std::invalid_argument("table_size must be a multiple of items_per_bundle");
        bundle_idx_count = table_size / items_per_bundle;
    }
};

// ----------------------------------------------------------------------------------------------
// receiver/apsu/bin_bundle.cpp:366-430  BatchedPlaintextPolyn ctor: column-wise plaintexts.
// batched_coeffs[i] is NTT form (plain level) unless i==0 (non-PS) / i % (ps_low+1)==0 (PS).
// ----------------------------------------------------------------------------------------------
struct BatchedPlaintextPolyn {
    std::vector<Plaintext> batched_coeffs;
    // polyns: one coefficient vector (degree ascending) per bin
    void build(const Context &ctx, const std::vector<std::vector<u64>> &polyns, uint32_t ps_low_degree)
    {
        size_t max_deg = 0;
        for (auto &p : polyns) max_deg = std::max(p.size(), max_deg + 1) - 1;
        size_t plain_L = ctx.level_for_chain_idx(std::min<size_t>(ctx.first_L - 1, ps_low_degree ? 2 : 1));
        Evaluator ev(ctx);
        batched_coeffs.clear();
        std::vector<u64> col(polyns.size());
        for (size_t i = 0; i < max_deg + 1; i++) {
            for (size_t b = 0; b < polyns.size(); b++) col[b] = i < polyns[b].size() ? polyns[b][i] : 0;
            Plaintext pt;
            std::vector<u64> enc(ctx.N);
            batch_encode(ctx, col.data(), col.size(), enc.data());
            bool to_ntt = (!ps_low_degree && i != 0) || (ps_low_degree && (i % (ps_low_degree + 1)));
            if (to_ntt) {
                pt.L = plain_L;
                pt.d.resize(plain_L * ctx.N);
                ev.plain_to_ntt(enc.data(), plain_L, pt.d.data());
            } else {
                pt.L = 0;
                pt.d = std::move(enc);
            }
            batched_coeffs.push_back(std::move(pt));
        }
    }
};

// receiver/apsu/bin_bundle.cpp:67-97  try_clear_irrelevant_bits (last level always has one prime)
static inline void try_clear_irrelevant_bits(const Context &ctx, Ciphertext &c)
{
    if (c.L != 1) return;
    int sig_bits_N = ctx.logN + 1; // get_significant_bit_count(N)
    int keep = ctx.t.bits + sig_bits_N - 1;
    int drop = Modulus(ctx.primes[0]).bits - keep;
    if (drop > 0) {
        u64 mask = ~(((u64)1 << drop) - 1);
        for (auto &x : c.d) x &= mask;
    }
}

using CiphertextPowers = std::vector<Ciphertext>; // index = exponent, [0] dummy

// ----------------------------------------------------------------------------------------------
// receiver/apsu/bin_bundle.cpp:106-174  BatchedPlaintextPolyn::eval
// ----------------------------------------------------------------------------------------------
static inline Ciphertext bp_eval(
    const Context &ctx, const BatchedPlaintextPolyn &bp, const CiphertextPowers &powers, const u64 *random_plain)
{
    if (powers.size() < std::max<size_t>(bp.batched_coeffs.size(), 2)) throw std::invalid_argument("not enough ciphertext powers available");
    Evaluator ev(ctx);
    size_t N = ctx.N;
    Ciphertext result, temp;
    result.resize(N, 2, powers[1].L);
    result.ntt = true;
    for (size_t deg = 1; deg < bp.batched_coeffs.size(); deg++) {
        ev.multiply_plain_ntt(powers[deg], bp.batched_coeffs[deg].d.data(), temp);
        ev.add_inplace(result, temp);
    }
    ev.from_ntt(result);
    ev.add_plain_inplace(result, bp.batched_coeffs[0].d.data());
    ev.add_plain_inplace(result, random_plain);
    while (result.L != 1) ev.mod_switch_to_next(result);
    try_clear_irrelevant_bits(ctx, result);
    return result;
}

// ----------------------------------------------------------------------------------------------
// receiver/apsu/bin_bundle.cpp:192-360  BatchedPlaintextPolyn::eval_patstock
// ----------------------------------------------------------------------------------------------
static inline Ciphertext bp_eval_patstock(
    const Context &ctx, const BatchedPlaintextPolyn &bp, const CiphertextPowers &powers, size_t ps_low_degree,
    const u64 *relin_keys, const u64 *random_plain)
{
    if (powers.size() < std::max<size_t>(bp.batched_coeffs.size(), 2)) throw std::invalid_argument("not enough ciphertext powers available");
    size_t degree = bp.batched_coeffs.size() - 1;
    if (ps_low_degree <= 1 || ps_low_degree >= degree) throw std::invalid_argument("ps_low_degree must be greater than 1 and less than the size of batched_coeffs");
    Evaluator ev(ctx);
    size_t N = ctx.N;
    bool relinearize = ctx.using_keyswitching();
    size_t high_L = ctx.level_for_chain_idx(1);
    size_t h = ps_low_degree + 1, H = degree / h;

    Ciphertext result, temp, temp_in;
    result.resize(N, 3, high_L);
    result.ntt = false;

    auto inner = [&](size_t i, size_t jmax) {
        for (size_t j = 1; j <= jmax; j++) {
            ev.multiply_plain_ntt(powers[j], bp.batched_coeffs[i * h + j].d.data(), temp);
            if (j == 1)
                temp_in = temp;
            else
                ev.add_inplace(temp_in, temp);
        }
        ev.from_ntt(temp_in);
        ev.mod_switch_to(temp_in, high_L);
        Ciphertext prod;
        ev.multiply(temp_in, powers[i * h], prod);
        ev.add_inplace(result, prod);
    };
    for (size_t i = 1; i < H; i++) inner(i, h - 1);         // :248-274
    if (degree % h > 0) inner(H, degree % h);                 // :279-304
    if (relinearize) ev.relinearize(result, relin_keys);      // :308-310
    for (size_t j = 1; j < h; j++) {                          // :314-324  per-term iNTT + mod-switch
        ev.multiply_plain_ntt(powers[j], bp.batched_coeffs[j].d.data(), temp);
        ev.from_ntt(temp);
        ev.mod_switch_to(temp, high_L);
        ev.add_inplace(result, temp);
    }
    for (size_t i = 1; i < H + 1; i++) {                      // :328-337
        ev.multiply_plain_normal(powers[i * h], bp.batched_coeffs[i * h].d.data(), temp);
        ev.mod_switch_to(temp, high_L);
        ev.add_inplace(result, temp);
    }
    ev.add_plain_inplace(result, bp.batched_coeffs[0].d.data()); // :340-345
    ev.add_plain_inplace(result, random_plain);                  // :346
    while (result.L != 1) ev.mod_switch_to_next(result);         // :354-356
    try_clear_irrelevant_bits(ctx, result);                      // :357
    return result;
}

// simple fork-join helper standing in for ThreadPoolMgr (-t workers)
static inline void parallel_for(size_t count, size_t threads, const std::function<void(size_t)> &fn)
{
    if (threads <= 1 || count <= 1) {
        for (size_t i = 0; i < count; i++) fn(i);
        return;
    }
    std::atomic<size_t> next{ 0 };
    std::vector<std::thread> pool;
    size_t nt = std::min(threads, count);
    for (size_t t = 0; t < nt; t++)
        pool.emplace_back([&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= count) return;
                fn(i);
            }
        });
    for (auto &th : pool) th.join();
}

// ----------------------------------------------------------------------------------------------
// receiver/apsu/receiver_ddh.cpp:390-483  Receiver::ComputePowers (one bundle index).
// powers: in/out, index = exponent; source powers present at the first data level, coefficient form.
// The DAG is applied level by level (any dependency-respecting order gives identical results,
// powers.h:158-278).
// ----------------------------------------------------------------------------------------------
static inline void compute_powers(
    const Context &ctx, const PowersDag &pd, uint32_t ps_low_degree, const u64 *relin_keys, CiphertextPowers &powers,
    size_t threads)
{
    Evaluator ev(ctx);
    bool relinearize = ctx.using_keyswitching();
    for (uint32_t d = 1; d <= pd.depth; d++) {
        std::vector<const PowersNode *> todo;
        for (auto &kv : pd.nodes)
            if (!kv.second.is_source() && kv.second.depth == d) todo.push_back(&kv.second);
        parallel_for(todo.size(), threads, [&](size_t i) {
            const PowersNode &nd = *todo[i];
            Ciphertext prod;
            ev.multiply(powers[nd.p1], powers[nd.p2], prod); // square == multiply(x,x)
            if (relinearize) ev.relinearize(prod, relin_keys);
            powers[nd.power] = std::move(prod);
        });
    }
    size_t high_L = ctx.level_for_chain_idx(1), low_L = ctx.level_for_chain_idx(2);
    std::vector<uint32_t> tv(pd.targets.begin(), pd.targets.end());
    parallel_for(tv.size(), threads, [&](size_t i) {
        uint32_t p = tv[i];
        if (!ps_low_degree) {
            ev.mod_switch_to(powers[p], high_L);
            ev.to_ntt(powers[p]);
        } else if (p <= ps_low_degree) {
            ev.mod_switch_to(powers[p], low_L);
            ev.to_ntt(powers[p]);
        } else {
            ev.mod_switch_to(powers[p], high_L);
        }
    });
}

// ----------------------------------------------------------------------------------------------
// The receiver DB as the hot path sees it: per bundle index a list of BinBundle caches
// (receiver/apsu/receiver_db.h:375, bin_bundle.h:137-171).
// ----------------------------------------------------------------------------------------------
struct ReceiverDB {
    std::vector<std::vector<BatchedPlaintextPolyn>> bin_bundles; // [bundle_idx][cache_idx]
    size_t bundle_count() const
    {
        size_t c = 0;
        for (auto &v : bin_bundles) c += v.size();
        return c;
    }
};

struct QueryResult {
    uint32_t bundle_idx, cache_idx;
    Ciphertext ct; // size 2, one prime, coefficient form
};

// ----------------------------------------------------------------------------------------------
// receiver/apsu/receiver_ddh.cpp:295-369 + :485-535  the HE part of RunQuery.
//   query[bundle_idx][source power] -> ciphertext at first data level
//   masks: dense [alpha_max][bundle_idx_count][N] coefficient-form plaintexts indexed by
//          pack_idx = bundle_idx + cache_idx*bundle_idx_count
// ----------------------------------------------------------------------------------------------
static inline std::vector<QueryResult> run_query_he(
    const Context &ctx, const PathParams &pp, const ReceiverDB &db, const PowersDag &pd,
    const std::vector<std::map<uint32_t, Ciphertext>> &query, const u64 *relin_keys, const u64 *masks, size_t threads,
    double *t_powers_ms = nullptr, double *t_eval_ms = nullptr)
{
    using clk = std::chrono::steady_clock;
    size_t N = ctx.N;
    uint32_t bic = pp.bundle_idx_count;
    std::vector<CiphertextPowers> all_powers(bic);
    auto t0 = clk::now();
    for (uint32_t b = 0; b < bic; b++) {
        all_powers[b].assign((size_t)pp.max_items_per_bin + 1, Ciphertext());
        for (auto &kv : query[b]) all_powers[b][kv.first] = kv.second;
    }
    for (uint32_t b = 0; b < bic; b++) {
        if (db.bin_bundles[b].empty()) continue; // :399-402
        compute_powers(ctx, pd, pp.ps_low_degree, relin_keys, all_powers[b], threads);
    }
    auto t1 = clk::now();
    struct Job {
        uint32_t b, c;
    };
    std::vector<Job> jobs;
    for (uint32_t b = 0; b < bic; b++)
        for (uint32_t c = 0; c < db.bin_bundles[b].size(); c++) jobs.push_back({ b, c });
    std::vector<QueryResult> results(jobs.size());
    parallel_for(jobs.size(), threads, [&](size_t k) {
        uint32_t b = jobs[k].b, c = jobs[k].c;
        const BatchedPlaintextPolyn &bp = db.bin_bundles[b][c];
        size_t pack_idx = b + (size_t)c * bic;
        const u64 *mask = masks + pack_idx * N;
        uint32_t degree = (uint32_t)bp.batched_coeffs.size() - 1;
        bool using_ps = pp.ps_low_degree > 1 && pp.ps_low_degree < degree; // :515-517
        results[k].bundle_idx = b;
        results[k].cache_idx = c;
        results[k].ct = using_ps ? bp_eval_patstock(ctx, bp, all_powers[b], pp.ps_low_degree, relin_keys, mask)
                                 : bp_eval(ctx, bp, all_powers[b], mask);
    });
    auto t2 = clk::now();
    if (t_powers_ms) *t_powers_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (t_eval_ms) *t_eval_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    return results;
}

} // namespace orc
