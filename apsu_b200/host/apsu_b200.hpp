// C++ host facade over the C ABI (include/apsu_b200.h): the reference's receiver-side surface for the
// query-evaluation path, same names / argument meaning / exception types, so that the call sites in
// receiver/apsu/receiver_ddh.cpp and receiver/apsu/zmq/receiver_dispatcher_ddh.cpp read unchanged.
//
//   apsu::PSUParams                      common/apsu/psu_params.h:31-222
//   apsu::PowersDag                      common/apsu/powers.h:41-293
//   apsu::receiver::BatchedPlaintextPolyn / BinBundleCache   receiver/apsu/bin_bundle.h:52-171
//   apsu::receiver::ReceiverDB           receiver/apsu/receiver_db.h:60-390 (query-time surface)
//   apsu::receiver::Query                receiver/apsu/query.h:26-89
//   apsu::network::ResultPackage         common/apsu/network/result_package.h:44-62
//   apsu::receiver::Receiver             receiver/apsu/receiver_ddh.h:184-243 (HE part of RunQuery)
//
// Header-only; link with libapsu_b200.so.  Ciphertexts/plaintexts are plain std::vector<uint64_t> in SEAL's
// in-memory layouts (see the C header) — (de)serialisation of SEAL objects stays with the caller.
#pragma once
#include "../../include/apsu_b200.h"
#include <cstdint>
#include <algorithm>
#include <exception>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

namespace apsu {

namespace detail {
// status -> the exception type the reference throws at the same place
inline void check(int rc)
{
    if (rc == APSU_B200_OK) return;
    std::string msg = apsu_b200_last_error();
    switch (rc) {
    case APSU_B200_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
    case APSU_B200_ERR_LOGIC: throw std::logic_error(msg);
    default: throw std::runtime_error(msg);
    }
}
} // namespace detail

class PSUParams {
public:
    struct TableParams {
        std::uint32_t hash_func_count, table_size, max_items_per_bin;
    };
    struct ItemParams {
        std::uint32_t felts_per_item;
    };
    struct QueryParams {
        std::uint32_t ps_low_degree;
        std::set<std::uint32_t> query_powers;
    };
    struct SEALParams {
        std::size_t poly_modulus_degree;
        std::uint64_t plain_modulus;
        std::vector<std::uint64_t> coeff_modulus;
    };

    // PSUParams::Load(const std::string&) — JSON text in the reference's schema
    static PSUParams Load(const std::string &json)
    {
        PSUParams p;
        detail::check(apsu_b200_params_load_json(json.c_str(), &p.c_));
        return p;
    }
    PSUParams(const ItemParams &ip, const TableParams &tp, const QueryParams &qp, const SEALParams &sp)
    {
        c_ = apsu_b200_params{};
        c_.poly_modulus_degree = static_cast<std::uint32_t>(sp.poly_modulus_degree);
        c_.plain_modulus = sp.plain_modulus;
        if (sp.coeff_modulus.size() > APSU_B200_MAX_COEFF_MODULUS) throw std::invalid_argument("coeff_modulus is too long");
        c_.coeff_modulus_count = static_cast<std::uint32_t>(sp.coeff_modulus.size());
        for (std::size_t i = 0; i < sp.coeff_modulus.size(); i++) c_.coeff_modulus[i] = sp.coeff_modulus[i];
        c_.hash_func_count = tp.hash_func_count;
        c_.table_size = tp.table_size;
        c_.max_items_per_bin = tp.max_items_per_bin;
        c_.felts_per_item = ip.felts_per_item;
        c_.ps_low_degree = qp.ps_low_degree;
        if (qp.query_powers.size() > APSU_B200_MAX_QUERY_POWERS) throw std::invalid_argument("too many query_powers");
        for (auto q : qp.query_powers) c_.query_powers[c_.query_power_count++] = q;
        detail::check(apsu_b200_params_validate(&c_)); // == PSUParams::initialize
    }

    TableParams table_params() const { return { c_.hash_func_count, c_.table_size, c_.max_items_per_bin }; }
    ItemParams item_params() const { return { c_.felts_per_item }; }
    QueryParams query_params() const
    {
        return { c_.ps_low_degree, std::set<std::uint32_t>(c_.query_powers, c_.query_powers + c_.query_power_count) };
    }
    SEALParams seal_params() const
    {
        return { c_.poly_modulus_degree, c_.plain_modulus,
                 std::vector<std::uint64_t>(c_.coeff_modulus, c_.coeff_modulus + c_.coeff_modulus_count) };
    }
    std::uint32_t items_per_bundle() const { return c_.items_per_bundle; }
    std::uint32_t bins_per_bundle() const { return c_.bins_per_bundle; }
    std::uint32_t bundle_idx_count() const { return c_.bundle_idx_count; }
    std::uint32_t item_bit_count() const { return c_.item_bit_count; }
    std::uint32_t item_bit_count_per_felt() const { return c_.item_bit_count_per_felt; }
    const apsu_b200_params &c_params() const { return c_; }

private:
    PSUParams() = default;
    apsu_b200_params c_{};
};

class PowersDag {
public:
    struct PowersNode {
        std::uint32_t power = 0, depth = 0;
        std::pair<std::uint32_t, std::uint32_t> parents{ 0, 0 };
        bool is_source() const { return !parents.first && !parents.second; }
    };
    // configure(params.query_powers, create_powers_set(ps_low_degree, max_items_per_bin)) as in query.cpp:78
    bool configure(const PSUParams &params)
    {
        const apsu_b200_params &c = params.c_params();
        std::uint32_t cap = c.max_items_per_bin + 1, n = 0;
        std::vector<std::uint32_t> pw(cap), dp(cap), p1(cap), p2(cap);
        if (apsu_b200_powers_dag(&c, cap, pw.data(), dp.data(), p1.data(), p2.data(), &n, &depth_) != APSU_B200_OK) return false;
        nodes_.clear();
        target_powers_.clear();
        for (std::uint32_t i = 0; i < n; i++) {
            nodes_[pw[i]] = PowersNode{ pw[i], dp[i], { p1[i], p2[i] } };
            target_powers_.insert(pw[i]);
        }
        configured_ = true;
        return true;
    }
    bool is_configured() const { return configured_; }
    std::uint32_t depth() const { need(); return depth_; }
    std::set<std::uint32_t> target_powers() const { need(); return target_powers_; }
    std::uint32_t source_count() const { return static_cast<std::uint32_t>(source_nodes().size()); }
    std::vector<PowersNode> source_nodes() const
    {
        need();
        std::vector<PowersNode> r;
        for (auto &kv : nodes_)
            if (kv.second.is_source()) r.push_back(kv.second);
        return r;
    }
    template <typename Func>
    void apply(Func &&func) const
    {
        need();
        for (auto p : target_powers_) func(nodes_.at(p));
    }

private:
    void need() const
    {
        if (!configured_) throw std::logic_error("PowersDag has not been configured");
    }
    std::unordered_map<std::uint32_t, PowersNode> nodes_;
    std::set<std::uint32_t> target_powers_;
    std::uint32_t depth_ = 0;
    bool configured_ = false;
};

namespace network {
// one per BinBundle; psu_result = size-2 ciphertext at the last level, uint64_t[2][1][N]
struct ResultPackage {
    std::uint32_t bundle_idx = 0, cache_idx = 0;
    std::uint32_t label_byte_count = 0, nonce_byte_count = 0;
    std::vector<std::uint64_t> psu_result;
};
} // namespace network

namespace receiver {

// batched_coeffs[k]: NTT form uint64_t[Lp][N] unless k==0 (no PS) / k % (ps_low_degree+1)==0 (PS): uint64_t[N]
struct BatchedPlaintextPolyn {
    std::vector<std::vector<std::uint64_t>> batched_coeffs;
    explicit operator bool() const { return !batched_coeffs.empty(); }
};
struct BinBundleCache {
    BatchedPlaintextPolyn batched_matching_polyn;
};

class ReceiverDB {
public:
    explicit ReceiverDB(PSUParams params, int device = 0) : params_(std::move(params))
    {
        detail::check(apsu_b200_ctx_create(&params_.c_params(), device, &ctx_));
    }
    ~ReceiverDB()
    {
        apsu_b200_ctx_destroy(ctx_);
        for (auto &b : pinned_) apsu_b200_host_free(b.p);
    }
    ReceiverDB(const ReceiverDB &) = delete;
    ReceiverDB &operator=(const ReceiverDB &) = delete;

    const PSUParams &get_params() const { return params_; }
    // uploads one BinBundle cache; it stays device-resident (bin_bundles_[bundle_idx].push_back)
    std::uint32_t add_bin_bundle(std::uint32_t bundle_idx, const BinBundleCache &cache)
    {
        std::unique_lock<std::shared_mutex> lock(db_lock_);
        std::vector<const std::uint64_t *> ptrs;
        for (auto &v : cache.batched_matching_polyn.batched_coeffs) ptrs.push_back(v.data());
        std::uint32_t ci = 0;
        detail::check(apsu_b200_db_add_binbundle(ctx_, bundle_idx, ptrs.data(), static_cast<std::uint32_t>(ptrs.size()), &ci));
        return ci;
    }
    std::uint32_t add_bin_bundle_synthetic(std::uint32_t bundle_idx, std::uint32_t ncoeffs, std::uint64_t seed)
    {
        std::unique_lock<std::shared_mutex> lock(db_lock_);
        std::uint32_t ci = 0;
        detail::check(apsu_b200_db_add_binbundle_synthetic(ctx_, bundle_idx, ncoeffs, seed, &ci));
        return ci;
    }
    // RunQuery's mask loop on the device (receiver_ddh.cpp:218-283): SEAL's blake2xb generator keyed with 64 bytes of
    // OS randomness (seed == nullptr, the reference's random_bytes) or with a caller-chosen 64-byte seed (reproducible
    // runs); returns random_matrix as (low, high) words per item, [alpha_max*bundle_idx_count][items_per_bundle][2];
    // the masks stay device-resident for the next query
    std::vector<std::uint64_t> generate_masks(const std::uint8_t *seed = nullptr)
    {
        std::unique_lock<std::shared_mutex> lock(db_lock_);
        const std::uint32_t bic = params_.bundle_idx_count();
        std::uint32_t alpha = 1;
        std::vector<std::uint32_t> cnt(bic);
        for (std::uint32_t b = 0; b < bic; b++) {
            detail::check(apsu_b200_db_bin_bundle_count(ctx_, b, &cnt[b]));
            alpha = std::max(alpha, cnt[b]);
        }
        std::vector<std::uint8_t> padded(static_cast<std::size_t>(alpha) * bic);
        for (std::uint32_t c = 0; c < alpha; c++)
            for (std::uint32_t b = 0; b < bic; b++) padded[b + static_cast<std::size_t>(c) * bic] = c >= cnt[b];
        std::vector<std::uint64_t> blocks(padded.size() * params_.items_per_bundle() * 2);
        detail::check(apsu_b200_generate_masks(ctx_, seed, padded.data(), static_cast<std::uint32_t>(padded.size()), blocks.data(), nullptr));
        return blocks;
    }
    // BinBundle::regen_cache on the device (bin_bundle.cpp:934-1041): item_bins -> matching polynomials ->
    // batched plaintexts, all built and kept on the GPU
    std::uint32_t add_bin_bundle_from_bins(std::uint32_t bundle_idx, const std::vector<std::vector<std::uint64_t>> &item_bins)
    {
        std::unique_lock<std::shared_mutex> lock(db_lock_);
        std::vector<std::uint32_t> sizes;
        std::vector<std::uint64_t> roots;
        for (auto &b : item_bins) {
            sizes.push_back(static_cast<std::uint32_t>(b.size()));
            roots.insert(roots.end(), b.begin(), b.end());
        }
        if (roots.empty()) roots.push_back(0);
        std::uint32_t ci = 0;
        detail::check(apsu_b200_db_add_binbundle_from_bins(ctx_, bundle_idx, sizes.data(), roots.data(), &ci));
        return ci;
    }
    std::size_t get_bin_bundle_count(std::uint32_t bundle_idx) const
    {
        std::uint32_t n = 0;
        detail::check(apsu_b200_db_bin_bundle_count(ctx_, bundle_idx, &n));
        return n;
    }
    std::size_t get_bin_bundle_count() const
    {
        std::uint32_t n = 0;
        detail::check(apsu_b200_db_total_bin_bundle_count(ctx_, &n));
        return n;
    }
    void clear()
    {
        std::unique_lock<std::shared_mutex> lock(db_lock_);
        detail::check(apsu_b200_db_clear(ctx_));
    }
    std::shared_lock<std::shared_mutex> get_reader_lock() const { return std::shared_lock<std::shared_mutex>(db_lock_); }
    apsu_b200_ctx *handle() const { return ctx_; }
    // a context is thread-compatible, not thread-safe: callers of the query path serialise on THIS context's mutex
    std::mutex &context_mutex() const { return ctx_mutex_; }
    // pinned host staging owned by the DB and reused between queries (slot 0: query ciphertexts, slot 1: results)
    std::uint64_t *pinned(int slot, std::size_t words) const
    {
        Pinned &b = pinned_[slot];
        if (words > b.words) {
            apsu_b200_host_free(b.p);
            b.p = nullptr;
            b.words = 0;
            void *q = nullptr;
            detail::check(apsu_b200_host_alloc(words * sizeof(std::uint64_t), &q));
            b.p = static_cast<std::uint64_t *>(q);
            b.words = words;
        }
        return b.p;
    }

private:
    struct Pinned {
        std::uint64_t *p = nullptr;
        std::size_t words = 0;
    };
    PSUParams params_;
    apsu_b200_ctx *ctx_ = nullptr;
    mutable std::shared_mutex db_lock_;
    mutable std::mutex ctx_mutex_;
    mutable Pinned pinned_[2];
};

// the already-extracted query: source power -> one ciphertext per bundle index (uint64_t[2][L][N]), relin keys
class Query {
public:
    Query(std::shared_ptr<ReceiverDB> db, std::unordered_map<std::uint32_t, std::vector<std::vector<std::uint64_t>>> data,
          std::vector<std::uint64_t> relin_keys)
        : receiver_db_(std::move(db)), data_(std::move(data)), relin_keys_(std::move(relin_keys))
    {
        if (!receiver_db_) throw std::invalid_argument("receiver_db cannot be null");
        const PSUParams &p = receiver_db_->get_params();
        std::set<std::uint32_t> given;
        for (auto &kv : data_) {
            given.insert(kv.first);
            if (kv.second.size() != p.bundle_idx_count()) return; // invalid: valid_ stays false (query.cpp:93-99)
        }
        if (given != p.query_params().query_powers) return; // query.cpp:68-111
        if (!pd_.configure(p)) return;
        valid_ = true;
    }
    explicit operator bool() const { return valid_; }
    std::shared_ptr<ReceiverDB> receiver_db() const { return receiver_db_; }
    const PowersDag &pd() const { return pd_; }
    const std::unordered_map<std::uint32_t, std::vector<std::vector<std::uint64_t>>> &data() const { return data_; }
    const std::vector<std::uint64_t> &relin_keys() const { return relin_keys_; }

private:
    std::shared_ptr<ReceiverDB> receiver_db_;
    std::unordered_map<std::uint32_t, std::vector<std::vector<std::uint64_t>>> data_;
    std::vector<std::uint64_t> relin_keys_;
    PowersDag pd_;
    bool valid_ = false;
};

using ResultPart = std::unique_ptr<network::ResultPackage>;

class Receiver {
public:
    // The HE part of Receiver::RunQuery (receiver_ddh.cpp:139-369): `masks` is the dense
    // [alpha_max_cache_count][bundle_idx_count][N] table of coefficient-form random plaintexts
    // (random_plain_list, indexed by pack_idx), or empty to keep the masks ReceiverDB::generate_masks left on the
    // device; send_rp_fun receives one ResultPackage per BinBundle AS ITS BinBundle FINISHES (the reference sends from
    // the pool thread that evaluated it, receiver_ddh.cpp:527-534): chunk by chunk, while later chunks are evaluated.
    static void RunQuery(const Query &query, const std::vector<std::uint64_t> &masks,
                         const std::function<void(ResultPart)> &send_rp_fun)
    {
        if (!query) throw std::invalid_argument("query is invalid");
        auto db = query.receiver_db();
        auto lock = db->get_reader_lock();
        std::lock_guard<std::mutex> guard(db->context_mutex()); // a context is thread-compatible: serialise per context
        const PSUParams &p = db->get_params();
        const std::size_t N = p.seal_params().poly_modulus_degree;
        std::uint32_t L = 0;
        detail::check(apsu_b200_ctx_level(db->handle(), 0, &L));
        const std::size_t ct_words = 2 * static_cast<std::size_t>(L) * N, bic = p.bundle_idx_count();
        const std::size_t n = db->get_bin_bundle_count();
        // one copy of the ciphertexts, straight into pinned staging kept by the DB (uploads from pageable memory are staged)
        std::uint64_t *cts = db->pinned(0, query.data().size() * bic * ct_words);
        std::vector<std::uint32_t> src;
        std::size_t at = 0;
        for (auto &kv : query.data()) {
            src.push_back(kv.first);
            for (auto &ct : kv.second) {
                if (ct.size() != ct_words) throw std::invalid_argument("query ciphertext has the wrong size");
                std::copy(ct.begin(), ct.end(), cts + at);
                at += ct_words;
            }
        }
        if (masks.size() % N) throw std::invalid_argument("mask table is not a multiple of poly_modulus_degree");
        detail::check(apsu_b200_query_begin(db->handle(), src.data(), static_cast<std::uint32_t>(src.size()), cts));
        detail::check(apsu_b200_set_relin_keys(db->handle(), query.relin_keys().empty() ? nullptr : query.relin_keys().data()));
        if (!masks.empty()) detail::check(apsu_b200_set_masks(db->handle(), masks.data(), static_cast<std::uint32_t>(masks.size() / N)));
        detail::check(apsu_b200_compute_powers(db->handle()));
        std::uint64_t *out = db->pinned(1, n * 2 * N);
        struct Sink {
            const std::function<void(ResultPart)> *send;
            std::size_t words;
        } sink{ &send_rp_fun, 2 * N };
        detail::check(apsu_b200_eval_all_stream(
            db->handle(), out,
            [](void *user, std::uint32_t bundle_idx, std::uint32_t cache_idx, const std::uint64_t *ct) {
                auto *sk = static_cast<Sink *>(user);
                auto rp = std::make_unique<network::ResultPackage>();
                rp->bundle_idx = bundle_idx;
                rp->cache_idx = cache_idx;
                rp->psu_result.assign(ct, ct + sk->words);
                (*sk->send)(std::move(rp));
            },
            &sink));
    }
};

// BinBundles sharded over the GPUs of one box: one ReceiverDB (context) per GPU, one host thread per
// GPU inside the calls, all exchanges inside the library (apsu_b200_mgpu_*, NCCL).  The unit that is distributed is the
// unit the reference hands to its thread pool: the BinBundle (receiver_ddh.cpp:340-364).
class MultiGpuReceiver {
public:
    MultiGpuReceiver(const PSUParams &params, const std::vector<int> &devices)
    {
        if (devices.empty()) throw std::invalid_argument("no devices");
        for (int d : devices) dbs_.push_back(std::make_shared<ReceiverDB>(params, d));
        gidx_.resize(devices.size());
        next_cache_.assign(params.bundle_idx_count(), 0);
    }
    ~MultiGpuReceiver()
    {
        for (auto m : mg_) apsu_b200_mgpu_destroy(m);
    }
    std::size_t world() const { return dbs_.size(); }
    std::shared_ptr<ReceiverDB> db(std::size_t rank) const { return dbs_.at(rank); }
    // appends a BinBundle to bundle index `bundle_idx` of the whole DB, stored on GPU `rank`; returns its cache index
    std::uint32_t add_bin_bundle(std::size_t rank, std::uint32_t bundle_idx, const BinBundleCache &cache)
    {
        dbs_.at(rank)->add_bin_bundle(bundle_idx, cache);
        return note(rank, bundle_idx);
    }
    std::uint32_t add_bin_bundle_synthetic(std::size_t rank, std::uint32_t bundle_idx, std::uint32_t ncoeffs, std::uint64_t seed)
    {
        dbs_.at(rank)->add_bin_bundle_synthetic(bundle_idx, ncoeffs, seed);
        return note(rank, bundle_idx);
    }
    // collective set-up after the DB is loaded (and after every change): dag_split as in apsu_b200_mgpu_commit
    void commit(int dag_split = -1)
    {
        if (mg_.empty()) {
            std::uint8_t id[APSU_B200_MGPU_ID_BYTES];
            detail::check(apsu_b200_mgpu_unique_id(id));
            mg_.assign(world(), nullptr);
            each_rank([&](std::size_t r) { detail::check(apsu_b200_mgpu_create(dbs_[r]->handle(), id, (std::uint32_t)r, (std::uint32_t)world(), &mg_[r])); });
        }
        each_rank([&](std::size_t r) {
            std::vector<std::uint32_t> g;
            for (auto &kv : gidx_[r]) g.insert(g.end(), kv.second.begin(), kv.second.end()); // bundle_idx major, insertion order
            detail::check(apsu_b200_mgpu_commit(mg_[r], g.empty() ? nullptr : g.data(), dag_split));
        });
    }
    // RunQuery over all GPUs: `query` was built against db(0); masks[r] = rank r's dense mask table for ITS local cache
    // indices (empty: keep the masks generated on that GPU).  One ResultPackage per BinBundle of the whole DB.
    void RunQuery(const Query &query, const std::vector<std::vector<std::uint64_t>> &masks, const std::function<void(ResultPart)> &send_rp_fun)
    {
        if (!query) throw std::invalid_argument("query is invalid");
        if (mg_.empty()) throw std::logic_error("MultiGpuReceiver::commit has not been called");
        const PSUParams &p = dbs_[0]->get_params();
        const std::size_t N = p.seal_params().poly_modulus_degree, bic = p.bundle_idx_count();
        std::uint32_t L = 0, total = 0;
        detail::check(apsu_b200_ctx_level(dbs_[0]->handle(), 0, &L));
        detail::check(apsu_b200_mgpu_info(mg_[0], &total, nullptr, nullptr, nullptr));
        const std::size_t ct_words = 2 * static_cast<std::size_t>(L) * N;
        std::uint64_t *cts = dbs_[0]->pinned(0, query.data().size() * bic * ct_words);
        std::vector<std::uint32_t> src;
        std::size_t at = 0;
        for (auto &kv : query.data()) {
            src.push_back(kv.first);
            for (auto &ct : kv.second) {
                if (ct.size() != ct_words) throw std::invalid_argument("query ciphertext has the wrong size");
                std::copy(ct.begin(), ct.end(), cts + at);
                at += ct_words;
            }
        }
        // every GPU delivers the result ciphertexts of its own BinBundles to its own pinned buffer (no gather: the
        // device-to-host copies of all GPUs run in parallel), then the packages are handed over in rank order
        std::vector<std::uint64_t *> out(world(), nullptr);
        std::vector<std::vector<std::uint32_t>> bidx(world()), cidx(world());
        each_rank([&](std::size_t r) {
            const std::vector<std::uint64_t> *m = r < masks.size() && !masks[r].empty() ? &masks[r] : nullptr;
            std::uint32_t mine = 0;
            detail::check(apsu_b200_mgpu_local_count(mg_[r], &mine));
            out[r] = dbs_[r]->pinned(1, static_cast<std::size_t>(std::max<std::uint32_t>(mine, 1)) * 2 * N);
            bidx[r].assign(mine, 0);
            cidx[r].assign(mine, 0);
            // the ranks are threads of this process: every GPU uploads its own part of the query in parallel
            detail::check(apsu_b200_mgpu_run_query_local(
                mg_[r], src.data(), static_cast<std::uint32_t>(src.size()), cts,
                !query.relin_keys().empty() ? query.relin_keys().data() : nullptr, m ? m->data() : nullptr,
                m ? static_cast<std::uint32_t>(m->size() / N) : 0, out[r], bidx[r].data(), cidx[r].data()));
        });
        std::uint32_t delivered = 0;
        for (std::size_t r = 0; r < world(); r++)
            for (std::size_t k = 0; k < bidx[r].size(); k++, delivered++) {
                auto rp = std::make_unique<network::ResultPackage>();
                rp->bundle_idx = bidx[r][k];
                rp->cache_idx = cidx[r][k];
                rp->psu_result.assign(out[r] + k * 2 * N, out[r] + (k + 1) * 2 * N);
                send_rp_fun(std::move(rp));
            }
        if (delivered != total) throw std::logic_error("MultiGpuReceiver: result count mismatch");
    }

private:
    std::uint32_t note(std::size_t rank, std::uint32_t bundle_idx)
    {
        const std::uint32_t g = next_cache_.at(bundle_idx)++;
        gidx_[rank][bundle_idx].push_back(g);
        return g;
    }
    // the collective calls block until every rank has entered them: one host thread per GPU
    template <typename F>
    void each_rank(F &&f)
    {
        std::vector<std::thread> th;
        std::vector<std::exception_ptr> err(world());
        for (std::size_t r = 0; r < world(); r++)
            th.emplace_back([&, r] {
                try {
                    f(r);
                } catch (...) {
                    err[r] = std::current_exception();
                }
            });
        for (auto &t : th) t.join();
        for (auto &e : err)
            if (e) std::rethrow_exception(e);
    }
    std::vector<std::shared_ptr<ReceiverDB>> dbs_;
    std::vector<apsu_b200_mgpu *> mg_;
    std::vector<std::map<std::uint32_t, std::vector<std::uint32_t>>> gidx_; // [rank][bundle_idx] -> global cache indices
    std::vector<std::uint32_t> next_cache_;
};

} // namespace receiver
} // namespace apsu
