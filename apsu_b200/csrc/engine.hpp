// Engine: device-resident ReceiverDB + the query-evaluation programs (ComputePowers and
// eval/eval_patstock for every BinBundle) of one APSU receiver on one B200.
// Reference: receiver/apsu/receiver_ddh.cpp:295-369, 390-535; receiver/apsu/bin_bundle.cpp:106-360.
#pragma once
#include "context.hpp"
#include <functional>
#include <memory>

namespace apsu_b200 {

struct ProgramBuilder;
struct KtGroup;
struct MulTermsJob;
struct FinalizeJob;

// 32-bit index arrays used by the kernels; collected on the host while a program is built and uploaded
// in one copy.  Handles are offsets into the pool.
struct IdxPool {
    std::vector<uint32_t> host;
    DBuf<uint32_t> dev;
    size_t add(const std::vector<uint32_t> &v)
    {
        size_t off = host.size();
        host.insert(host.end(), v.begin(), v.end());
        return off;
    }
    void upload(cudaStream_t st) { dev.upload(host, st); }
    const uint32_t *at(size_t off) const { return dev.p + off; }
    void clear() { host.clear(); }
};

// bump allocator over the polynomial arena (units: one polynomial of N words)
struct Arena {
    DBuf<u64> buf;
    size_t top = 0, high_water = 0;
    uint32_t take(size_t polys)
    {
        size_t at = top;
        top += polys;
        if (top > high_water) high_water = top;
        if (top >= (1ull << 32)) throw std::runtime_error("polynomial arena exceeds 2^32 polynomials");
        return (uint32_t)at;
    }
};

struct BinBundleStore {
    uint32_t bundle_idx = 0, cache_idx = 0;
    uint32_t ncoeffs = 0;
    // NTT-form plaintexts in degree order (coefficient-form degrees skipped), tile-major and split
    // (db_stream.cuh): [low_L*N/128][n_ntt][128]
    DBuf<u64> ntt_coeffs;
    // coefficient-form plaintexts (degree 0 and, with PS, multiples of ps_low+1): [n_plain][N]
    DBuf<u64> plain_coeffs;
    // PS only: the coefficient-form plaintexts of degree i*(ps_low+1), i>=1, lifted and NTT'd at the high
    // level once at upload (multiply_plain on coefficient-form operands, bin_bundle.cpp:328-337), tile-major
    // and split: [high_L*N/128][n_plain-1][128]
    DBuf<u64> plain_high_ntt;
    uint32_t n_ntt = 0, n_plain = 0;
};

class Engine {
public:
    Engine(const apsu_b200_params &p, int device);
    ~Engine();

    DeviceContext ctx;
    PowersDag dag;

    // ---- DB ----
    std::vector<std::vector<std::unique_ptr<BinBundleStore>>> db; // [bundle_idx][cache_idx]
    uint32_t add_binbundle(uint32_t bundle_idx, const uint64_t *const *coeffs, uint32_t ncoeffs);
    uint32_t add_binbundle_synthetic(uint32_t bundle_idx, uint32_t ncoeffs, uint64_t seed);
    uint32_t add_binbundle_from_bins(uint32_t bundle_idx, const uint32_t *bin_sizes, const uint64_t *roots);
    uint32_t add_binbundle_from_bins_device(uint32_t bundle_idx, const uint32_t *d_first, const uint32_t *d_size, const u64 *d_roots, uint32_t max_deg);
    void throw_if_build_invalid();
    // ReceiverDB::set_data / insert_or_assign on an empty DB (receiver_db.cpp:330-438, 966): first-fit insertion of the
    // algebraised items into BinBundles and the build of every cache, on the device (dbbuild.cu)
    void set_data(const uint64_t *felts, const uint64_t *cuckoo_idx, size_t n, bool on_device, uint32_t *bundle_counts);
    uint32_t total_bundles() const;
    uint64_t stream_bytes() const;
    void clear_db();
    int db_split() const { return split_; } // bit position the stored NTT-form words are split at
    bool is_ntt_degree(uint32_t k) const
    {
        uint32_t ps = ctx.params.ps_low_degree;
        return ps ? (k % (ps + 1)) != 0 : k != 0;
    }

    // ---- query ----
    void set_relin_keys(const void *keys, bool on_device);
    void query_begin(const uint32_t *src_powers, uint32_t nsrc, const void *cts, bool on_device);
    void query_begin_seeded(const uint32_t *src_powers, uint32_t nsrc, const uint64_t *c0, const uint8_t *seeds64);
    void set_relin_keys_seeded(const uint64_t *c0, const uint8_t *seeds64);
    // relinearisation keys filled in place by the caller (multi-GPU broadcast): reserve, write/receive on the context
    // stream, then relin_keys_loaded() (runs the residue check)
    void reserve_relin_keys() { relin_keys_.ensure((size_t)(ctx.K - 1) * 2 * ctx.K * ctx.N); }
    u64 *relin_keys_device() { return relin_keys_.p; }
    void relin_keys_loaded()
    {
        check_range(relin_keys_.p, (ctx.K - 1) * 2 * ctx.K, ctx.params.coeff_modulus, ctx.K);
        have_keys_ = true;
    }
    void query_begin_partial(const uint32_t *src_powers, uint32_t nsrc);
    void query_load_index(uint32_t bundle_idx, const void *cts_device);
    const std::vector<std::pair<uint32_t, uint32_t>> &result_order()
    {
        if (!plan_valid_) build_plan();
        return result_order_;
    }
    void throw_if_query_invalid(); // reads back the is_valid_for flag of the loaded query / keys (synchronises)
    void set_masks(const void *masks, uint32_t npack, bool on_device);
    // host masks uploaded on the copy stream, behind ComputePowers: eval_all waits for them (apsu_b200_run_query)
    void set_masks_overlapped(const void *masks, uint32_t npack);
    void join_masks_upload();
    void encode_masks(const uint64_t *slot_values, uint32_t npack, uint64_t *out);
    void generate_masks(const uint8_t *seed64, const uint8_t *padded, uint32_t npack, uint64_t *blocks_out, uint64_t *values_out, bool synchronise = true);
    void decrypt_results(const uint64_t *secret_ntt_q0, const uint64_t *cts, uint32_t n, uint64_t *values_out, uint64_t *blocks_out, int32_t *budget_out);
    void compute_powers();
    // PowersDag split over the ranks that share a bundle index (collective C2)
    void set_powers_partition(uint32_t rank, uint32_t size);
    uint32_t powers_stage_count();
    void compute_powers_stage(uint32_t stage);
    uint32_t powers_exchange_regions(uint32_t level, void **ptrs, uint64_t *chunk_bytes, uint32_t capacity);
    // split PowersDag exchanged through NVLink peer memory instead of a collective: the key-switch epilogue of every DAG
    // level stores its products into the peers' arenas too, a flag barrier on the stream closes the level (kernels.cuh)
    void set_powers_p2p(const std::vector<u64 *> &arenas, const std::vector<uint32_t *> &flags);
    bool powers_p2p_enabled() const { return p2p_.enabled; }
    u64 *arena_base();
    uint32_t *p2p_flags();
    void eval_all();
    void fetch_results(uint64_t *out, uint32_t *bundle_idx, uint32_t *cache_idx);
    void eval_all_stream(uint64_t *out, void (*fn)(void *, uint32_t, uint32_t, const uint64_t *), void *user);
    void set_eval_chunk(uint32_t n)
    {
        eval_chunk_ = n ? n : 32;
        invalidate_plan();
    }
    void get_power(uint32_t bundle_idx, uint32_t power, uint64_t *out, uint32_t *L, int *is_ntt);
    void results_device(void **ptr, uint64_t *bytes);
    void collect_timings();

    // ---- stand-alone batched evaluator operations ----
    void op_ntt(uint64_t *polys, uint32_t count, const uint32_t *modulus_index, uint32_t pattern_len, bool inverse);
    void op_multiply(uint32_t L, const uint64_t *a, const uint64_t *b, uint64_t *out, uint32_t n_ops);
    void op_relinearize(uint32_t L, const uint64_t *in, uint64_t *out, uint32_t n_ops);
    void op_mod_switch_next(uint32_t L, const uint64_t *in, uint64_t *out, uint32_t n_polys);
    void op_prng_stream(const uint8_t *seed64, uint64_t counter0, uint64_t *out, size_t n_words);
    void op_expand_seeds(uint32_t L, const uint8_t *seeds64, uint32_t n, uint64_t *out);

    apsu_b200_timings timings{};
    bool profiling = false;

    struct Step {
        std::function<void()> run;
    };

private:
    friend struct ProgramBuilder;
    // device state
    DBuf<u64> relin_keys_; // [K-1][2][K][N]
    bool have_keys_ = false;
    DBuf<u64> masks_; // [npack][N]
    uint32_t npack_ = 0, npack_needed_ = 0;
    DBuf<u64> results_; // [total_bundles][2][N]
    std::vector<std::pair<uint32_t, uint32_t>> result_order_;
    DBuf<LevelConsts> levels_dev_;

    // the polynomial arena and the two programs
    Arena arena_;
    IdxPool idx_;
    std::vector<uint8_t> desc_host_; // kernel descriptor structs (KtGroup, FinalizeJob, ...)
    DBuf<uint8_t> desc_dev_;
    std::vector<Step> powers_prog_, eval_prog_;
    std::vector<size_t> powers_stage_end_; // powers_prog_ index after each DAG level and after the tail
    uint32_t powers_part_rank_ = 0, powers_part_size_ = 1;
    struct ExchangeRegion {
        uint32_t level, region, chunk_polys;
    };
    std::vector<ExchangeRegion> powers_exchange_;
    bool plan_valid_ = false;
    bool query_loaded_ = false, powers_done_ = false, eval_done_ = false;
    uint32_t query_region_ = 0; // arena index of the uploaded query [nsrc][bic][2][first_L]
    // where each power lives after ComputePowers
    struct PowerLoc {
        uint32_t idx = 0, L = 0;
        bool ntt = false, valid = false;
    };
    std::vector<std::vector<PowerLoc>> final_power_; // [bic][max_items_per_bin+1]

    // timing
    cudaEvent_t ev_[4] = { nullptr, nullptr, nullptr, nullptr };
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> mac_events_;
    size_t mac_events_used_ = 0;
    std::vector<uint64_t> mac_step_bytes_;
    uint64_t timed_mac_bytes_ = 0;
    uint32_t powers_launches_ = 0, eval_launches_ = 0;

    // The two programs are fixed launch sequences: they are captured once into CUDA graphs (one per ComputePowers
    // stage, one for the evaluation) and replayed per query; the small parameter sets are launch-bound.
    struct ProgGraph {
        cudaGraphExec_t exec = nullptr;
        uint32_t launches = 0;
        size_t mac_events = 0;
        uint64_t mac_bytes = 0;
        const void *key[4] = { nullptr, nullptr, nullptr, nullptr }; // arena, masks, relin keys, results at capture
        bool profiling = false;
    };
    std::vector<ProgGraph> powers_graphs_;
    ProgGraph eval_graph_;
    bool use_graphs_ = true;
    uint32_t eval_chunk_ = 32; // BinBundles per Paterson-Stockmeyer chunk (= per result delivery of eval_all_stream)
    struct FinGroup { // one finalize launch: the result slots it completes
        std::vector<uint32_t> slots;
        cudaEvent_t done = nullptr, copied = nullptr;
    };
    std::vector<FinGroup> fin_groups_;
    cudaStream_t copy_stream_ = nullptr;
    cudaEvent_t masks_free_ = nullptr, masks_ready_ = nullptr; // set_masks_overlapped
    bool masks_pending_ = false;
    // element-wise producers fused into the transforms that consume them (ntt.cuh: NttFuse).  Bit-exact, but measured
    // SLOWER on B200 (16M-4096: 6.38 vs 5.88 ms per query; 256K-512: 0.224 vs 0.209 ms): the prologue runs at the
    // transform's low occupancy and costs a shared-memory pass more than the launch it saves.  Off unless APSU_B200_FUSE=1.
    bool fuse_ = false;
    unsigned fuse_mask_ = 7; // which producers fuse_ covers: 1 extension, 2 tensor product, 4 key-switch inner product
    size_t fuse_max_ = ~size_t(0); // with fuse_: only launches of at most this many polynomials (the latency-bound ones)
    void run_steps(std::vector<Step> &prog, size_t lo, size_t hi, ProgGraph &g);
    void drop_graphs();
    void invalidate_plan()
    {
        plan_valid_ = false;
        drop_graphs();
    }
    void build_plan();
    void prepare_plain_high(BinBundleStore &s);
    void pack_tile(const u64 *src, u64 *dst, uint32_t rows, uint32_t L);
    DBuf<u64> stage_;         // staging for uploads in the standard layout
    DBuf<uint32_t> build_first_, build_size_, build_rows_, build_rows2_; // scratch of add_binbundle_from_bins, kept between calls
    DBuf<u64> build_roots_, build_M_, build_enc_;
    DBuf<int> build_bad_;
    DBuf<u64> aux_[6];        // scratch of the mask / decrypt entry points, kept between calls
    DBuf<int> aux_int_;
    DBuf<uint32_t> aux_idx_;
    DBuf<unsigned char> aux_bytes_;
    // seed expansion (blake2.cuh) and query validation
    DBuf<unsigned char> seed_buf_;
    DBuf<uint32_t> seed_dst_, rej_; // rej_: [count per polynomial][positions]
    DBuf<int> query_bad_;           // [0] residue out of range, [1] rejection-list overflow, [2] peer barrier timeout
    bool query_checked_ = false;
    int *flags_host_ = nullptr; // pinned landing zone of query_bad_
    void enqueue_flag_read();
    void check_flags_after_sync();
    std::vector<uint32_t> partial_rank_; // sorted rank of every source power of a partially loaded query
    std::vector<uint32_t> check_query_powers(const uint32_t *src_powers, uint32_t nsrc);
    void check_range(const u64 *base, uint32_t n_polys, const uint64_t *moduli, uint32_t nmods);
    void expand_seeds(uint32_t L, const std::vector<uint32_t> &dst, const uint8_t *seeds64, u64 *base, const uint64_t *moduli);
    int split_ = 30;          // bit position the DB-stream operands are split at
    uint32_t fold_stages_ = 1; // ring stages between lane folds in the DB-stream kernel
    uint32_t kt_grid_cap_ = 1; // resident CTAs of the DB-stream kernel on this device (persistent grid)
    size_t add_desc(const void *data, size_t bytes);
    void emit_mac(ProgramBuilder &pb, uint32_t L, std::vector<KtGroup> &groups, uint64_t bytes);
    void emit_mul_terms(ProgramBuilder &pb, uint32_t L, std::vector<MulTermsJob> &jobs, uint32_t nterms);
    void emit_finalize(ProgramBuilder &pb, uint32_t Ls, std::vector<FinalizeJob> &jobs);
    template <typename Build>
    void run_scratch_program(Build &&build);

    // batched primitives used by the program builder and the op_* API (index arrays on device)
    void run_extend(uint32_t L, uint32_t n_polys, const uint32_t *src, const uint32_t *dst);
    void run_tensor(uint32_t L, uint32_t n_ops, const uint32_t *a, const uint32_t *b, const uint32_t *d);
    void run_scale_down(uint32_t L, uint32_t n_polys, const uint32_t *src, const uint32_t *dst);
    void run_ks_mac(uint32_t L, uint32_t n_ops, const uint32_t *dig, const uint32_t *out);
    void run_ks_moddown(uint32_t L, uint32_t n_ops, const uint32_t *acc, const uint32_t *ct, const uint32_t *dst, bool mirror = false);
    void run_peer_barrier();
    struct P2P {
        std::vector<u64 *> arena;       // arena base of every rank of the PowersDag partition (peer-mapped)
        std::vector<uint32_t *> flags;  // their barrier flag arrays
        int me = 0;
        bool enabled = false;
    } p2p_;
    DBuf<uint32_t> p2p_flags_; // [kMaxPeers + 1] epochs published by the peers, then this rank's own epoch counter
    void run_mod_switch_next(uint32_t L, uint32_t n_polys, const uint32_t *src, const uint32_t *dst);
    void note_launch() { ctx.launches++; }
};

} // namespace apsu_b200
