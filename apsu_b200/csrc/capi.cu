// C ABI of libapsu_b200.so (include/apsu_b200.h).  Thin: argument checks, exception -> status mapping.
#include "../../include/apsu_b200.h"
#include "engine.hpp"
#include <cstring>
#include <string>

using namespace apsu_b200;

struct apsu_b200_ctx {
    std::unique_ptr<Engine> eng;
};

namespace {
thread_local std::string g_last_error;

template <typename F>
int guarded(F &&f)
{
    try {
        f();
        return APSU_B200_OK;
    } catch (const CudaError &e) {
        g_last_error = e.what();
        return APSU_B200_ERR_CUDA;
    } catch (const std::invalid_argument &e) {
        g_last_error = e.what();
        return APSU_B200_ERR_INVALID_ARGUMENT;
    } catch (const std::out_of_range &e) {
        g_last_error = e.what();
        return APSU_B200_ERR_INVALID_ARGUMENT;
    } catch (const std::logic_error &e) {
        g_last_error = e.what();
        return APSU_B200_ERR_LOGIC;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return APSU_B200_ERR_RUNTIME;
    }
}
Engine &E(apsu_b200_ctx *c)
{
    if (!c || !c->eng) throw std::invalid_argument("context is null");
    APSU_CUDA_CHECK(cudaSetDevice(c->eng->ctx.device));
    return *c->eng;
}
const Engine &E(const apsu_b200_ctx *c)
{
    if (!c || !c->eng) throw std::invalid_argument("context is null");
    return *c->eng;
}
template <typename T>
T *need(T *p, const char *what)
{
    if (!p) throw std::invalid_argument(std::string(what) + " is null");
    return p;
}
} // namespace

namespace apsu_b200 {
// for the other translation units of the C ABI (mgpu.cu)
Engine &engine_of(apsu_b200_ctx *ctx) { return E(ctx); }
int guarded_call(const std::function<void()> &f) { return guarded(f); }
} // namespace apsu_b200

extern "C" {

const char *apsu_b200_last_error(void) { return g_last_error.c_str(); }
const char *apsu_b200_version(void) { return "apsu_b200 0.1 (sm_100a)"; }

int apsu_b200_params_load_json(const char *json_text, apsu_b200_params *out)
{
    return guarded([&] { params_load_json(need(json_text, "json_text"), *need(out, "out")); });
}
int apsu_b200_params_validate(apsu_b200_params *params)
{
    return guarded([&] { params_validate(*need(params, "params")); });
}
int apsu_b200_coeff_modulus_create(uint32_t N, const int *bit_sizes, uint32_t count, uint64_t *out)
{
    return guarded([&] {
        auto v = coeff_modulus_create(N, std::vector<int>(need(bit_sizes, "bit_sizes"), bit_sizes + count));
        std::copy(v.begin(), v.end(), need(out, "out"));
    });
}
int apsu_b200_plain_modulus_batching(uint32_t N, int bit_size, uint64_t *out)
{
    return guarded([&] { *need(out, "out") = plain_modulus_batching(N, bit_size); });
}
int apsu_b200_powers_dag(
    const apsu_b200_params *params, uint32_t capacity, uint32_t *power, uint32_t *depth, uint32_t *parent1,
    uint32_t *parent2, uint32_t *count, uint32_t *dag_depth)
{
    return guarded([&] {
        const apsu_b200_params &p = *need(params, "params");
        PowersDag pd;
        std::set<uint32_t> src(p.query_powers, p.query_powers + p.query_power_count);
        if (!pd.configure(src, create_powers_set(p.ps_low_degree, p.max_items_per_bin)))
            throw std::invalid_argument("failed to configure PowersDag");
        uint32_t n = 0;
        for (uint32_t e : pd.target_powers()) {
            if (n >= capacity) throw std::invalid_argument("capacity is too small");
            const PowersNode &nd = pd.node(e);
            if (power) power[n] = nd.power;
            if (depth) depth[n] = nd.depth;
            if (parent1) parent1[n] = nd.parent1;
            if (parent2) parent2[n] = nd.parent2;
            n++;
        }
        if (count) *count = n;
        if (dag_depth) *dag_depth = pd.depth();
    });
}

int apsu_b200_ctx_create(const apsu_b200_params *params, int device, apsu_b200_ctx **out)
{
    return guarded([&] {
        apsu_b200_params p = *need(params, "params");
        params_validate(p);
        need(out, "out");
        auto c = std::make_unique<apsu_b200_ctx>();
        c->eng = std::make_unique<Engine>(p, device);
        *out = c.release();
    });
}
void apsu_b200_ctx_destroy(apsu_b200_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->eng) {
        cudaSetDevice(ctx->eng->ctx.device);
        cudaStreamSynchronize(ctx->eng->ctx.stream);
    }
    delete ctx;
}
/* pinned host memory for query / result buffers (cudaHostAlloc): host<->device copies from pageable memory are staged
 * and run at a fraction of the PCIe rate */
int apsu_b200_host_alloc(size_t bytes, void **out)
{
    return guarded([&] { APSU_CUDA_CHECK(cudaHostAlloc(need(out, "out"), bytes ? bytes : 1, cudaHostAllocDefault)); });
}
void apsu_b200_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}
int apsu_b200_ctx_set_stream(apsu_b200_ctx *ctx, void *cuda_stream)
{
    return guarded([&] {
        Engine &e = E(ctx);
        APSU_CUDA_CHECK(cudaStreamSynchronize(e.ctx.stream));
        if (e.ctx.owns_stream && e.ctx.stream) cudaStreamDestroy(e.ctx.stream);
        e.ctx.stream = (cudaStream_t)cuda_stream;
        e.ctx.owns_stream = false;
    });
}
int apsu_b200_ctx_get_stream(apsu_b200_ctx *ctx, void **cuda_stream)
{
    return guarded([&] { *need(cuda_stream, "cuda_stream") = (void *)E(ctx).ctx.stream; });
}
int apsu_b200_ctx_synchronize(apsu_b200_ctx *ctx)
{
    return guarded([&] { APSU_CUDA_CHECK(cudaStreamSynchronize(E(ctx).ctx.stream)); });
}
int apsu_b200_ctx_level(const apsu_b200_ctx *ctx, int which, uint32_t *num_primes)
{
    return guarded([&] {
        const Engine &e = E(ctx);
        need(num_primes, "num_primes");
        switch (which) {
        case 0: *num_primes = e.ctx.first_L; break;
        case 1: *num_primes = e.ctx.low_L; break;
        case 2: *num_primes = e.ctx.high_L; break;
        case 3: *num_primes = e.ctx.K; break;
        default: throw std::invalid_argument("unknown level selector");
        }
    });
}

int apsu_b200_db_add_binbundle(apsu_b200_ctx *ctx, uint32_t bundle_idx, const uint64_t *const *coeffs, uint32_t ncoeffs, uint32_t *cache_idx)
{
    return guarded([&] {
        uint32_t ci = E(ctx).add_binbundle(bundle_idx, coeffs, ncoeffs);
        if (cache_idx) *cache_idx = ci;
    });
}
int apsu_b200_db_add_binbundle_synthetic(apsu_b200_ctx *ctx, uint32_t bundle_idx, uint32_t ncoeffs, uint64_t seed, uint32_t *cache_idx)
{
    return guarded([&] {
        uint32_t ci = E(ctx).add_binbundle_synthetic(bundle_idx, ncoeffs, seed);
        if (cache_idx) *cache_idx = ci;
    });
}
int apsu_b200_db_add_binbundle_from_bins(apsu_b200_ctx *ctx, uint32_t bundle_idx, const uint32_t *bin_sizes, const uint64_t *roots, uint32_t *cache_idx)
{
    return guarded([&] {
        uint32_t ci = E(ctx).add_binbundle_from_bins(bundle_idx, need(bin_sizes, "bin_sizes"), need(roots, "roots"));
        if (cache_idx) *cache_idx = ci;
    });
}
int apsu_b200_db_set_data(apsu_b200_ctx *ctx, const uint64_t *felts, const uint64_t *cuckoo_idx, uint64_t n_items, uint32_t *bundle_counts)
{
    return guarded([&] { E(ctx).set_data(felts, cuckoo_idx, (size_t)n_items, false, bundle_counts); });
}
int apsu_b200_db_set_data_device(apsu_b200_ctx *ctx, const void *felts_device, const void *cuckoo_idx_device, uint64_t n_items, uint32_t *bundle_counts)
{
    return guarded([&] { E(ctx).set_data((const uint64_t *)felts_device, (const uint64_t *)cuckoo_idx_device, (size_t)n_items, true, bundle_counts); });
}
int apsu_b200_db_bin_bundle_count(const apsu_b200_ctx *ctx, uint32_t bundle_idx, uint32_t *count)
{
    return guarded([&] {
        const Engine &e = E(ctx);
        if (bundle_idx >= e.db.size()) throw std::out_of_range("bundle_idx is out of range");
        *need(count, "count") = (uint32_t)e.db[bundle_idx].size();
    });
}
int apsu_b200_db_total_bin_bundle_count(const apsu_b200_ctx *ctx, uint32_t *count)
{
    return guarded([&] { *need(count, "count") = E(ctx).total_bundles(); });
}
int apsu_b200_db_binbundle_ncoeffs(const apsu_b200_ctx *ctx, uint32_t bundle_idx, uint32_t cache_idx, uint32_t *ncoeffs)
{
    return guarded([&] {
        const Engine &e = E(ctx);
        if (bundle_idx >= e.db.size() || cache_idx >= e.db[bundle_idx].size()) throw std::out_of_range("no such BinBundle");
        *need(ncoeffs, "ncoeffs") = e.db[bundle_idx][cache_idx]->ncoeffs;
    });
}
int apsu_b200_db_binbundle_coeff(const apsu_b200_ctx *ctx, uint32_t bundle_idx, uint32_t cache_idx, uint32_t k, uint64_t *out, uint32_t *num_primes)
{
    return guarded([&] {
        const Engine &e = E(ctx);
        if (bundle_idx >= e.db.size() || cache_idx >= e.db[bundle_idx].size()) throw std::out_of_range("no such BinBundle");
        const BinBundleStore &s = *e.db[bundle_idx][cache_idx];
        if (k >= s.ncoeffs) throw std::out_of_range("coefficient index is out of range");
        APSU_CUDA_CHECK(cudaSetDevice(e.ctx.device));
        const uint32_t N = e.ctx.N, Ll = e.ctx.low_L;
        uint32_t rank = 0;
        for (uint32_t i = 0; i < k; i++) rank += (e.is_ntt_degree(i) == e.is_ntt_degree(k));
        if (e.is_ntt_degree(k)) {
            if (num_primes) *num_primes = Ll;
            if (out) {
                // device plaintexts are tile-major ([Ll*N/128][n_ntt][128]) and split (db_stream.cuh); hand back
                // plain residues in the standard [Ll][N] order
                const size_t row_bytes = 128 * sizeof(uint64_t);
                APSU_CUDA_CHECK(cudaMemcpy2D(out, row_bytes, s.ntt_coeffs.p + (size_t)rank * 128, (size_t)s.n_ntt * row_bytes, row_bytes, (size_t)Ll * N / 128,
                                             cudaMemcpyDeviceToHost));
                const int sp = e.db_split();
                for (size_t i = 0; i < (size_t)Ll * N; i++) out[i] = (out[i] & 0xFFFFFFFFull) | ((out[i] >> 32) << sp);
            }
        } else {
            if (num_primes) *num_primes = 0;
            if (out) APSU_CUDA_CHECK(cudaMemcpy(out, s.plain_coeffs.p + (size_t)rank * N, (size_t)N * 8, cudaMemcpyDeviceToHost));
        }
    });
}
int apsu_b200_db_stream_bytes(const apsu_b200_ctx *ctx, uint64_t *bytes)
{
    return guarded([&] { *need(bytes, "bytes") = E(ctx).stream_bytes(); });
}
int apsu_b200_db_clear(apsu_b200_ctx *ctx)
{
    return guarded([&] { E(ctx).clear_db(); });
}

int apsu_b200_set_relin_keys(apsu_b200_ctx *ctx, const uint64_t *keys)
{
    return guarded([&] {
        Engine &e = E(ctx);
        e.set_relin_keys(keys, false);
        e.throw_if_query_invalid();
    });
}
int apsu_b200_set_relin_keys_device(apsu_b200_ctx *ctx, const void *keys_device)
{
    return guarded([&] { E(ctx).set_relin_keys(keys_device, true); });
}
int apsu_b200_query_begin(apsu_b200_ctx *ctx, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *cts)
{
    return guarded([&] {
        Engine &e = E(ctx);
        e.query_begin(src_powers, nsrc, cts, false);
        e.throw_if_query_invalid();
    });
}
int apsu_b200_query_begin_device(apsu_b200_ctx *ctx, const uint32_t *src_powers, uint32_t nsrc, const void *cts_device)
{
    return guarded([&] { E(ctx).query_begin(src_powers, nsrc, cts_device, true); });
}
int apsu_b200_compute_powers(apsu_b200_ctx *ctx)
{
    return guarded([&] { E(ctx).compute_powers(); });
}
int apsu_b200_get_power(apsu_b200_ctx *ctx, uint32_t bundle_idx, uint32_t power, uint64_t *out, uint32_t *num_primes, int *is_ntt_form)
{
    return guarded([&] { E(ctx).get_power(bundle_idx, power, out, num_primes, is_ntt_form); });
}
int apsu_b200_set_masks(apsu_b200_ctx *ctx, const uint64_t *masks, uint32_t npack)
{
    return guarded([&] {
        Engine &e = E(ctx);
        e.set_masks(masks, npack, false);
        APSU_CUDA_CHECK(cudaStreamSynchronize(e.ctx.stream));
    });
}
int apsu_b200_set_masks_device(apsu_b200_ctx *ctx, const void *masks_device, uint32_t npack)
{
    return guarded([&] { E(ctx).set_masks(masks_device, npack, true); });
}
int apsu_b200_encode_masks(apsu_b200_ctx *ctx, const uint64_t *slot_values, uint32_t npack, uint64_t *masks_out)
{
    return guarded([&] { E(ctx).encode_masks(slot_values, npack, masks_out); });
}
int apsu_b200_decrypt_results(
    apsu_b200_ctx *ctx, const uint64_t *secret_key_ntt_q0, const uint64_t *cts, uint32_t n, uint64_t *slot_values, uint64_t *blocks, int32_t *noise_budget)
{
    return guarded([&] { E(ctx).decrypt_results(secret_key_ntt_q0, cts, n, slot_values, blocks, noise_budget); });
}
int apsu_b200_set_powers_partition(apsu_b200_ctx *ctx, uint32_t rank, uint32_t size)
{
    return guarded([&] { E(ctx).set_powers_partition(rank, size); });
}
int apsu_b200_powers_stage_count(apsu_b200_ctx *ctx, uint32_t *count)
{
    return guarded([&] { *need(count, "count") = E(ctx).powers_stage_count(); });
}
int apsu_b200_compute_powers_stage(apsu_b200_ctx *ctx, uint32_t stage)
{
    return guarded([&] { E(ctx).compute_powers_stage(stage); });
}
int apsu_b200_powers_exchange_regions(apsu_b200_ctx *ctx, uint32_t level, void **device_ptrs, uint64_t *chunk_bytes, uint32_t capacity, uint32_t *count)
{
    return guarded([&] { *need(count, "count") = E(ctx).powers_exchange_regions(level, device_ptrs, chunk_bytes, capacity); });
}
int apsu_b200_query_begin_seeded(apsu_b200_ctx *ctx, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *c0, const uint8_t *seeds)
{
    return guarded([&] {
        Engine &e = E(ctx);
        e.query_begin_seeded(src_powers, nsrc, c0, seeds);
        e.throw_if_query_invalid();
    });
}
int apsu_b200_set_relin_keys_seeded(apsu_b200_ctx *ctx, const uint64_t *c0, const uint8_t *seeds)
{
    return guarded([&] {
        Engine &e = E(ctx);
        e.set_relin_keys_seeded(c0, seeds);
        e.throw_if_query_invalid();
    });
}
int apsu_b200_generate_masks(apsu_b200_ctx *ctx, const uint8_t *seed, const uint8_t *padded, uint32_t npack, uint64_t *random_matrix, uint64_t *slot_values)
{
    return guarded([&] { E(ctx).generate_masks(seed, padded, npack, random_matrix, slot_values); });
}
int apsu_b200_eval_all(apsu_b200_ctx *ctx)
{
    return guarded([&] { E(ctx).eval_all(); });
}
int apsu_b200_eval_all_stream(apsu_b200_ctx *ctx, uint64_t *out, apsu_b200_result_fn fn, void *user)
{
    return guarded([&] { E(ctx).eval_all_stream(out, fn, user); });
}
int apsu_b200_ctx_set_eval_chunk(apsu_b200_ctx *ctx, uint32_t bin_bundles)
{
    return guarded([&] { E(ctx).set_eval_chunk(bin_bundles); });
}
int apsu_b200_fetch_results(apsu_b200_ctx *ctx, uint64_t *out, uint32_t *bundle_idx, uint32_t *cache_idx)
{
    return guarded([&] { E(ctx).fetch_results(out, bundle_idx, cache_idx); });
}
int apsu_b200_results_device(apsu_b200_ctx *ctx, void **ptr, uint64_t *bytes)
{
    return guarded([&] { E(ctx).results_device(ptr, bytes); });
}
int apsu_b200_copy_results_device(apsu_b200_ctx *ctx, void *dst_device)
{
    return guarded([&] {
        Engine &e = E(ctx);
        void *src = nullptr;
        uint64_t bytes = 0;
        e.results_device(&src, &bytes);
        if (bytes) APSU_CUDA_CHECK(cudaMemcpyAsync(need(dst_device, "dst_device"), src, bytes, cudaMemcpyDeviceToDevice, e.ctx.stream));
    });
}
int apsu_b200_run_query(
    apsu_b200_ctx *ctx, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *cts, const uint64_t *relin_keys,
    const uint64_t *masks, uint32_t npack, uint64_t *out, uint32_t *bundle_idx, uint32_t *cache_idx)
{
    return guarded([&] {
        Engine &e = E(ctx);
        e.query_begin(src_powers, nsrc, cts, false);
        e.set_relin_keys(relin_keys, false);
        e.set_masks_overlapped(masks, npack); // uploaded behind ComputePowers: only the last kernel of the evaluation reads them
        e.compute_powers();
        e.eval_all();
        e.fetch_results(out, bundle_idx, cache_idx);
    });
}

int apsu_b200_run_query_seeded(
    apsu_b200_ctx *ctx, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *c0, const uint8_t *seeds, const uint64_t *relin_c0,
    const uint8_t *relin_seeds, const uint8_t *mask_seed, const uint8_t *padded, uint32_t npack, uint64_t *random_matrix, uint64_t *out,
    uint32_t *bundle_idx, uint32_t *cache_idx)
{
    return guarded([&] {
        Engine &e = E(ctx);
        e.query_begin_seeded(src_powers, nsrc, c0, seeds);
        e.set_relin_keys_seeded(relin_c0, relin_seeds);
        e.generate_masks(mask_seed, padded, npack, random_matrix, nullptr, /*synchronise=*/false);
        e.compute_powers();
        e.eval_all();
        e.fetch_results(out, bundle_idx, cache_idx);
    });
}

int apsu_b200_ctx_modulus_index(const apsu_b200_ctx *ctx, int kind, uint32_t i, uint32_t *index)
{
    return guarded([&] {
        const Engine &e = E(ctx);
        need(index, "index");
        switch (kind) {
        case 0:
            if (i >= e.ctx.K) throw std::out_of_range("coeff modulus index is out of range");
            *index = i;
            break;
        case 1: *index = e.ctx.idx_msk; break;
        case 2:
            if (i >= e.ctx.nB) throw std::out_of_range("auxiliary base index is out of range");
            *index = e.ctx.idx_B0 + i;
            break;
        case 3: *index = e.ctx.idx_t; break;
        default: throw std::invalid_argument("unknown modulus kind");
        }
    });
}
int apsu_b200_op_ntt(apsu_b200_ctx *ctx, uint64_t *polys, uint32_t count, const uint32_t *modulus_index, uint32_t pattern_len, int inverse)
{
    return guarded([&] { E(ctx).op_ntt(polys, count, modulus_index, pattern_len, inverse != 0); });
}
int apsu_b200_op_multiply(apsu_b200_ctx *ctx, uint32_t num_primes, const uint64_t *a, const uint64_t *b, uint64_t *out, uint32_t n_ops)
{
    return guarded([&] { E(ctx).op_multiply(num_primes, a, b, out, n_ops); });
}
int apsu_b200_op_relinearize(apsu_b200_ctx *ctx, uint32_t num_primes, const uint64_t *in, uint64_t *out, uint32_t n_ops)
{
    return guarded([&] { E(ctx).op_relinearize(num_primes, in, out, n_ops); });
}
int apsu_b200_op_mod_switch_next(apsu_b200_ctx *ctx, uint32_t num_primes, const uint64_t *in, uint64_t *out, uint32_t n_polys)
{
    return guarded([&] { E(ctx).op_mod_switch_next(num_primes, in, out, n_polys); });
}

int apsu_b200_op_prng_stream(apsu_b200_ctx *ctx, const uint8_t *seed, uint64_t first_refill, uint64_t *out, uint64_t n_words)
{
    return guarded([&] { E(ctx).op_prng_stream(seed, first_refill, out, (size_t)n_words); });
}
int apsu_b200_op_expand_seeds(apsu_b200_ctx *ctx, uint32_t num_primes, const uint8_t *seeds, uint32_t n, uint64_t *out)
{
    return guarded([&] { E(ctx).op_expand_seeds(num_primes, seeds, n, out); });
}

int apsu_b200_last_timings(apsu_b200_ctx *ctx, apsu_b200_timings *out)
{
    return guarded([&] {
        Engine &e = E(ctx);
        e.collect_timings();
        *need(out, "out") = e.timings;
    });
}
int apsu_b200_bench_ntt(apsu_b200_ctx *ctx, uint32_t count, uint32_t iters, int inverse, float *ms)
{
    return guarded([&] {
        Engine &e = E(ctx);
        if (!count || !iters) throw std::invalid_argument("count and iters must be positive");
        DBuf<u64> buf;
        buf.alloc((size_t)count * e.ctx.N);
        APSU_CUDA_CHECK(cudaMemsetAsync(buf.p, 0, buf.n * 8, e.ctx.stream));
        auto pat = e.ctx.pattern_q(e.ctx.first_L);
        e.ctx.ntt(buf.p, buf.p, count, pat, inverse != 0); // warm-up
        cudaEvent_t a, b;
        APSU_CUDA_CHECK(cudaEventCreate(&a));
        APSU_CUDA_CHECK(cudaEventCreate(&b));
        APSU_CUDA_CHECK(cudaEventRecord(a, e.ctx.stream));
        for (uint32_t i = 0; i < iters; i++) e.ctx.ntt(buf.p, buf.p, count, pat, inverse != 0);
        APSU_CUDA_CHECK(cudaEventRecord(b, e.ctx.stream));
        APSU_CUDA_CHECK(cudaEventSynchronize(b));
        float t = 0;
        APSU_CUDA_CHECK(cudaEventElapsedTime(&t, a, b));
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        *need(ms, "ms") = t / iters;
    });
}
int apsu_b200_set_profiling(apsu_b200_ctx *ctx, int enabled)
{
    return guarded([&] { E(ctx).profiling = enabled != 0; });
}

} // extern "C"
