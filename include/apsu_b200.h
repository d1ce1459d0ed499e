/*
 * apsu_b200 — C ABI of the B200-native receiver-side homomorphic query evaluation of APSU.
 *
 * The reference (real-world-cryprography/APSU) has no FFI/plugin seam on this path; the seam is cut
 * at the C++ calls below.  Every entry point names the reference interface it
 * replaces (paths relative to the reference tree).  All buffers are caller-owned HOST memory unless
 * the name ends in _device; all integers are little-endian uint64_t residues in the layouts SEAL
 * uses at that seam:
 *     ciphertext            uint64_t[size][L][N]      (poly-major, RNS prime, coefficient)
 *     NTT-form plaintext    uint64_t[L][N]            (L = plaintext level of the parameter set)
 *     coefficient plaintext uint64_t[N]               (values mod t)
 *     relinearisation keys  uint64_t[K-1][2][K][N]    (RelinKeys.data()[0][J] = size-2 ct, key level, NTT form)
 *
 * Every function returns 0 on success or a negative apsu_b200_status; apsu_b200_last_error() gives
 * the message of the last failure on the calling thread.  The status classes map 1:1 onto the C++
 * exception types the reference throws (std::invalid_argument / logic_error / runtime_error), which
 * the C++ facade (apsu_b200/host/apsu_b200.hpp) rethrows.  A context is thread-compatible: calls on
 * one context must be serialised by the caller (the facade does).
 *
 * There is NO CPU fallback: without a CUDA device apsu_b200_ctx_create fails with
 * APSU_B200_ERR_CUDA and nothing else can be called.
 */
#ifndef APSU_B200_H
#define APSU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APSU_B200_MAX_COEFF_MODULUS 8
#define APSU_B200_MAX_QUERY_POWERS 1024

typedef enum apsu_b200_status {
    APSU_B200_OK = 0,
    APSU_B200_ERR_INVALID_ARGUMENT = -1, /* std::invalid_argument */
    APSU_B200_ERR_LOGIC = -2,            /* std::logic_error      */
    APSU_B200_ERR_RUNTIME = -3,          /* std::runtime_error    */
    APSU_B200_ERR_CUDA = -4              /* CUDA runtime failure / no device (fails loudly, no fallback) */
} apsu_b200_status;

/* PSUParams — common/apsu/psu_params.h:31-222.  Filled by apsu_b200_params_load_json or by hand and
 * then completed by apsu_b200_params_validate (== PSUParams::initialize, psu_params.cpp:95-180). */
typedef struct apsu_b200_params {
    /* seal_params */
    uint32_t poly_modulus_degree;
    uint32_t coeff_modulus_count; /* K */
    uint64_t plain_modulus;
    uint64_t coeff_modulus[APSU_B200_MAX_COEFF_MODULUS];
    /* table_params / item_params / query_params */
    uint32_t hash_func_count;
    uint32_t table_size;
    uint32_t max_items_per_bin;
    uint32_t felts_per_item;
    uint32_t ps_low_degree;
    uint32_t query_power_count;
    uint32_t query_powers[APSU_B200_MAX_QUERY_POWERS]; /* ascending, always contains 1 */
    /* derived (psu_params.cpp:148-179) */
    uint32_t item_bit_count_per_felt;
    uint32_t item_bit_count;
    uint32_t items_per_bundle;
    uint32_t bins_per_bundle;
    uint32_t bundle_idx_count;
} apsu_b200_params;

typedef struct apsu_b200_ctx apsu_b200_ctx;

const char *apsu_b200_last_error(void);
const char *apsu_b200_version(void);

/* ---- parameters ---------------------------------------------------------------------------- */

/* PSUParams::Load(const std::string&) — common/apsu/psu_params.cpp:290-374 (same JSON schema;
 * CoeffModulus::Create / PlainModulus::Batching prime selection restated). */
int apsu_b200_params_load_json(const char *json_text, apsu_b200_params *out);
/* PSUParams::initialize — common/apsu/psu_params.cpp:95-180. */
int apsu_b200_params_validate(apsu_b200_params *params);
/* seal::CoeffModulus::Create / seal::PlainModulus::Batching (call sites psu_params.cpp:355,363). */
int apsu_b200_coeff_modulus_create(uint32_t poly_modulus_degree, const int *bit_sizes, uint32_t count, uint64_t *out);
int apsu_b200_plain_modulus_batching(uint32_t poly_modulus_degree, int bit_size, uint64_t *out);

/* PowersDag::configure(query_powers, create_powers_set(ps_low_degree, max_items_per_bin)) —
 * common/apsu/powers.cpp:22-107, common/apsu/util/utils.cpp:146-177, call site query.cpp:78.
 * Writes one entry per target power (ascending); p1==p2==0 marks a source node. */
int apsu_b200_powers_dag(
    const apsu_b200_params *params, uint32_t capacity, uint32_t *power, uint32_t *depth, uint32_t *parent1,
    uint32_t *parent2, uint32_t *count, uint32_t *dag_depth);

/* ---- context (CryptoContext + SEALContext + Evaluator) ------------------------------------- */

/* CryptoContext(SEALContext(parms, true, tc128)) + set_evaluator — common/apsu/crypto_context.h:28-125,
 * ReceiverDB ctor receiver/apsu/receiver_db.cpp:640-674.  Builds the modulus chain, NTT tables and
 * BEHZ constants on `device` (a CUDA ordinal). */
int apsu_b200_ctx_create(const apsu_b200_params *params, int device, apsu_b200_ctx **out);
void apsu_b200_ctx_destroy(apsu_b200_ctx *ctx);
/* Pinned host memory for query / result buffers (host<->device copies from pageable memory run at a fraction of the
 * PCIe rate). */
int apsu_b200_host_alloc(size_t bytes, void **out);
void apsu_b200_host_free(void *p);
/* Run all work of this context on an existing cudaStream_t (e.g. torch's current stream).  NULL (the legacy default
 * stream) is accepted but cannot be captured into CUDA graphs: the context then launches its kernels one by one. */
int apsu_b200_ctx_set_stream(apsu_b200_ctx *ctx, void *cuda_stream);
/* The cudaStream_t the context runs on (its own non-blocking stream unless one was set): anything that touches the
 * context's device buffers from another stream (e.g. an external collective on apsu_b200_powers_exchange_regions)
 * must order itself with this stream in both directions. */
int apsu_b200_ctx_get_stream(apsu_b200_ctx *ctx, void **cuda_stream);
int apsu_b200_ctx_synchronize(apsu_b200_ctx *ctx);
/* number of RNS primes of: 0 = first data level, 1 = DB plaintexts / low powers, 2 = high powers,
 * 3 = key level (get_parms_id_for_chain_idx, common/apsu/util/utils.cpp:179-189). */
int apsu_b200_ctx_level(const apsu_b200_ctx *ctx, int which, uint32_t *num_primes);

/* ---- ReceiverDB: device-resident BinBundle caches ------------------------------------------- */

/* One BatchedPlaintextPolyn (receiver/apsu/bin_bundle.h:52-134): coeffs[k] is the k-th entry of
 * batched_coeffs with the SEAL header stripped — NTT form uint64_t[Lp][N] unless k==0 (ps_low_degree==0)
 * or k % (ps_low_degree+1)==0 (ps_low_degree>0), in which case coefficient form uint64_t[N]
 * (bin_bundle.cpp:413).  Appends a BinBundle at bundle_idx (ReceiverDB::bin_bundles_[bundle_idx],
 * receiver_db.h:375) and returns its cache index. */
int apsu_b200_db_add_binbundle(
    apsu_b200_ctx *ctx, uint32_t bundle_idx, const uint64_t *const *coeffs, uint32_t ncoeffs, uint32_t *cache_idx);
/* Synthetic BinBundle for throughput runs: every plaintext word uniform in [0,q_j) / [0,t), generated on
 * the device from a counter-based splitmix64 stream (same stream as the oracle's synthetic fill). */
int apsu_b200_db_add_binbundle_synthetic(
    apsu_b200_ctx *ctx, uint32_t bundle_idx, uint32_t ncoeffs, uint64_t seed, uint32_t *cache_idx);
/* Row f1 — BinBundle::regen_cache (bin_bundle.cpp:934-1041): build the cache on the device from
 * raw bins: polyn_with_roots per bin (interpolate.cpp:63-80), column gather, BatchEncoder::encode,
 * transform_to_ntt (bin_bundle.cpp:366-430); the plaintexts are bit-identical to the ones the reference
 * builds.  bin_sizes[bins_per_bundle] (each below max_items_per_bin, receiver_db.cpp:388-389), roots = the
 * bins' field elements concatenated (each below plain_modulus, else APSU_B200_ERR_INVALID_ARGUMENT). */
int apsu_b200_db_add_binbundle_from_bins(
    apsu_b200_ctx *ctx, uint32_t bundle_idx, const uint32_t *bin_sizes, const uint64_t *roots, uint32_t *cache_idx);
/* Row f1 — ReceiverDB::set_data / insert_or_assign on an empty DB (receiver/apsu/receiver_db.cpp:966, dispatch :446,
 * insert_or_assign_worker :330-438) followed by generate_caches (:808): the whole DB build from the algebraised items,
 * on the device.  Input = the reference's `data_with_indices` (preprocess_unlabeled_data, :246-307): item k has
 * felts[k][0..felts_per_item) (each below plain_modulus) and cuckoo_idx[k] = location * felts_per_item, the first bin of
 * the table slot it hashed to, in the order the reference would walk them (the order decides which BinBundle an item
 * lands in: every item goes to the NEWEST BinBundle of its bundle index whose bins stay below max_items_per_bin, a new
 * one is appended when none has room, :370-432).  The resulting BinBundles — items per bin and all cached plaintexts —
 * are those of the reference, bit for bit.  Replaces the DB.  bundle_counts (optional): [bundle_idx_count] BinBundles
 * created per bundle index. */
int apsu_b200_db_set_data(apsu_b200_ctx *ctx, const uint64_t *felts, const uint64_t *cuckoo_idx, uint64_t n_items, uint32_t *bundle_counts);
int apsu_b200_db_set_data_device(apsu_b200_ctx *ctx, const void *felts_device, const void *cuckoo_idx_device, uint64_t n_items, uint32_t *bundle_counts);
/* ReceiverDB::get_bin_bundle_count(bundle_idx) / () — receiver_db.cpp:742-760. */
int apsu_b200_db_bin_bundle_count(const apsu_b200_ctx *ctx, uint32_t bundle_idx, uint32_t *count);
int apsu_b200_db_total_bin_bundle_count(const apsu_b200_ctx *ctx, uint32_t *count);
/* batched_coeffs.size() of one BinBundle and a copy of one plaintext back to the host (tests). */
int apsu_b200_db_binbundle_ncoeffs(const apsu_b200_ctx *ctx, uint32_t bundle_idx, uint32_t cache_idx, uint32_t *ncoeffs);
int apsu_b200_db_binbundle_coeff(
    const apsu_b200_ctx *ctx, uint32_t bundle_idx, uint32_t cache_idx, uint32_t k, uint64_t *out, uint32_t *num_primes);
/* bytes of DB plaintext the evaluation stream reads per query (SURVEY.md §8d algorithmic bytes). */
int apsu_b200_db_stream_bytes(const apsu_b200_ctx *ctx, uint64_t *bytes);
int apsu_b200_db_clear(apsu_b200_ctx *ctx);

/* ---- query ---------------------------------------------------------------------------------- */

/* CryptoContext::set_evaluator(query.relin_keys()) — receiver/apsu/receiver_ddh.cpp:175-176. */
int apsu_b200_set_relin_keys(apsu_b200_ctx *ctx, const uint64_t *keys);
/* Load Q_i^e into all_powers[bundle_idx][e] — receiver_ddh.cpp:295-322.
 * cts: uint64_t[nsrc][bundle_idx_count][2][L_first][N], coefficient form; src_powers must equal the
 * parameter set's query_powers (Query validation, receiver/apsu/query.cpp:68-111). */
int apsu_b200_query_begin(apsu_b200_ctx *ctx, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *cts);
/* Receiver::ComputePowers for every bundle index — receiver_ddh.cpp:325-333, 390-483. */
int apsu_b200_compute_powers(apsu_b200_ctx *ctx);
/* Debug/parity: copy all_powers[bundle_idx][power] (size-2 ciphertext) to the host. */
int apsu_b200_get_power(
    apsu_b200_ctx *ctx, uint32_t bundle_idx, uint32_t power, uint64_t *out, uint32_t *num_primes, int *is_ntt_form);
/* Masks: dense table uint64_t[alpha_max_cache_count][bundle_idx_count][N] of coefficient-form
 * plaintexts, indexed pack_idx = bundle_idx + cache_idx*bundle_idx_count (receiver_ddh.cpp:346;
 * SURVEY.md Appendix C.1).  apsu_b200_encode_masks does BatchEncoder::encode (receiver_ddh.cpp:275)
 * on the device from slot values uint64_t[npack][N]. */
int apsu_b200_set_masks(apsu_b200_ctx *ctx, const uint64_t *masks, uint32_t npack);
int apsu_b200_encode_masks(apsu_b200_ctx *ctx, const uint64_t *slot_values, uint32_t npack, uint64_t *masks_out);
/* "next" row f4 — the sender's side of the exchange for a batch of results: ResultPackage::extract
 * (common/apsu/network/result_package.cpp:175-213: Decryptor::decrypt, invariant_noise_budget, BatchEncoder::decode)
 * and the items' 128-bit blocks (vec_to_std_block, sender/apsu/sender_ddh.cpp:588-594).  cts: [n][2][N] result
 * ciphertexts at the last level (what apsu_b200_fetch_results returns); secret_key_ntt_q0: the secret key in NTT form
 * modulo the first coefficient prime (the first N words of seal::SecretKey::data()); slot_values: [n][N];
 * blocks (optional): [n][items_per_bundle][2] (low, high); noise_budget (optional): [n] bits. */
int apsu_b200_decrypt_results(
    apsu_b200_ctx *ctx, const uint64_t *secret_key_ntt_q0, const uint64_t *cts, uint32_t n, uint64_t *slot_values, uint64_t *blocks,
    int32_t *noise_budget);
/* Multi-GPU, collective C2 of SURVEY.md §8e: when more GPUs than bundle indices serve a DB, the ranks that share a
 * bundle index split the PowersDag of ComputePowers (receiver_ddh.cpp:390-483) instead of each recomputing it.
 * Rank r of `size` computes chunk r of every DAG level (ceil(n/size) consecutive nodes of PowersDag's level order);
 * stage s < dag_depth of apsu_b200_compute_powers_stage runs level s+1, after which the caller all-gathers that
 * level's exchange regions between the ranks (one region per populated bundle index: a device buffer of `size`
 * chunks of chunk_bytes, this rank's chunk at rank*chunk_bytes — e.g. ncclAllGather in place); the last stage
 * (mod-switches, NTTs, power tables) needs every power and runs on every rank.  size == 1 (default) keeps
 * apsu_b200_compute_powers as one call. */
int apsu_b200_set_powers_partition(apsu_b200_ctx *ctx, uint32_t rank, uint32_t size);
int apsu_b200_powers_stage_count(apsu_b200_ctx *ctx, uint32_t *count);
int apsu_b200_compute_powers_stage(apsu_b200_ctx *ctx, uint32_t stage);
int apsu_b200_powers_exchange_regions(
    apsu_b200_ctx *ctx, uint32_t level, void **device_ptrs, uint64_t *chunk_bytes, uint32_t capacity, uint32_t *count);
/* Row f3 — the mask generation of RunQuery on the device (receiver/apsu/receiver_ddh.cpp:218-283): SEAL's default
 * generator (Blake2xbPRNG) keyed with the 64 bytes of `seed`; seed == NULL draws them from the OS (getrandom) exactly
 * as the reference does with random_bytes (:221-225) — that is the production setting; a caller-supplied seed makes
 * masks (and with them result ciphertexts) reproducible for tests.  The pairs are visited in (cache_idx, bundle_idx)
 * order, i.e. ascending pack index p = bundle_idx + cache_idx*bundle_idx_count, padded pairs (padded[p] != 0) draw
 * nothing and get Block::all_one_block (:247-252); every other pair draws N 32-bit words, r = word % plain_modulus per
 * slot (:256-262), BatchEncoder::encode of it (:275; kept device-resident as the masks of the next evaluation, like
 * apsu_b200_set_masks) and random_matrix[p][item] = vec_to_std_block of the item's felts (:70-92, :263-270) as
 * (low, high) words.  random_matrix: [npack][items_per_bundle][2]; slot_values (optional, tests): [npack][N].
 * [SEAL-RECALL] Blake2xbPRNG: 4096-byte refills blake2xb(key = seed, input = 64-bit refill counter). */
int apsu_b200_generate_masks(
    apsu_b200_ctx *ctx, const uint8_t *seed /*[64] or NULL*/, const uint8_t *padded, uint32_t npack, uint64_t *random_matrix, uint64_t *slot_values);
/* Row f2 — the query as it is on the wire (seal::Serializable<Ciphertext>: sender/apsu/plaintext_powers.cpp:45,
 * common/apsu/seal_object.h:161-219, common/apsu/network/receiver_operation.cpp:187-345): the second polynomial of a
 * symmetric-key ciphertext is replaced by the 64-byte seed of the generator that sampled it, and
 * Ciphertext::unsafe_load re-samples it (expand_seed -> sample_poly_uniform).  The expansion runs on the device, so
 * only c0 and the seeds are uploaded (half the bytes).  c0: uint64_t[nsrc][bundle_idx_count][L_first][N];
 * seeds: uint8_t[nsrc][bundle_idx_count][64].  Relinearisation keys likewise (KeyGenerator::create_relin_keys returns
 * Serializable<RelinKeys>): c0 uint64_t[K-1][K][N], seeds uint8_t[K-1][64].  Host-side parsing of the SEAL byte
 * layout: apsu_b200/host/seal_wire.hpp. */
int apsu_b200_query_begin_seeded(apsu_b200_ctx *ctx, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *c0, const uint8_t *seeds);
int apsu_b200_set_relin_keys_seeded(apsu_b200_ctx *ctx, const uint64_t *c0, const uint8_t *seeds);
/* Receiver::ProcessBinBundleCache for every BinBundle — receiver_ddh.cpp:340-369, 485-535 →
 * BatchedPlaintextPolyn::eval / eval_patstock (bin_bundle.cpp:106-174, 192-360).  Results stay on the
 * device until fetched.  Order of results: bundle_idx major, cache_idx minor. */
int apsu_b200_eval_all(apsu_b200_ctx *ctx);
/* The same with delivery PER BinBundle, as the reference hands each ResultPackage to the channel the moment its
 * BinBundle is done (send_rp_fun, receiver_ddh.cpp:527-534): the result ciphertexts are copied to `out`
 * (uint64_t[total_bin_bundle_count][2][N], result order; pinned memory recommended) chunk by chunk while later chunks
 * are still being evaluated, and `fn(user, bundle_idx, cache_idx, ct)` is called on the calling thread for every
 * BinBundle as soon as its ciphertext is on the host.  A chunk is what one finalize launch completes: all directly
 * evaluated BinBundles first, then the Paterson-Stockmeyer ones apsu_b200_ctx_set_eval_chunk at a time (default 32:
 * larger chunks are faster in total, smaller ones deliver earlier). */
typedef void (*apsu_b200_result_fn)(void *user, uint32_t bundle_idx, uint32_t cache_idx, const uint64_t *ct);
int apsu_b200_eval_all_stream(apsu_b200_ctx *ctx, uint64_t *out, apsu_b200_result_fn fn, void *user);
int apsu_b200_ctx_set_eval_chunk(apsu_b200_ctx *ctx, uint32_t bin_bundles);
/* rp->psu_result for every BinBundle: out = uint64_t[total_bin_bundle_count][2][N] (one prime), plus
 * the ResultPackage bundle_idx / cache_idx fields (network/result_package.h:44-62). Either index
 * array may be NULL. */
int apsu_b200_fetch_results(apsu_b200_ctx *ctx, uint64_t *out, uint32_t *bundle_idx, uint32_t *cache_idx);
/* The whole HE part of Receiver::RunQuery (receiver_ddh.cpp:295-369) through host buffers:
 * query_begin + set_relin_keys + set_masks + compute_powers + eval_all + fetch_results. */
int apsu_b200_run_query(
    apsu_b200_ctx *ctx, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *cts, const uint64_t *relin_keys,
    const uint64_t *masks, uint32_t npack, uint64_t *out, uint32_t *bundle_idx, uint32_t *cache_idx);
/* The same, fed as the reference's RunQuery is: the query as it is on the wire (seeded ciphertexts and keys,
 * apsu_b200_query_begin_seeded / apsu_b200_set_relin_keys_seeded) and the masks drawn inside the call
 * (apsu_b200_generate_masks; mask_seed NULL = OS randomness as in receiver_ddh.cpp:221-225).  Uploads c0 + seeds only,
 * returns the result ciphertexts and random_matrix (the PEQT hand-off, [npack][items_per_bundle][2]). */
int apsu_b200_run_query_seeded(
    apsu_b200_ctx *ctx, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *c0, const uint8_t *seeds, const uint64_t *relin_c0,
    const uint8_t *relin_seeds, const uint8_t *mask_seed, const uint8_t *padded, uint32_t npack, uint64_t *random_matrix, uint64_t *out,
    uint32_t *bundle_idx, uint32_t *cache_idx);
/* Device-resident variant used for sharded runs: cts/keys/masks already on this context's GPU.  These calls are
 * ASYNCHRONOUS: they queue device-to-device copies on the context stream and return; the source buffers must stay
 * valid (and unmodified) until apsu_b200_ctx_synchronize or a fetch of the results.  The residue-range check of the
 * query (seal::is_valid_for, receiver/apsu/query.cpp:45-66) runs on the device with every load and is reported by the
 * next synchronising call (the host-buffer variants report it themselves). */
int apsu_b200_query_begin_device(apsu_b200_ctx *ctx, const uint32_t *src_powers, uint32_t nsrc, const void *cts_device);
int apsu_b200_set_relin_keys_device(apsu_b200_ctx *ctx, const void *keys_device);
int apsu_b200_set_masks_device(apsu_b200_ctx *ctx, const void *masks_device, uint32_t npack);
/* device pointer + byte size of the result buffer filled by eval_all (valid until the next eval_all) */
int apsu_b200_results_device(apsu_b200_ctx *ctx, void **ptr, uint64_t *bytes);
/* asynchronous device-to-device copy of the results into a caller-owned device buffer (e.g. a torch tensor
 * that is then gathered with NCCL); ordered on the context stream. */
int apsu_b200_copy_results_device(apsu_b200_ctx *ctx, void *dst_device);

/* ---- multi-GPU: one rank per B200, BinBundles sharded ------------------------------------------ */
/* The reference fans ProcessBinBundleCache out over a thread pool (receiver/apsu/receiver_ddh.cpp:340-364); the unit
 * of that fan-out, the BinBundle, is what shards across GPUs: every rank holds the BinBundles it was given in its own
 * context (apsu_b200_db_add_*), recomputes the query powers of the bundle indices it owns (ComputePowers,
 * receiver_ddh.cpp:325-333) and evaluates its BinBundles.  All exchanges run inside the library, in C++, over NCCL
 * (bound at run time with dlopen("libnccl.so.2"); no link-time dependency): the query goes from the root rank (0) to
 * the ranks that need it (only the ciphertexts of the bundle indices a rank owns; relinearisation keys to all), the
 * result ciphertexts come back to the root unpadded.  When several ranks hold BinBundles of ONE bundle index only,
 * they may split its PowersDag and all-gather every DAG level in place (collective C2).
 * One apsu_b200_mgpu per rank; ranks may be processes (one per GPU) or threads of one process (each with its own
 * context); every function below is COLLECTIVE: all ranks call it, in the same order. */
typedef struct apsu_b200_mgpu apsu_b200_mgpu;
#define APSU_B200_MGPU_ID_BYTES 128
/* ncclGetUniqueId on one rank; hand the 128 bytes to every rank (any transport). */
int apsu_b200_mgpu_unique_id(uint8_t *id /*[128]*/);
int apsu_b200_mgpu_create(apsu_b200_ctx *ctx, const uint8_t *id, uint32_t rank, uint32_t world, apsu_b200_mgpu **out);
void apsu_b200_mgpu_destroy(apsu_b200_mgpu *m);
/* After the rank's BinBundles are loaded (and after every DB change): publishes which BinBundles each rank holds.
 * global_cache_idx[k] = cache index, in the whole DB, of this rank's k-th BinBundle in result order (bundle_idx major,
 * local cache index minor); NULL = the local indices.  dag_split: 1 = ranks sharing one bundle index split its
 * PowersDag, 0 = every rank recomputes it, -1 = split when the levels can be exchanged through NVLink peer memory (see
 * apsu_b200_mgpu_info) or, failing that, only large DAGs (>= 128 products: below that ncclAllGather per level costs more
 * than the products it saves, profiles/README.md). */
int apsu_b200_mgpu_commit(apsu_b200_mgpu *m, const uint32_t *global_cache_idx, int dag_split);
/* dag_exchange: how a split PowersDag is exchanged — 0 not split, 1 ncclAllGather per DAG level, 2 through NVLink peer
 * memory: the key-switch epilogue of every level stores its products straight into the peers' arenas (CUDA IPC between
 * processes, peer access between threads) and a flag barrier on the stream closes the level; chosen automatically at
 * commit when every rank of the group can map the others, APSU_B200_NO_P2P=1 forces the NCCL exchange. */
int apsu_b200_mgpu_info(const apsu_b200_mgpu *m, uint32_t *total_bin_bundles, uint32_t *dag_group_size, int *dag_exchange, int *nccl_version);
/* The HE part of Receiver::RunQuery (receiver_ddh.cpp:295-369) over all ranks, host buffers in and out.  Root passes
 * cts / relin_keys (layouts of apsu_b200_run_query) and receives out = uint64_t[total][2][N] with the ResultPackage
 * indices of every result (rank-major, each rank's results in its result order); the other ranks pass NULL for them.
 * masks_local: this rank's dense mask table for its local cache indices (NULL keeps the resident masks, e.g. after
 * apsu_b200_generate_masks). */
int apsu_b200_mgpu_run_query(
    apsu_b200_mgpu *m, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *cts, const uint64_t *relin_keys, const uint64_t *masks_local,
    uint32_t npack_local, uint64_t *out, uint32_t *bundle_idx, uint32_t *cache_idx);
/* The same when EVERY rank can read the query in host memory (threads of one process, or processes that map one shared
 * segment the receiving process wrote the query into): every rank passes cts / relin_keys and uploads the ciphertexts of
 * its own bundle indices over its own PCIe link — the uploads run in parallel and no scatter / broadcast happens.
 * Results are gathered on the root as above. */
int apsu_b200_mgpu_run_query_shared(
    apsu_b200_mgpu *m, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *cts, const uint64_t *relin_keys, const uint64_t *masks_local,
    uint32_t npack_local, uint64_t *out, uint32_t *bundle_idx, uint32_t *cache_idx);
/* The shared-query call with LOCAL result delivery: every rank copies the result ciphertexts of ITS OWN BinBundles to its
 * own host buffer — out_local uint64_t[local_count][2][N], bundle_idx_local / cache_idx_local [local_count] with the
 * GLOBAL cache indices, in the order of the rank's span of the gathered layout — and nothing is gathered: the
 * device-to-host copies of all GPUs run in parallel.  This is what a host with one thread (or process) per GPU needs when
 * each of them forwards its ResultPackages itself, as the reference's workers do (receiver/apsu/receiver_ddh.cpp:340-364,
 * 527-534: send_rp_fun is called by the pool thread that finished the BinBundle). */
int apsu_b200_mgpu_run_query_local(
    apsu_b200_mgpu *m, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *cts, const uint64_t *relin_keys, const uint64_t *masks_local,
    uint32_t npack_local, uint64_t *out_local, uint32_t *bundle_idx_local, uint32_t *cache_idx_local);
/* number of BinBundles of this rank (after apsu_b200_mgpu_commit) */
int apsu_b200_mgpu_local_count(const apsu_b200_mgpu *m, uint32_t *count);
/* ComputePowers of this rank with the per-level exchange of a split PowersDag (query already loaded). */
int apsu_b200_mgpu_compute_powers(apsu_b200_mgpu *m);

/* ---- SEAL Evaluator calls on the path, as stand-alone batched device operations (K2–K8) ----- */
/* modulus selector for apsu_b200_op_ntt: index into [coeff_modulus[0..K-1], m_sk, B_0.., plain_modulus] */
int apsu_b200_ctx_modulus_index(const apsu_b200_ctx *ctx, int kind /*0=coeff,1=m_sk,2=B,3=plain*/, uint32_t i, uint32_t *index);
/* Evaluator::transform_to_ntt_inplace / transform_from_ntt_inplace on `count` polynomials uint64_t[count][N],
 * polynomial p reduced modulo modulus_index[p % pattern_len]. */
int apsu_b200_op_ntt(apsu_b200_ctx *ctx, uint64_t *polys, uint32_t count, const uint32_t *modulus_index, uint32_t pattern_len, int inverse);
/* Evaluator::multiply / square (size 2 x size 2 -> size 3) at level `num_primes`, n_ops independent products.
 * a,b: uint64_t[n_ops][2][L][N]; out: uint64_t[n_ops][3][L][N]; coefficient form. */
int apsu_b200_op_multiply(apsu_b200_ctx *ctx, uint32_t num_primes, const uint64_t *a, const uint64_t *b, uint64_t *out, uint32_t n_ops);
/* Evaluator::relinearize_inplace (size 3 -> 2): in uint64_t[n_ops][3][L][N], out uint64_t[n_ops][2][L][N]. */
int apsu_b200_op_relinearize(apsu_b200_ctx *ctx, uint32_t num_primes, const uint64_t *in, uint64_t *out, uint32_t n_ops);
/* Evaluator::mod_switch_to_next_inplace on n_polys polynomials uint64_t[n_polys][L][N] -> [n_polys][L-1][N]. */
int apsu_b200_op_mod_switch_next(apsu_b200_ctx *ctx, uint32_t num_primes, const uint64_t *in, uint64_t *out, uint32_t n_polys);

/* SEAL's Blake2xbPRNG as a raw stream: n_words 64-bit words starting at refill `first_refill` (4096 bytes each), and
 * sample_poly_uniform for n seeded polynomials over coeff_modulus[0..num_primes-1]: out uint64_t[n][num_primes][N]. */
int apsu_b200_op_prng_stream(apsu_b200_ctx *ctx, const uint8_t *seed /*[64]*/, uint64_t first_refill, uint64_t *out, uint64_t n_words);
int apsu_b200_op_expand_seeds(apsu_b200_ctx *ctx, uint32_t num_primes, const uint8_t *seeds /*[n][64]*/, uint32_t n, uint64_t *out);

/* ---- measurement ----------------------------------------------------------------------------- */
typedef struct apsu_b200_timings {
    float compute_powers_ms;  /* "Receiver::ComputePowers" scope, device time on the context stream */
    float eval_ms;            /* all "Receiver::ProcessBinBundleCache" scopes */
    float db_stream_ms;       /* sum of DB-stream MAC kernel launches inside eval */
    uint64_t db_stream_bytes; /* algorithmic plaintext bytes those launches covered */
    uint32_t db_stream_launches;
    uint32_t kernel_launches; /* kernels launched by the last compute_powers + eval_all */
} apsu_b200_timings;
int apsu_b200_last_timings(apsu_b200_ctx *ctx, apsu_b200_timings *out);
/* NTT micro-benchmark on device-resident data: `count` polynomials over the first-level primes, `iters`
 * back-to-back launches timed with CUDA events; *ms = average per launch.  (BASELINE metric "NTT GB/s":
 * 16*N bytes per polynomial per transform.) */
int apsu_b200_bench_ntt(apsu_b200_ctx *ctx, uint32_t count, uint32_t iters, int inverse, float *ms);
/* when enabled, DB-stream launches are individually timed with CUDA events (adds event overhead only) */
int apsu_b200_set_profiling(apsu_b200_ctx *ctx, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* APSU_B200_H */
