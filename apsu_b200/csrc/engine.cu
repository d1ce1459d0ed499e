#include "engine.hpp"
#include "blake2.cuh"
#include "db_stream.cuh"
#include "eval_kernels.cuh"
#include "hostmath.hpp"
#include "kernels.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <array>
#include <map>
#include <sys/random.h>

namespace apsu_b200 {

// ------------------------------------------------------------------------------------------------
// kernel launch wrappers
// ------------------------------------------------------------------------------------------------
#define APSU_LAUNCH_CHECK()                                                                                            \
    do {                                                                                                               \
        APSU_CUDA_CHECK(cudaGetLastError());                                                                           \
        note_launch();                                                                                                 \
    } while (0)

// specialised (|q|, |Bsk|) instances of the BEHZ kernels; anything else runs the generic <0,0> instance
#define APSU_DISPATCH_LS(KERN, L, S, GRID, ...)                                                                        \
    do {                                                                                                               \
        if ((L) == 1 && (S) == 2) launch_pdl(KERN<1, 2>, GRID, dim3(kEwThreads), 0, ctx.stream, __VA_ARGS__);          \
        else if ((L) == 2 && (S) == 3) launch_pdl(KERN<2, 3>, GRID, dim3(kEwThreads), 0, ctx.stream, __VA_ARGS__);     \
        else if ((L) == 3 && (S) == 4) launch_pdl(KERN<3, 4>, GRID, dim3(kEwThreads), 0, ctx.stream, __VA_ARGS__);     \
        else if ((L) == 4 && (S) == 5) launch_pdl(KERN<4, 5>, GRID, dim3(kEwThreads), 0, ctx.stream, __VA_ARGS__);     \
        else launch_pdl(KERN<0, 0>, GRID, dim3(kEwThreads), 0, ctx.stream, __VA_ARGS__);                               \
    } while (0)

void Engine::run_extend(uint32_t L, uint32_t n, const uint32_t *src, const uint32_t *dst)
{
    if (!n) return;
    APSU_DISPATCH_LS(k_behz_extend, L, (uint32_t)ctx.level[L].S, dim3(ctx.N / kEwThreads, n), arena_.buf.p, src, dst, ctx.level[L], (int)ctx.N);
    APSU_LAUNCH_CHECK();
}
void Engine::run_tensor(uint32_t L, uint32_t n_ops, const uint32_t *a, const uint32_t *b, const uint32_t *d)
{
    if (!n_ops) return;
    const LevelConsts &c = ctx.level[L];
    launch_pdl(k_tensor, dim3(ctx.N / kEwThreads, c.L + c.S, n_ops), dim3(kEwThreads), 0, ctx.stream, arena_.buf.p, a, b, d, c, (int)ctx.N);
    APSU_LAUNCH_CHECK();
}
void Engine::run_scale_down(uint32_t L, uint32_t n, const uint32_t *src, const uint32_t *dst)
{
    if (!n) return;
    APSU_DISPATCH_LS(k_behz_scale_down, L, (uint32_t)ctx.level[L].S, dim3(ctx.N / kEwThreads, n), arena_.buf.p, src, dst, ctx.level[L], (int)ctx.N);
    APSU_LAUNCH_CHECK();
}
void Engine::run_ks_mac(uint32_t L, uint32_t n_ops, const uint32_t *dig, const uint32_t *out)
{
    if (!n_ops) return;
    launch_pdl(k_ks_mac, dim3(ctx.N / kEwThreads, L + 1, n_ops), dim3(kEwThreads), 0, ctx.stream, arena_.buf.p, dig, out, (const u64 *)relin_keys_.p, ctx.ks[L], (int)ctx.N);
    APSU_LAUNCH_CHECK();
}
void Engine::run_ks_moddown(uint32_t L, uint32_t n_ops, const uint32_t *acc, const uint32_t *ct, const uint32_t *dst, bool mirror)
{
    if (!n_ops) return;
    PeerArenas pa;
    std::memset(&pa, 0, sizeof(pa));
    if (mirror && p2p_.enabled) {
        for (size_t k = 0; k < p2p_.arena.size(); k++)
            if ((int)k != p2p_.me) pa.base[pa.n++] = p2p_.arena[k];
        launch_pdl(k_ks_moddown<true>, dim3(ctx.N / kEwThreads, 2, n_ops), dim3(kEwThreads), 0, ctx.stream, arena_.buf.p, acc, ct, dst, ctx.ks[L], (int)ctx.N, pa);
    } else {
        launch_pdl(k_ks_moddown<false>, dim3(ctx.N / kEwThreads, 2, n_ops), dim3(kEwThreads), 0, ctx.stream, arena_.buf.p, acc, ct, dst, ctx.ks[L], (int)ctx.N, pa);
    }
    APSU_LAUNCH_CHECK();
}
// barrier between the ranks that split the PowersDag over peer memory (no-op unless set_powers_p2p enabled it)
void Engine::run_peer_barrier()
{
    if (!p2p_.enabled) return;
    PeerFlags f;
    std::memset(&f, 0, sizeof(f));
    f.n = (int)p2p_.flags.size();
    f.me = p2p_.me;
    for (int k = 0; k < f.n; k++) f.peer[k] = p2p_.flags[k];
    f.mine = p2p_flags_.p;
    f.epoch = p2p_flags_.p + kMaxPeers + 1;
    f.error = query_bad_.p + 2;
    k_xgpu_barrier<<<1, 32, 0, ctx.stream>>>(f, 4000000000ll); // ~2 s at 2 GHz
    APSU_LAUNCH_CHECK();
    query_checked_ = true;
}
uint32_t *Engine::p2p_flags()
{
    if (!p2p_flags_.p) {
        p2p_flags_.alloc(kMaxPeers + 2);
        APSU_CUDA_CHECK(cudaMemsetAsync(p2p_flags_.p, 0, (kMaxPeers + 2) * sizeof(uint32_t), ctx.stream));
        APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    }
    return p2p_flags_.p;
}
u64 *Engine::arena_base()
{
    if (!plan_valid_) build_plan();
    return arena_.buf.p;
}
// peer arenas / flag arrays of the ranks of the PowersDag partition, indexed by partition rank (own entries ignored);
// empty vectors switch the peer-memory exchange off (the caller all-gathers the levels itself)
void Engine::set_powers_p2p(const std::vector<u64 *> &arenas, const std::vector<uint32_t *> &flags)
{
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    if (arenas.size() > (size_t)kMaxPeers + 1 || arenas.size() != flags.size()) throw std::invalid_argument("set_powers_p2p: bad peer tables");
    p2p_.arena = arenas;
    p2p_.flags = flags;
    p2p_.me = (int)powers_part_rank_;
    p2p_.enabled = arenas.size() > 1 && arenas.size() == powers_part_size_;
    drop_graphs(); // the captured launches carry the peer pointers
}
void Engine::run_mod_switch_next(uint32_t L, uint32_t n, const uint32_t *src, const uint32_t *dst)
{
    if (!n) return;
    launch_pdl(k_mod_switch_next, dim3(ctx.N / kEwThreads, n), dim3(kEwThreads), 0, ctx.stream, arena_.buf.p, src, dst, ctx.level[L], (int)ctx.N);
    APSU_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------------
// Program builder: turns batches of ciphertext operations into kernel launches over index arrays.
// ------------------------------------------------------------------------------------------------
struct ProgramBuilder {
    Engine &e;
    std::vector<Engine::Step> &prog;
    DeviceContext &ctx;
    Arena &arena;
    IdxPool &idx;
    ProgramBuilder(Engine &eng, std::vector<Engine::Step> &p) : e(eng), prog(p), ctx(eng.ctx), arena(eng.arena_), idx(eng.idx_) {}

    void step(std::function<void()> f) { prog.push_back(Engine::Step{ std::move(f) }); }

    // NTT over `count` polynomials with explicit source/destination indices
    void ntt(const std::vector<uint32_t> &src, const std::vector<uint32_t> &dst, const std::vector<uint32_t> &pattern, bool inverse, bool reduce = false)
    {
        if (src.empty()) return;
        size_t so = idx.add(src), dn = idx.add(dst);
        uint32_t count = (uint32_t)src.size();
        Engine *en = &e;
        step([=] { en->ctx.ntt(en->arena_.buf.p, en->arena_.buf.p, count, pattern, inverse, en->idx_.at(so), en->idx_.at(dn), reduce); });
    }
    // in-place NTT over a contiguous run of polynomials
    void ntt_run(uint32_t first, uint32_t count, const std::vector<uint32_t> &pattern, bool inverse)
    {
        if (!count) return;
        Engine *en = &e;
        step([=] {
            u64 *base = en->arena_.buf.p + (size_t)first * en->ctx.N;
            en->ctx.ntt(base, base, count, pattern, inverse);
        });
    }

    // BEHZ steps (1)-(3): size-2 ciphertexts at level L (arena idx of [2][L][N]) -> extended NTT form
    // entries [2][L+S][N]
    void extend(uint32_t L, const std::vector<uint32_t> &cts, const std::vector<uint32_t> &exts)
    {
        if (cts.empty()) return;
        const uint32_t S = (uint32_t)ctx.level[L].S, LS = L + S;
        std::vector<uint32_t> esrc, edst, nsrc, ndst;
        for (size_t k = 0; k < cts.size(); k++)
            for (uint32_t c = 0; c < 2; c++) {
                esrc.push_back(cts[k] + c * L);
                edst.push_back(exts[k] + c * LS + L);
                for (uint32_t j = 0; j < LS; j++) {
                    nsrc.push_back(j < L ? cts[k] + c * L + j : exts[k] + c * LS + j);
                    ndst.push_back(exts[k] + c * LS + j);
                }
            }
        Engine *en = &e;
        if (e.fuse_ && (e.fuse_mask_ & 1) && ndst.size() <= e.fuse_max_) {
            // one launch: the transform of polynomial (rp, j) computes its own input (ntt.cuh: kNttExtend)
            size_t so = idx.add(esrc), dn = idx.add(ndst);
            const uint32_t count = (uint32_t)ndst.size();
            const std::vector<uint32_t> pat = ctx.pattern_ext(L);
            step([=] {
                NttFuse f{ en->levels_dev_.p + L, nullptr, nullptr, nullptr, (int)L, (int)S, 0 };
                en->ctx.ntt_fused(kNttExtend, en->arena_.buf.p, count, pat, en->idx_.at(so), en->idx_.at(dn), en->arena_.buf.p, f);
            });
            return;
        }
        size_t so = idx.add(esrc), dn = idx.add(edst);
        uint32_t n = (uint32_t)esrc.size();
        step([=] { en->run_extend(L, n, en->idx_.at(so), en->idx_.at(dn)); });
        ntt(nsrc, ndst, ctx.pattern_ext(L), false);
    }

    // BEHZ steps (4)-(8): products of extended entries a[o] x b[o] -> size-3 ciphertexts prod[o] ([3][L][N]).
    // scratch: arena index of n_ops*3*(L+S) polynomials.
    void multiply(uint32_t L, const std::vector<uint32_t> &a, const std::vector<uint32_t> &b, const std::vector<uint32_t> &prod, uint32_t scratch)
    {
        if (a.empty()) return;
        const uint32_t S = (uint32_t)ctx.level[L].S, LS = L + S, n_ops = (uint32_t)a.size();
        std::vector<uint32_t> d(n_ops), ssrc, sdst;
        for (uint32_t o = 0; o < n_ops; o++) {
            d[o] = scratch + o * 3 * LS;
            for (uint32_t c = 0; c < 3; c++) {
                ssrc.push_back(d[o] + c * LS);
                sdst.push_back(prod[o] + c * L);
            }
        }
        size_t ao = idx.add(a), bo = idx.add(b), dof = idx.add(d), so = idx.add(ssrc), sd = idx.add(sdst);
        Engine *en = &e;
        if (e.fuse_ && (e.fuse_mask_ & 2) && (size_t)n_ops * 3 * LS <= e.fuse_max_) {
            // tensor product computed on the way into the inverse transform (ntt.cuh: kNttTensor)
            const std::vector<uint32_t> pat = ctx.pattern_ext(L);
            step([=] {
                NttFuse f{ nullptr, en->idx_.at(ao), en->idx_.at(bo), nullptr, (int)L, (int)S, 0 };
                en->ctx.ntt_fused(kNttTensor, en->arena_.buf.p, n_ops * 3 * LS, pat, nullptr, nullptr, en->arena_.buf.p + (size_t)scratch * en->ctx.N, f);
            });
        } else {
            step([=] { en->run_tensor(L, n_ops, en->idx_.at(ao), en->idx_.at(bo), en->idx_.at(dof)); });
            ntt_run(scratch, n_ops * 3 * LS, ctx.pattern_ext(L), true);
        }
        step([=] { en->run_scale_down(L, n_ops * 3, en->idx_.at(so), en->idx_.at(sd)); });
    }
    static uint32_t multiply_scratch(const DeviceContext &c, uint32_t L, uint32_t n_ops) { return n_ops * 3 * (L + (uint32_t)c.level[L].S); }

    // relinearize_inplace: size-3 ct3[o] ([3][L][N]) -> size-2 dst[o] ([2][L][N]).
    // scratch: n_ops*(L + 2)*(L + 1) polynomials
    void relinearize(uint32_t L, const std::vector<uint32_t> &ct3, const std::vector<uint32_t> &dst, uint32_t scratch, bool mirror = false)
    {
        if (ct3.empty()) return;
        const uint32_t R = L + 1, n_ops = (uint32_t)ct3.size();
        const uint32_t dig0 = scratch, acc0 = scratch + n_ops * L * R;
        std::vector<uint32_t> nsrc, ndst, dig(n_ops), acc(n_ops);
        for (uint32_t o = 0; o < n_ops; o++) {
            dig[o] = dig0 + o * L * R;
            acc[o] = acc0 + o * 2 * R;
            for (uint32_t J = 0; J < L; J++)
                for (uint32_t I = 0; I < R; I++) {
                    nsrc.push_back(ct3[o] + 2 * L + J);
                    ndst.push_back(dig[o] + J * R + I);
                }
        }
        ntt(nsrc, ndst, ctx.pattern_ks(L), false, /*reduce=*/true);
        size_t dg = idx.add(dig), ac = idx.add(acc), ct = idx.add(ct3), ds = idx.add(dst);
        Engine *en = &e;
        if (e.fuse_ && (e.fuse_mask_ & 4) && (size_t)n_ops * 2 * R <= e.fuse_max_) {
            // inner product with the keys computed on the way into the inverse transform (ntt.cuh: kNttKsMac)
            const std::vector<uint32_t> pat = ctx.pattern_ks(L);
            step([=] {
                NttFuse f{ nullptr, en->idx_.at(dg), nullptr, en->relin_keys_.p, (int)L, 0, (int)en->ctx.K };
                en->ctx.ntt_fused(kNttKsMac, en->arena_.buf.p, n_ops * 2 * R, pat, nullptr, nullptr, en->arena_.buf.p + (size_t)acc0 * en->ctx.N, f);
            });
        } else {
            step([=] { en->run_ks_mac(L, n_ops, en->idx_.at(dg), en->idx_.at(ac)); });
            ntt_run(acc0, n_ops * 2 * R, ctx.pattern_ks(L), true);
        }
        step([=] { en->run_ks_moddown(L, n_ops, en->idx_.at(ac), en->idx_.at(ct), en->idx_.at(ds), mirror); });
    }
    static uint32_t relin_scratch(uint32_t L, uint32_t n_ops) { return n_ops * (L + 2) * (L + 1); }

    // mod_switch_to_next on RNS polynomials: src[k] ([L][N]) -> dst[k] ([L-1][N])
    void mod_switch_next(uint32_t L, const std::vector<uint32_t> &src, const std::vector<uint32_t> &dst)
    {
        if (src.empty()) return;
        size_t so = idx.add(src), dn = idx.add(dst);
        uint32_t n = (uint32_t)src.size();
        Engine *en = &e;
        step([=] { en->run_mod_switch_next(L, n, en->idx_.at(so), en->idx_.at(dn)); });
    }

    // tile-major split power tables (k_pack_powers): table z gathers T*2 (term, component) sources
    void pack_powers(uint32_t L, uint32_t T, const std::vector<uint32_t> &src, const std::vector<uint32_t> &dst)
    {
        if (dst.empty() || !T) return;
        size_t so = idx.add(src), dn = idx.add(dst);
        uint32_t nz = (uint32_t)dst.size();
        Engine *en = &e;
        step([=] {
            launch_pdl(k_pack_powers, dim3(L * en->ctx.N / kKtCols, (T * 2 + kPackPer - 1) / kPackPer, nz), dim3(kKtCols), 0, en->ctx.stream, en->arena_.buf.p, (const u32 *)en->idx_.at(so),
                       (const u32 *)en->idx_.at(dn), T, (int)en->ctx.N, en->split_);
            APSU_CUDA_CHECK(cudaGetLastError());
            en->ctx.launches++;
        });
    }

    // dst[k] = sum of the RNS polynomials ([L][N]) listed in terms[k]
    void sum_polys(uint32_t L, const std::vector<std::vector<uint32_t>> &terms, const std::vector<uint32_t> &dst)
    {
        if (dst.empty()) return;
        std::vector<uint32_t> flat, first, cnt;
        for (auto &tl : terms) {
            first.push_back((uint32_t)flat.size());
            cnt.push_back((uint32_t)tl.size());
            flat.insert(flat.end(), tl.begin(), tl.end());
        }
        if (flat.empty()) flat.push_back(0);
        size_t fo = idx.add(flat), fi = idx.add(first), cn = idx.add(cnt), ds = idx.add(dst);
        uint32_t n = (uint32_t)dst.size();
        Engine *en = &e;
        step([=] {
            launch_pdl(k_sum_polys, dim3(en->ctx.N / kEwThreads, L, n), dim3(kEwThreads), 0, en->ctx.stream, en->arena_.buf.p, (const u32 *)en->idx_.at(fo),
                       (const u32 *)en->idx_.at(fi), (const u32 *)en->idx_.at(cn), (const u32 *)en->idx_.at(ds), en->ctx.level[L], (int)en->ctx.N);
            APSU_CUDA_CHECK(cudaGetLastError());
            en->ctx.launches++;
        });
    }
};

// ------------------------------------------------------------------------------------------------
// Engine
// ------------------------------------------------------------------------------------------------
Engine::Engine(const apsu_b200_params &p, int device) : ctx(p, device)
{
    std::set<uint32_t> sources(p.query_powers, p.query_powers + p.query_power_count);
    if (!dag.configure(sources, create_powers_set(p.ps_low_degree, p.max_items_per_bin)))
        throw std::invalid_argument("failed to configure PowersDag");
    if (dag.depth() > 0 && !ctx.using_keyswitching())
        throw std::invalid_argument("parameters need ciphertext multiplications but provide no key-switching prime");
    db.resize(p.bundle_idx_count);
    {
        // DB-stream operand split and lane-fold period (db_stream.cuh) from the largest data prime
        int b = 0;
        const uint32_t ndata = ctx.K > 1 ? ctx.K - 1 : 1;
        for (uint32_t j = 0; j < ndata; j++) b = std::max(b, hm::bit_length(p.coeff_modulus[j]));
        if (b > 60) throw std::invalid_argument("coeff_modulus primes must have at most 60 bits");
        split_ = std::max(1, (b + 1) / 2);
        const unsigned __int128 one = 1;
        const unsigned __int128 sum_max = ((one << split_) - 1) + ((one << std::max(0, b - split_)) - 1);
        const unsigned __int128 prod_max = sum_max * sum_max;
        const unsigned __int128 room = ((one << 64) - 1) - (one << (split_ + 1));
        const unsigned __int128 cap = prod_max ? room / prod_max : room; // terms the kk lane can take after a fold
        if (cap < (unsigned)kKtTS) throw std::logic_error("DB-stream lanes cannot hold one ring stage");
        fold_stages_ = (uint32_t)std::min<unsigned __int128>(cap / kKtTS, 0x7FFFFFFFu);
    }
    levels_dev_.upload(ctx.level, ctx.stream);
    query_bad_.alloc(3);
    APSU_CUDA_CHECK(cudaMemsetAsync(query_bad_.p, 0, 3 * sizeof(int), ctx.stream));
    {
        // function attributes are per device: set once per context, not per launch
        auto kern = k_db_mac_kt<kKtStages, kKtCtasPerSm>;
        constexpr size_t smem = kt_smem_bytes(kKtStages);
        APSU_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        APSU_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kKtThreads, smem));
        per_sm = std::max(1, std::min(per_sm, kKtCtasPerSm));
        kt_grid_cap_ = (uint32_t)(ctx.sms * per_sm);
    }
    if (const char *ev = std::getenv("APSU_B200_NO_GRAPH")) use_graphs_ = atoi(ev) == 0;
    if (const char *ev = std::getenv("APSU_B200_CHUNK")) eval_chunk_ = (uint32_t)std::max(1, atoi(ev));
    if (const char *ev = std::getenv("APSU_B200_FUSE")) fuse_ = atoi(ev) != 0; // A/B: element-wise producers fused into the transforms
    if (const char *ev = std::getenv("APSU_B200_FUSE_MASK")) fuse_mask_ = (unsigned)atoi(ev); // 1 extension, 2 tensor product, 4 key-switch MAC
    if (const char *ev = std::getenv("APSU_B200_FUSE_MAX")) fuse_max_ = (size_t)std::max(0, atoi(ev)); // ... for launches of at most this many polynomials
    for (auto &ev : ev_) APSU_CUDA_CHECK(cudaEventCreate(&ev));
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
}

Engine::~Engine()
{
    drop_graphs();
    for (auto &g : fin_groups_) {
        if (g.done) cudaEventDestroy(g.done);
        if (g.copied) cudaEventDestroy(g.copied);
    }
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    if (masks_free_) cudaEventDestroy(masks_free_);
    if (masks_ready_) cudaEventDestroy(masks_ready_);
    if (flags_host_) cudaFreeHost(flags_host_);
    for (auto &e : ev_)
        if (e) cudaEventDestroy(e);
    for (auto &pr : mac_events_) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
}

uint32_t Engine::total_bundles() const
{
    uint32_t n = 0;
    for (auto &v : db) n += (uint32_t)v.size();
    return n;
}

uint64_t Engine::stream_bytes() const
{
    uint64_t words = 0;
    for (auto &v : db)
        for (auto &s : v) words += (uint64_t)s->n_ntt * ctx.low_L * ctx.N + (uint64_t)s->n_plain * ctx.N;
    return words * 8;
}

void Engine::clear_db()
{
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    for (auto &v : db) v.clear();
    for (DBuf<u64> *b : { &build_roots_, &build_M_, &build_enc_, &stage_ }) b->release(); // build scratch
    invalidate_plan();
}

// lifts + NTTs the high-degree coefficient-form plaintexts once (multiply_plain_normal operand)
void Engine::prepare_plain_high(BinBundleStore &s)
{
    const uint32_t N = ctx.N, Lh = ctx.high_L;
    if (!ctx.params.ps_low_degree || s.n_plain <= 1) return;
    uint32_t cnt = s.n_plain - 1;
    s.plain_high_ntt.alloc((size_t)cnt * Lh * N);
    stage_.ensure((size_t)cnt * Lh * N); // reuse is ordered by the stream; a reallocation (cudaFree) synchronises by itself
    k_plain_lift<<<dim3(N / kEwThreads, Lh, cnt), kEwThreads, 0, ctx.stream>>>(s.plain_coeffs.p + N, stage_.p, ctx.level[Lh], ctx.t, (int)N);
    APSU_LAUNCH_CHECK();
    ctx.ntt(stage_.p, stage_.p, cnt * Lh, ctx.pattern_q(Lh), false);
    pack_tile(stage_.p, s.plain_high_ntt.p, cnt, Lh);
}

uint32_t Engine::add_binbundle(uint32_t bundle_idx, const uint64_t *const *coeffs, uint32_t ncoeffs)
{
    if (bundle_idx >= db.size()) throw std::invalid_argument("bundle_idx is out of range");
    if (!coeffs || !ncoeffs) throw std::invalid_argument("a BinBundle needs at least the constant coefficient");
    if (ncoeffs > ctx.params.max_items_per_bin) throw std::invalid_argument("too many coefficients: degree exceeds max_items_per_bin - 1");
    const uint32_t N = ctx.N, Ll = ctx.low_L;
    auto s = std::make_unique<BinBundleStore>();
    s->bundle_idx = bundle_idx;
    s->cache_idx = (uint32_t)db[bundle_idx].size();
    s->ncoeffs = ncoeffs;
    for (uint32_t k = 0; k < ncoeffs; k++) (is_ntt_degree(k) ? s->n_ntt : s->n_plain)++;
    s->ntt_coeffs.alloc((size_t)s->n_ntt * Ll * N);
    s->plain_coeffs.alloc((size_t)s->n_plain * N);
    uint32_t in = 0, ip = 0;
    for (uint32_t k = 0; k < ncoeffs; k++)
        if (!coeffs[k]) throw std::invalid_argument("null plaintext pointer");
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    stage_.ensure((size_t)s->n_ntt * Ll * N);
    for (uint32_t k = 0; k < ncoeffs; k++) {
        if (is_ntt_degree(k))
            APSU_CUDA_CHECK(cudaMemcpyAsync(stage_.p + (size_t)(in++) * Ll * N, coeffs[k], (size_t)Ll * N * 8, cudaMemcpyHostToDevice, ctx.stream));
        else
            APSU_CUDA_CHECK(cudaMemcpyAsync(s->plain_coeffs.p + (size_t)(ip++) * N, coeffs[k], (size_t)N * 8, cudaMemcpyHostToDevice, ctx.stream));
    }
    pack_tile(stage_.p, s->ntt_coeffs.p, s->n_ntt, Ll);
    prepare_plain_high(*s);
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    uint32_t ci = s->cache_idx;
    db[bundle_idx].push_back(std::move(s));
    invalidate_plan();
    return ci;
}

// NTT-form DB words live tile-major and split, the form the DB-stream kernel consumes (db_stream.cuh):
// standard [rows][L][N] -> [L*N/128][rows][128]
void Engine::pack_tile(const u64 *src, u64 *dst, uint32_t rows, uint32_t L)
{
    if (!rows) return;
    k_pack_tile<<<dim3(L * ctx.N / kKtCols, (rows + 7) / 8), kKtCols, 0, ctx.stream>>>(src, dst, rows, L * ctx.N, split_);
    APSU_LAUNCH_CHECK();
}

uint32_t Engine::add_binbundle_synthetic(uint32_t bundle_idx, uint32_t ncoeffs, uint64_t seed)
{
    if (bundle_idx >= db.size()) throw std::invalid_argument("bundle_idx is out of range");
    if (!ncoeffs || ncoeffs > ctx.params.max_items_per_bin) throw std::invalid_argument("ncoeffs is out of range");
    const uint32_t N = ctx.N, Ll = ctx.low_L, ps = ctx.params.ps_low_degree;
    auto s = std::make_unique<BinBundleStore>();
    s->bundle_idx = bundle_idx;
    s->cache_idx = (uint32_t)db[bundle_idx].size();
    s->ncoeffs = ncoeffs;
    for (uint32_t k = 0; k < ncoeffs; k++) (is_ntt_degree(k) ? s->n_ntt : s->n_plain)++;
    s->ntt_coeffs.alloc((size_t)s->n_ntt * Ll * N);
    s->plain_coeffs.alloc((size_t)s->n_plain * N);
    FillMods fm;
    std::memset(&fm, 0, sizeof(fm));
    for (uint32_t j = 0; j < Ll; j++) fm.q[j] = ctx.params.coeff_modulus[j];
    fm.t = ctx.t;
    fm.L = (int)Ll;
    size_t cn = s->ntt_coeffs.n, cp = s->plain_coeffs.n;
    if (cn) {
        k_fill_db<<<(unsigned)((cn + 255) / 256), 256, 0, ctx.stream>>>(s->ntt_coeffs.p, cn, seed, 0, ps ? ps + 1 : 0, fm, (int)N, s->n_ntt, split_);
        APSU_LAUNCH_CHECK();
    }
    k_fill_db<<<(unsigned)((cp + 255) / 256), 256, 0, ctx.stream>>>(s->plain_coeffs.p, cp, seed, 1, ps ? ps + 1 : 0, fm, (int)N, 0, split_);
    APSU_LAUNCH_CHECK();
    prepare_plain_high(*s);
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    uint32_t ci = s->cache_idx;
    db[bundle_idx].push_back(std::move(s));
    invalidate_plan();
    return ci;
}

// BinBundle::regen_cache on the device (bin_bundle.cpp:934-1041): regen_polyns = polyn_with_roots per bin,
// regen_plaintexts = BatchedPlaintextPolyn ctor (bin_bundle.cpp:366-430): plaintext i = BatchEncoder::encode of
// the degree-i coefficients of all bins, NTT form at the plaintext level unless i == 0 / i % (ps_low+1) == 0.
uint32_t Engine::add_binbundle_from_bins(uint32_t bundle_idx, const uint32_t *bin_sizes, const uint64_t *roots)
{
    if (bundle_idx >= db.size()) throw std::invalid_argument("bundle_idx is out of range");
    const apsu_b200_params &p = ctx.params;
    const uint32_t nbins = p.bins_per_bundle;
    std::vector<uint32_t> first(nbins);
    uint32_t max_deg = 0;
    size_t total = 0;
    for (uint32_t b = 0; b < nbins; b++) {
        first[b] = (uint32_t)total;
        total += bin_sizes[b];
        max_deg = std::max(max_deg, bin_sizes[b]);
        // a bin holds fewer than max_items_per_bin items (receiver_db.cpp:388-389)
        if (bin_sizes[b] >= p.max_items_per_bin) throw std::invalid_argument("a bin holds max_items_per_bin or more items");
    }
    if (total >= (1ull << 32)) throw std::invalid_argument("too many items in one BinBundle");
    // scratch of the build is kept between calls (a DB build is dozens of BinBundles of the same shape)
    build_first_.upload(first, ctx.stream);
    build_size_.ensure(nbins);
    APSU_CUDA_CHECK(cudaMemcpyAsync(build_size_.p, bin_sizes, nbins * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx.stream));
    build_roots_.ensure(std::max<size_t>(total, 1));
    if (total) APSU_CUDA_CHECK(cudaMemcpyAsync(build_roots_.p, roots, total * 8, cudaMemcpyHostToDevice, ctx.stream));
    const uint32_t ci = add_binbundle_from_bins_device(bundle_idx, build_first_.p, build_size_.p, build_roots_.p, max_deg);
    throw_if_build_invalid(); // synchronises
    return ci;
}

void Engine::throw_if_build_invalid()
{
    int bad = 0;
    build_bad_.ensure(1);
    APSU_CUDA_CHECK(cudaMemcpyAsync(&bad, build_bad_.p, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    if (bad) {
        APSU_CUDA_CHECK(cudaMemsetAsync(build_bad_.p, 0, sizeof(int), ctx.stream));
        throw std::invalid_argument("bin item is not a field element (>= plain_modulus)");
    }
}

// The device part of the build: bins already on the device (first / size per bin, roots concatenated), max_deg =
// the largest bin.  Asynchronous on the context stream: no host synchronisation (a non-field-element root is
// recorded in build_bad_ and reported by throw_if_build_invalid).
uint32_t Engine::add_binbundle_from_bins_device(uint32_t bundle_idx, const uint32_t *d_first, const uint32_t *d_size, const u64 *d_roots, uint32_t max_deg)
{
    if (bundle_idx >= db.size()) throw std::invalid_argument("bundle_idx is out of range");
    const apsu_b200_params &p = ctx.params;
    const uint32_t N = ctx.N, Ll = ctx.low_L, nbins = p.bins_per_bundle;
    if (max_deg >= p.max_items_per_bin) throw std::invalid_argument("a bin holds max_items_per_bin or more items");
    const uint32_t ncoeffs = max_deg + 1;

    auto s = std::make_unique<BinBundleStore>();
    s->bundle_idx = bundle_idx;
    s->cache_idx = (uint32_t)db[bundle_idx].size();
    s->ncoeffs = ncoeffs;
    std::vector<uint32_t> ntt_rows, plain_rows;
    for (uint32_t k = 0; k < ncoeffs; k++) (is_ntt_degree(k) ? ntt_rows : plain_rows).push_back(k);
    s->n_ntt = (uint32_t)ntt_rows.size();
    s->n_plain = (uint32_t)plain_rows.size();
    s->ntt_coeffs.alloc((size_t)s->n_ntt * Ll * N);
    s->plain_coeffs.alloc((size_t)s->n_plain * N);

    DBuf<u64> &M = build_M_, &enc = build_enc_;
    if (!build_bad_.p) {
        build_bad_.alloc(1);
        APSU_CUDA_CHECK(cudaMemsetAsync(build_bad_.p, 0, sizeof(int), ctx.stream));
    }
    M.ensure((size_t)ncoeffs * N);
    enc.ensure((size_t)ncoeffs * N);
    APSU_CUDA_CHECK(cudaMemsetAsync(M.p, 0, (size_t)ncoeffs * N * 8, ctx.stream)); // (enc is written whole by the slot scatter)
    {
        // one warp per bin: registers for plain moduli below 2^30 and up to 2048 coefficients, shared memory otherwise
        const DMod mt = ctx.mod_host[ctx.idx_t];
        const uint32_t need_regs = (max_deg + 1 + 31) / 32;
        if (ctx.t < (1ull << 30) && need_regs <= 64) {
            const int warps = 8;
            const unsigned grid = (nbins + warps - 1) / warps;
            auto launch = [&](auto kern) { kern<<<grid, warps * 32, 0, ctx.stream>>>(d_first, d_size, d_roots, M.p, nbins, (u32)ctx.t, (int)N, build_bad_.p); };
            if (need_regs <= 4) launch(k_polyn_with_roots_reg<4>);
            else if (need_regs <= 8) launch(k_polyn_with_roots_reg<8>);
            else if (need_regs <= 16) launch(k_polyn_with_roots_reg<16>);
            else if (need_regs <= 32) launch(k_polyn_with_roots_reg<32>);
            else if (need_regs <= 44) launch(k_polyn_with_roots_reg<44>);
            else launch(k_polyn_with_roots_reg<64>);
        } else {
            // as many warps per block as the polynomials leave shared memory for
            const size_t per_warp = (size_t)(max_deg + 2) * 8;
            int warps = (int)std::max<size_t>(1, std::min<size_t>(8, (160 * 1024) / per_warp));
            const size_t smem = per_warp * warps;
            const bool small = ctx.t < (1ull << 32);
            auto kern = small ? k_polyn_with_roots<true> : k_polyn_with_roots<false>;
            APSU_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(nbins + warps - 1) / warps, warps * 32, smem, ctx.stream>>>(d_first, d_size, d_roots, M.p, nbins, max_deg, mt, (int)N, build_bad_.p);
        }
        APSU_LAUNCH_CHECK();
    }
    // BatchEncoder::encode: slot permutation + inverse NTT modulo t
    k_slot_scatter<<<dim3(N / 256, ncoeffs), 256, 0, ctx.stream>>>(M.p, enc.p, ctx.slot_map.p, (int)N);
    APSU_LAUNCH_CHECK();
    ctx.ntt(enc.p, enc.p, ncoeffs, { ctx.idx_t }, true);
    // coefficient-form plaintexts: rows plain_rows of enc, gathered in one launch
    build_rows_.upload(plain_rows, ctx.stream);
    k_gather_rows<<<dim3(N / 256, s->n_plain), 256, 0, ctx.stream>>>(enc.p, s->plain_coeffs.p, build_rows_.p, (int)N);
    APSU_LAUNCH_CHECK();
    if (s->n_ntt) {
        // transform_to_ntt_inplace(pt, plaintext level): lift + NTT, then the tile-major split store
        build_rows2_.upload(ntt_rows, ctx.stream);
        stage_.ensure((size_t)s->n_ntt * Ll * N);
        k_plain_lift<<<dim3(N / kEwThreads, Ll, s->n_ntt), kEwThreads, 0, ctx.stream>>>(enc.p, stage_.p, ctx.level[Ll], ctx.t, (int)N, build_rows2_.p);
        APSU_LAUNCH_CHECK();
        ctx.ntt(stage_.p, stage_.p, s->n_ntt * Ll, ctx.pattern_q(Ll), false);
        pack_tile(stage_.p, s->ntt_coeffs.p, s->n_ntt, Ll);
    }
    prepare_plain_high(*s);
    uint32_t ci = s->cache_idx;
    db[bundle_idx].push_back(std::move(s));
    invalidate_plan();
    return ci;
}

// ------------------------------------------------------------------------------------------------
// query inputs
// ------------------------------------------------------------------------------------------------
void Engine::set_relin_keys(const void *keys, bool on_device)
{
    if (!ctx.using_keyswitching()) {
        have_keys_ = false;
        return;
    }
    if (!keys) throw std::invalid_argument("relinearization keys are required for this parameter set");
    size_t words = (size_t)(ctx.K - 1) * 2 * ctx.K * ctx.N;
    relin_keys_.ensure(words);
    APSU_CUDA_CHECK(cudaMemcpyAsync(relin_keys_.p, keys, words * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx.stream));
    check_range(relin_keys_.p, (ctx.K - 1) * 2 * ctx.K, ctx.params.coeff_modulus, ctx.K); // is_valid_for(relin_keys), query.cpp:45-51
    have_keys_ = true;
}

void Engine::set_masks(const void *masks, uint32_t npack, bool on_device)
{
    if (!masks || !npack) throw std::invalid_argument("masks are required");
    join_masks_upload();
    masks_.ensure((size_t)npack * ctx.N);
    APSU_CUDA_CHECK(cudaMemcpyAsync(masks_.p, masks, (size_t)npack * ctx.N * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx.stream));
    npack_ = npack;
}

// The masks are only read by the very last kernel of the evaluation, so their upload does not have to precede
// ComputePowers: it runs on the copy stream (after everything already queued on the context stream, which may still
// read the old masks) and eval_all makes the context stream wait for it.
void Engine::set_masks_overlapped(const void *masks, uint32_t npack)
{
    if (!masks || !npack) throw std::invalid_argument("masks are required");
    masks_.ensure((size_t)npack * ctx.N);
    if (!copy_stream_) APSU_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    if (!masks_free_) {
        APSU_CUDA_CHECK(cudaEventCreateWithFlags(&masks_free_, cudaEventDisableTiming));
        APSU_CUDA_CHECK(cudaEventCreateWithFlags(&masks_ready_, cudaEventDisableTiming));
    }
    APSU_CUDA_CHECK(cudaEventRecord(masks_free_, ctx.stream));
    APSU_CUDA_CHECK(cudaStreamWaitEvent(copy_stream_, masks_free_, 0));
    APSU_CUDA_CHECK(cudaMemcpyAsync(masks_.p, masks, (size_t)npack * ctx.N * 8, cudaMemcpyHostToDevice, copy_stream_));
    APSU_CUDA_CHECK(cudaEventRecord(masks_ready_, copy_stream_));
    masks_pending_ = true;
    npack_ = npack;
}

// orders the context stream behind an upload started by set_masks_overlapped (no-op otherwise)
void Engine::join_masks_upload()
{
    if (!masks_pending_) return;
    APSU_CUDA_CHECK(cudaStreamWaitEvent(ctx.stream, masks_ready_, 0));
    masks_pending_ = false;
}

void Engine::encode_masks(const uint64_t *slot_values, uint32_t npack, uint64_t *out)
{
    if (!slot_values || !out || !npack) throw std::invalid_argument("encode_masks: bad arguments");
    const uint32_t N = ctx.N;
    DBuf<u64> &in = aux_[0], &enc = aux_[1]; // scratch kept between calls
    in.ensure((size_t)npack * N);
    enc.ensure((size_t)npack * N);
    APSU_CUDA_CHECK(cudaMemcpyAsync(in.p, slot_values, (size_t)npack * N * 8, cudaMemcpyHostToDevice, ctx.stream));
    k_slot_scatter<<<dim3(N / 256, npack), 256, 0, ctx.stream>>>(in.p, enc.p, ctx.slot_map.p, (int)N);
    APSU_LAUNCH_CHECK();
    ctx.ntt(enc.p, enc.p, npack, { ctx.idx_t }, true);
    APSU_CUDA_CHECK(cudaMemcpyAsync(out, enc.p, (size_t)npack * N * 8, cudaMemcpyDeviceToHost, ctx.stream));
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
}

// RunQuery's mask generation on the device (receiver_ddh.cpp:218-283): SEAL's blake2xb generator keyed with 64 seed
// bytes (the reference takes them from random_bytes, :221-225; a NULL seed does the same here through getrandom),
// one 32-bit draw per slot in (cache_idx, bundle_idx) order skipping padded pairs, the values' encodings (kept
// resident as the masks of the next evaluation) and the PEQT blocks.
void Engine::generate_masks(const uint8_t *seed64, const uint8_t *padded, uint32_t npack, uint64_t *blocks_out, uint64_t *values_out, bool synchronise)
{
    if (!padded || !npack) throw std::invalid_argument("generate_masks: bad arguments");
    const uint32_t N = ctx.N, ipb = ctx.params.items_per_bundle;
    PrngSeed seed;
    if (seed64) {
        std::memcpy(seed.w, seed64, sizeof(seed.w));
    } else {
        size_t got = 0; // random_bytes (receiver_ddh.cpp:223)
        while (got < sizeof(seed.w)) {
            ssize_t r = getrandom(reinterpret_cast<unsigned char *>(seed.w) + got, sizeof(seed.w) - got, 0);
            if (r < 0) throw std::runtime_error("getrandom failed");
            got += (size_t)r;
        }
    }
    std::vector<uint32_t> seq(npack);
    uint32_t s = 0;
    for (uint32_t p = 0; p < npack; p++) seq[p] = padded[p] ? 0xFFFFFFFFu : s++;
    DBuf<u64> &values = aux_[0], &blocks = aux_[1];
    values.ensure((size_t)npack * N);
    blocks.ensure((size_t)npack * ipb * 2);
    aux_bytes_.ensure(npack);
    join_masks_upload();
    masks_.ensure((size_t)npack * N);
    aux_idx_.upload(seq, ctx.stream);
    APSU_CUDA_CHECK(cudaMemcpyAsync(aux_bytes_.p, padded, npack, cudaMemcpyHostToDevice, ctx.stream));
    k_prng_mask_values<<<dim3(N / 1024, npack), 64, 0, ctx.stream>>>(values.p, aux_idx_.p, seed, ctx.t, (int)N);
    APSU_LAUNCH_CHECK();
    k_masks_scatter_blocks<<<dim3(N / 256, npack), 256, 0, ctx.stream>>>(values.p, masks_.p, blocks.p, aux_bytes_.p, ctx.slot_map.p, ctx.t,
                                                                          ctx.params.felts_per_item, ipb, (int)N);
    APSU_LAUNCH_CHECK();
    ctx.ntt(masks_.p, masks_.p, npack, { ctx.idx_t }, true);
    npack_ = npack;
    if (blocks_out) APSU_CUDA_CHECK(cudaMemcpyAsync(blocks_out, blocks.p, (size_t)npack * ipb * 2 * 8, cudaMemcpyDeviceToHost, ctx.stream));
    if (values_out) APSU_CUDA_CHECK(cudaMemcpyAsync(values_out, values.p, (size_t)npack * N * 8, cudaMemcpyDeviceToHost, ctx.stream));
    if (synchronise) APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream)); // else: the outputs are valid after the next synchronising call
}

// sample_poly_uniform for `n` seeded polynomials of L primes each into arena blocks dst[k] ([L][N]); seeds on the host
void Engine::expand_seeds(uint32_t L, const std::vector<uint32_t> &dst, const uint8_t *seeds64, u64 *base, const uint64_t *moduli)
{
    const uint32_t n = (uint32_t)dst.size(), N = ctx.N;
    if (!n) return;
    UniformMods um;
    std::memset(&um, 0, sizeof(um));
    um.L = (int)L;
    for (uint32_t j = 0; j < L; j++) {
        const uint64_t q = moduli[j];
        for (size_t m = 0; m < ctx.mod_values.size(); m++)
            if (ctx.mod_values[m] == q) um.q[j] = ctx.mod_host[m];
        um.max_multiple[j] = ~0ull - (~0ull % q) - 1;
    }
    seed_buf_.ensure((size_t)n * sizeof(PrngSeed));
    rej_.ensure((size_t)n * (kMaxRej + 1));
    APSU_CUDA_CHECK(cudaMemcpyAsync(seed_buf_.p, seeds64, (size_t)n * sizeof(PrngSeed), cudaMemcpyHostToDevice, ctx.stream));
    APSU_CUDA_CHECK(cudaMemsetAsync(rej_.p, 0, (size_t)n * sizeof(uint32_t), ctx.stream));
    seed_dst_.upload(dst, ctx.stream);
    const PrngSeed *sd = reinterpret_cast<const PrngSeed *>(seed_buf_.p);
    uint32_t *rej_count = rej_.p, *rej_pos = rej_.p + n;
    k_prng_sample_uniform<<<dim3(L * N / 512, n), 64, 0, ctx.stream>>>(base, seed_dst_.p, sd, um, (int)N, rej_count, rej_pos);
    APSU_LAUNCH_CHECK();
    k_prng_fix_rejects<<<(n + 63) / 64, 64, 0, ctx.stream>>>(base, seed_dst_.p, sd, um, (int)N, rej_count, rej_pos, n, query_bad_.p + 1);
    query_checked_ = true;
    APSU_LAUNCH_CHECK();
}

// ResultPackage::extract for a batch of result ciphertexts (result_package.cpp:175-213, sender_ddh.cpp:580-605):
// decrypt at the last level with the secret key (NTT form modulo q0, i.e. the first N words of seal::SecretKey),
// BatchEncoder::decode, and the items' 128-bit blocks.
void Engine::decrypt_results(const uint64_t *secret_ntt_q0, const uint64_t *cts, uint32_t n, uint64_t *values_out, uint64_t *blocks_out, int32_t *budget_out)
{
    if (!secret_ntt_q0 || !cts || !n || !values_out) throw std::invalid_argument("decrypt_results: bad arguments");
    const uint32_t N = ctx.N, ipb = ctx.params.items_per_bundle;
    const DMod q0 = ctx.mod_host[0];
    DBuf<u64> &d_ct = aux_[0], &d_s = aux_[1], &tmp = aux_[2], &plain = aux_[3], &values = aux_[4], &blocks = aux_[5]; // kept between calls
    DBuf<int> &budget = aux_int_;
    DBuf<uint32_t> &src_idx = aux_idx_;
    d_ct.ensure((size_t)n * 2 * N);
    d_s.ensure(N);
    tmp.ensure((size_t)n * N);
    plain.ensure((size_t)n * N);
    values.ensure((size_t)n * N);
    blocks.ensure((size_t)n * ipb * 2);
    budget.ensure(n);
    std::vector<uint32_t> idx(n);
    for (uint32_t k = 0; k < n; k++) idx[k] = 2 * k + 1; // the c1 polynomials
    src_idx.upload(idx, ctx.stream);
    std::vector<int> big(n, 1 << 20);
    APSU_CUDA_CHECK(cudaMemcpyAsync(d_ct.p, cts, (size_t)n * 2 * N * 8, cudaMemcpyHostToDevice, ctx.stream));
    APSU_CUDA_CHECK(cudaMemcpyAsync(d_s.p, secret_ntt_q0, (size_t)N * 8, cudaMemcpyHostToDevice, ctx.stream));
    APSU_CUDA_CHECK(cudaMemcpyAsync(budget.p, big.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
    ctx.ntt(d_ct.p, tmp.p, n, { 0 }, false, src_idx.p, nullptr);
    k_mul_secret<<<dim3(N / 256, n), 256, 0, ctx.stream>>>(tmp.p, d_s.p, q0, (int)N);
    APSU_LAUNCH_CHECK();
    ctx.ntt(tmp.p, tmp.p, n, { 0 }, true);
    k_decrypt_round<<<dim3(N / 256, n), 256, 0, ctx.stream>>>(d_ct.p, tmp.p, plain.p, budget.p, q0, ctx.t, (int)N);
    APSU_LAUNCH_CHECK();
    ctx.ntt(plain.p, plain.p, n, { ctx.idx_t }, false);
    k_decode_gather<<<dim3(N / 256, n), 256, 0, ctx.stream>>>(plain.p, values.p, blocks_out ? blocks.p : nullptr, ctx.slot_map.p, ctx.t, ctx.params.felts_per_item, ipb, (int)N);
    APSU_LAUNCH_CHECK();
    APSU_CUDA_CHECK(cudaMemcpyAsync(values_out, values.p, (size_t)n * N * 8, cudaMemcpyDeviceToHost, ctx.stream));
    if (blocks_out) APSU_CUDA_CHECK(cudaMemcpyAsync(blocks_out, blocks.p, (size_t)n * ipb * 2 * 8, cudaMemcpyDeviceToHost, ctx.stream));
    if (budget_out) APSU_CUDA_CHECK(cudaMemcpyAsync(budget_out, budget.p, n * sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
}

// Query validation (receiver/apsu/query.cpp:68-111): the source powers must be exactly the parameter set's
std::vector<uint32_t> Engine::check_query_powers(const uint32_t *src_powers, uint32_t nsrc)
{
    const apsu_b200_params &p = ctx.params;
    if (!src_powers) throw std::invalid_argument("query is invalid");
    std::set<uint32_t> given(src_powers, src_powers + nsrc), want(p.query_powers, p.query_powers + p.query_power_count);
    if (given.size() != nsrc || given != want) throw std::invalid_argument("query powers do not match the query_powers of the parameters");
    if (!plan_valid_) build_plan();
    std::vector<uint32_t> sorted(want.begin(), want.end()), rank(nsrc);
    for (uint32_t k = 0; k < nsrc; k++) rank[k] = (uint32_t)(std::lower_bound(sorted.begin(), sorted.end(), src_powers[k]) - sorted.begin());
    return rank; // the plan addresses sources by sorted rank; the caller's block keeps its own order
}

// is_valid_for of the loaded query ciphertexts / keys (query.cpp:45-66): residues below their moduli, checked on
// the device; the flag is read back by the next synchronising call (throw_if_query_invalid)
void Engine::check_range(const u64 *base, uint32_t n_polys, const uint64_t *moduli, uint32_t nmods)
{
    if (!n_polys) return;
    RangeMods rm;
    std::memset(&rm, 0, sizeof(rm));
    rm.n = (int)nmods;
    for (uint32_t j = 0; j < nmods; j++) rm.q[j] = moduli[j];
    k_check_range<<<dim3(ctx.N / kEwThreads, n_polys), kEwThreads, 0, ctx.stream>>>(base, rm, (int)ctx.N, query_bad_.p);
    APSU_LAUNCH_CHECK();
    query_checked_ = true;
}
void Engine::throw_if_query_invalid()
{
    if (!query_checked_) return;
    enqueue_flag_read();
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    check_flags_after_sync();
}
// the flag read rides on whatever synchronisation the caller does next: enqueue (pinned destination, truly
// asynchronous), let the caller enqueue its own copies and synchronise once, then check
void Engine::enqueue_flag_read()
{
    if (!query_checked_) return;
    if (!flags_host_) APSU_CUDA_CHECK(cudaHostAlloc(&flags_host_, 3 * sizeof(int), cudaHostAllocDefault));
    APSU_CUDA_CHECK(cudaMemcpyAsync(flags_host_, query_bad_.p, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
}
void Engine::check_flags_after_sync()
{
    if (!query_checked_) return;
    query_checked_ = false;
    const int bad[3] = { flags_host_[0], flags_host_[1], flags_host_[2] }; // [0] residue out of range, [1] rejection list overflow, [2] peer barrier timed out
    if (bad[0] | bad[1] | bad[2]) APSU_CUDA_CHECK(cudaMemsetAsync(query_bad_.p, 0, sizeof(bad), ctx.stream));
    if (bad[0]) {
        query_loaded_ = false;
        throw std::invalid_argument("query ciphertexts or keys are not valid for the encryption parameters (residue >= modulus)");
    }
    if (bad[1]) {
        query_loaded_ = false;
        throw std::runtime_error("seed expansion: more rejected samples than the device list holds");
    }
    if (bad[2]) {
        query_loaded_ = false;
        throw std::runtime_error("multi-GPU ComputePowers: a peer of the PowersDag partition did not reach the barrier");
    }
}

void Engine::query_begin(const uint32_t *src_powers, uint32_t nsrc, const void *cts, bool on_device)
{
    const apsu_b200_params &p = ctx.params;
    if (!cts) throw std::invalid_argument("query is invalid");
    const std::vector<uint32_t> rank = check_query_powers(src_powers, nsrc);
    const size_t ct_words = (size_t)2 * ctx.first_L * ctx.N, row = (size_t)p.bundle_idx_count * ct_words;
    u64 *region = arena_.buf.p + (size_t)query_region_ * ctx.N;
    for (uint32_t k = 0; k < nsrc; k++) {
        const u64 *src = (const u64 *)cts + (size_t)k * row;
        APSU_CUDA_CHECK(cudaMemcpyAsync(region + (size_t)rank[k] * row, src, row * 8, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx.stream));
    }
    check_range(region, nsrc * p.bundle_idx_count * 2 * ctx.first_L, p.coeff_modulus, ctx.first_L);
    query_loaded_ = true;
    powers_done_ = eval_done_ = false;
}

// Multi-GPU loading (mgpu.cu): the ranks of a sharded DB receive only the ciphertexts of the bundle indices they own.
// query_begin_partial validates the source powers and opens the query; query_load_index places the ciphertexts of
// ONE bundle index, device buffer [nsrc][2][L_first][N] in the caller's source order (asynchronous, context stream).
void Engine::query_begin_partial(const uint32_t *src_powers, uint32_t nsrc)
{
    partial_rank_ = check_query_powers(src_powers, nsrc);
    query_loaded_ = true;
    powers_done_ = eval_done_ = false;
}
void Engine::query_load_index(uint32_t bundle_idx, const void *cts_device)
{
    const apsu_b200_params &p = ctx.params;
    if (!query_loaded_ || partial_rank_.empty()) throw std::logic_error("query_load_index called before query_begin_partial");
    if (bundle_idx >= p.bundle_idx_count || !cts_device) throw std::invalid_argument("query_load_index: bad arguments");
    const size_t ct_words = (size_t)2 * ctx.first_L * ctx.N;
    u64 *region = arena_.buf.p + (size_t)query_region_ * ctx.N;
    const uint32_t nsrc = (uint32_t)partial_rank_.size();
    for (uint32_t k = 0; k < nsrc; k++) {
        u64 *dst = region + ((size_t)partial_rank_[k] * p.bundle_idx_count + bundle_idx) * ct_words;
        APSU_CUDA_CHECK(cudaMemcpyAsync(dst, (const u64 *)cts_device + (size_t)k * ct_words, ct_words * 8, cudaMemcpyDeviceToDevice, ctx.stream));
    }
    check_range((const u64 *)cts_device, nsrc * 2 * ctx.first_L, p.coeff_modulus, ctx.first_L);
}

// Row f2: the query as it arrives on the wire (seal::Serializable<Ciphertext>, common/apsu/seal_object.h:161-219,
// sender/apsu/plaintext_powers.cpp:45): c1 of every ciphertext is a 64-byte PRNG seed, expanded here on the device
// (sample_poly_uniform) instead of on the host: half the upload.  c0: [nsrc][bundle_idx_count][L][N].
void Engine::query_begin_seeded(const uint32_t *src_powers, uint32_t nsrc, const uint64_t *c0, const uint8_t *seeds64)
{
    const apsu_b200_params &p = ctx.params;
    if (!c0 || !seeds64) throw std::invalid_argument("query is invalid");
    const std::vector<uint32_t> rank = check_query_powers(src_powers, nsrc);
    const uint32_t bic = p.bundle_idx_count, Lf = ctx.first_L;
    const size_t poly_bytes = (size_t)Lf * ctx.N * 8;
    u64 *region = arena_.buf.p + (size_t)query_region_ * ctx.N;
    std::vector<uint32_t> dst;
    for (uint32_t k = 0; k < nsrc; k++) {
        u64 *row = region + (size_t)rank[k] * bic * 2 * Lf * ctx.N;
        APSU_CUDA_CHECK(cudaMemcpy2DAsync(row, 2 * poly_bytes, c0 + (size_t)k * bic * Lf * ctx.N, poly_bytes, poly_bytes, bic, cudaMemcpyHostToDevice, ctx.stream));
        for (uint32_t b = 0; b < bic; b++) dst.push_back(query_region_ + ((rank[k] * bic + b) * 2 + 1) * Lf);
    }
    expand_seeds(Lf, dst, seeds64, arena_.buf.p, p.coeff_modulus);
    check_range(region, nsrc * bic * 2 * Lf, p.coeff_modulus, Lf);
    query_loaded_ = true;
    powers_done_ = eval_done_ = false;
}

// Serializable<RelinKeys> (KeyGenerator::create_relin_keys): the second polynomial of every key is a seed, sampled in
// NTT form at the key level.  c0: [K-1][K][N], seeds: [K-1][64].
void Engine::set_relin_keys_seeded(const uint64_t *c0, const uint8_t *seeds64)
{
    if (!ctx.using_keyswitching()) {
        have_keys_ = false;
        return;
    }
    if (!c0 || !seeds64) throw std::invalid_argument("relinearization keys are required for this parameter set");
    const uint32_t K = ctx.K, N = ctx.N;
    relin_keys_.ensure((size_t)(K - 1) * 2 * K * N);
    const size_t poly_bytes = (size_t)K * N * 8;
    APSU_CUDA_CHECK(cudaMemcpy2DAsync(relin_keys_.p, 2 * poly_bytes, c0, poly_bytes, poly_bytes, K - 1, cudaMemcpyHostToDevice, ctx.stream));
    std::vector<uint32_t> dst;
    for (uint32_t J = 0; J + 1 < K; J++) dst.push_back((J * 2 + 1) * K);
    expand_seeds(K, dst, seeds64, relin_keys_.p, ctx.params.coeff_modulus);
    check_range(relin_keys_.p, (K - 1) * 2 * K, ctx.params.coeff_modulus, K);
    have_keys_ = true;
}

// ------------------------------------------------------------------------------------------------
// plan: arena layout + the two kernel programs.  Depends on the parameter set and on which
// BinBundles exist, so it is rebuilt lazily after DB changes.
// ------------------------------------------------------------------------------------------------
void Engine::build_plan()
{
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    const apsu_b200_params &p = ctx.params;
    const uint32_t N = ctx.N, bic = p.bundle_idx_count, Lf = ctx.first_L, Ll = ctx.low_L, Lh = ctx.high_L;
    const uint32_t ps = p.ps_low_degree, h = ps + 1;
    const uint32_t Sf = (uint32_t)ctx.level[Lf].S, LSf = Lf + Sf;
    const uint32_t Sh = (uint32_t)ctx.level[Lh].S, LSh = Lh + Sh;
    arena_.top = 0;
    arena_.high_water = 0;
    idx_.clear();
    desc_host_.clear();
    powers_prog_.clear();
    eval_prog_.clear();
    mac_step_bytes_.clear();
    for (auto &g : fin_groups_) {
        if (g.done) cudaEventDestroy(g.done);
        if (g.copied) cudaEventDestroy(g.copied);
    }
    fin_groups_.clear();

    const std::set<uint32_t> &targets = dag.target_powers();
    std::vector<uint32_t> sources(p.query_powers, p.query_powers + p.query_power_count);
    const uint32_t nsrc = (uint32_t)sources.size();
    std::vector<bool> active(bic);
    for (uint32_t b = 0; b < bic; b++) active[b] = !db[b].empty(); // receiver_ddh.cpp:399-402

    // ---- persistent regions ----
    query_region_ = arena_.take((size_t)nsrc * bic * 2 * Lf);
    const uint32_t maxp = p.max_items_per_bin;
    std::vector<std::vector<uint32_t>> loc(bic, std::vector<uint32_t>(maxp + 1, kNoSrc)); // [2][Lf][N] at first level
    for (uint32_t k = 0; k < nsrc; k++)
        for (uint32_t b = 0; b < bic; b++) loc[b][sources[k]] = query_region_ + (k * bic + b) * 2 * Lf;
    // Products live level by level, node by node, in one contiguous region per (level, bundle index): with the
    // PowersDag split over `powers_part_size_` ranks rank r computes the nodes of its chunk
    // [r*chunk, (r+1)*chunk) of every level and the region is exactly the all-gather buffer of that level.
    const uint32_t part = std::max<uint32_t>(1, powers_part_size_), prank = powers_part_rank_;
    auto dag_levels = dag.levels();
    powers_exchange_.clear();
    std::set<uint32_t> parents;
    for (uint32_t d = 1; d < dag_levels.size(); d++) {
        const uint32_t n = (uint32_t)dag_levels[d].size(), chunk = (n + part - 1) / part;
        for (uint32_t k = 0; k < n; k++) {
            const PowersNode &nd = dag_levels[d][k];
            if (k / chunk == prank) { // extended forms are only needed for the products this rank computes
                parents.insert(nd.parent1);
                parents.insert(nd.parent2);
            }
        }
        for (uint32_t b = 0; b < bic; b++) {
            if (!active[b]) continue;
            const uint32_t region = arena_.take((size_t)part * chunk * 2 * Lf);
            for (uint32_t k = 0; k < n; k++) loc[b][dag_levels[d][k].power] = region + k * 2 * Lf;
            powers_exchange_.push_back(ExchangeRegion{ d, region, chunk * 2 * Lf });
        }
    }
    std::vector<std::map<uint32_t, uint32_t>> ext_loc(bic);
    for (uint32_t b = 0; b < bic; b++)
        if (active[b])
            for (uint32_t e : parents) ext_loc[b][e] = arena_.take(2 * LSf);

    // final powers as the evaluation reads them
    const uint32_t nlow = ps ? ps : maxp;          // powers 1..nlow in NTT form at the low level
    const uint32_t nhigh = ps ? maxp / h : 0;      // powers h, 2h, .. in coefficient form at the high level
    // LOWT / HIGHT: the same powers tile-major and split, the tables the DB-stream kernel reads (db_stream.cuh)
    std::vector<uint32_t> low_base(bic, kNoSrc), highext_base(bic, kNoSrc), lowt_base(bic, kNoSrc), hight_base(bic, kNoSrc);
    std::vector<std::vector<uint32_t>> highc_loc(bic, std::vector<uint32_t>(nhigh + 1, kNoSrc));
    for (uint32_t b = 0; b < bic; b++) {
        if (!active[b]) continue;
        low_base[b] = arena_.take((size_t)nlow * 2 * Ll);
        lowt_base[b] = arena_.take((size_t)nlow * 2 * Ll);
        if (nhigh) {
            highext_base[b] = arena_.take((size_t)nhigh * 2 * LSh);
            hight_base[b] = arena_.take((size_t)nhigh * 2 * Lh);
        }
    }
    final_power_.assign(bic, std::vector<PowerLoc>(maxp + 1));

    const uint32_t persistent_top = (uint32_t)arena_.top;

    // ---- ComputePowers program (receiver_ddh.cpp:390-483) ----
    {
        ProgramBuilder pb(*this, powers_prog_);
        std::set<uint32_t> extended;
        auto &levels = dag_levels;
        powers_stage_end_.clear();
        for (uint32_t d = 1; d < levels.size(); d++) {
            arena_.top = persistent_top;
            const uint32_t n = (uint32_t)levels[d].size(), chunk = (n + part - 1) / part;
            std::vector<PowersNode> mine; // this rank's chunk of the level
            for (uint32_t k = 0; k < n; k++)
                if (k / chunk == prank) mine.push_back(levels[d][k]);
            std::vector<uint32_t> ext_cts, ext_dst, a, bb, prod, dst;
            std::set<uint32_t> need;
            for (auto &nd : mine) {
                if (!extended.count(nd.parent1)) need.insert(nd.parent1);
                if (!extended.count(nd.parent2)) need.insert(nd.parent2);
            }
            for (uint32_t b = 0; b < bic; b++) {
                if (!active[b]) continue;
                for (uint32_t e : need) {
                    ext_cts.push_back(loc[b][e]);
                    ext_dst.push_back(ext_loc[b][e]);
                }
            }
            extended.insert(need.begin(), need.end());
            uint32_t n_ops = 0;
            for (uint32_t b = 0; b < bic; b++)
                if (active[b]) n_ops += (uint32_t)mine.size();
            uint32_t prod0 = arena_.take((size_t)n_ops * 3 * Lf);
            uint32_t mscr = arena_.take(ProgramBuilder::multiply_scratch(ctx, Lf, n_ops));
            uint32_t rscr = arena_.take(ProgramBuilder::relin_scratch(Lf, n_ops));
            uint32_t o = 0;
            for (uint32_t b = 0; b < bic; b++) {
                if (!active[b]) continue;
                for (auto &nd : mine) {
                    a.push_back(ext_loc[b][nd.parent1]);
                    bb.push_back(ext_loc[b][nd.parent2]);
                    prod.push_back(prod0 + o * 3 * Lf);
                    dst.push_back(loc[b][nd.power]);
                    o++;
                }
            }
            // split PowersDag over peer memory: the first level must not overwrite a peer's level regions while that peer
            // still reads them (tail of its previous query); every level ends with the barrier that makes the mirrored
            // products of all ranks visible (run_peer_barrier is a no-op without set_powers_p2p)
            Engine *self = this;
            if (part > 1 && d == 1) pb.step([=] { self->run_peer_barrier(); });
            pb.extend(Lf, ext_cts, ext_dst);
            pb.multiply(Lf, a, bb, prod, mscr);
            pb.relinearize(Lf, prod, dst, rscr, /*mirror=*/part > 1); // relinearize == using_keyswitching (checked in ctor)
            if (part > 1) pb.step([=] { self->run_peer_barrier(); });
            powers_stage_end_.push_back(powers_prog_.size());
        }
        // tail: mod-switch every target to its level; low powers to NTT form (:446-478)
        arena_.top = persistent_top;
        std::vector<uint32_t> ntt_src, ntt_dst, hx_cts, hx_dst;
        // cur[b][e] = (arena idx, level) while switching down
        struct Cur {
            uint32_t idx, L, b, e, target_L;
            bool low;
        };
        std::vector<Cur> curs;
        for (uint32_t b = 0; b < bic; b++) {
            if (!active[b]) continue;
            for (uint32_t e : targets) {
                bool low = !ps || e <= ps;
                curs.push_back(Cur{ loc[b][e], Lf, b, e, low ? Ll : Lh, low });
            }
        }
        for (uint32_t L = Lf; L > 1; L--) {
            std::vector<uint32_t> ms_src, ms_dst;
            for (auto &c : curs) {
                if (c.L != L || c.target_L >= L) continue;
                uint32_t nxt = arena_.take(2 * (L - 1));
                for (uint32_t comp = 0; comp < 2; comp++) {
                    ms_src.push_back(c.idx + comp * L);
                    ms_dst.push_back(nxt + comp * (L - 1));
                }
                c.idx = nxt;
                c.L = L - 1;
            }
            pb.mod_switch_next(L, ms_src, ms_dst);
        }
        // note: the switched-down high powers live in the scratch area above persistent_top; everything the
        // evaluation needs from them (extended NTT forms) is produced below, before eval scratch reuses it —
        // except the coefficient-form high powers themselves, which no later step reads.
        for (auto &c : curs) {
            PowerLoc &fl = final_power_[c.b][c.e];
            fl.valid = true;
            fl.L = c.L;
            if (c.low) {
                uint32_t dstp = low_base[c.b] + (c.e - 1) * 2 * Ll;
                for (uint32_t k = 0; k < 2 * Ll; k++) {
                    ntt_src.push_back(c.idx + k);
                    ntt_dst.push_back(dstp + k);
                }
                fl.idx = dstp;
                fl.ntt = true;
            } else {
                uint32_t i = c.e / h;
                highc_loc[c.b][i] = c.idx;
                hx_cts.push_back(c.idx);
                hx_dst.push_back(highext_base[c.b] + (i - 1) * 2 * LSh);
                fl.idx = c.idx;
                fl.ntt = false;
            }
        }
        pb.ntt(ntt_src, ntt_dst, ctx.pattern_q(Ll), false);
        pb.extend(Lh, hx_cts, hx_dst);
        // tile-major split copies for the DB stream
        {
            std::vector<uint32_t> lsrc, ldst, hsrc, hdst;
            for (uint32_t b = 0; b < bic; b++) {
                if (!active[b]) continue;
                ldst.push_back(lowt_base[b]);
                for (uint32_t tc = 0; tc < nlow * 2; tc++) lsrc.push_back(low_base[b] + tc * Ll);
                if (nhigh) {
                    hdst.push_back(hight_base[b]);
                    for (uint32_t tc = 0; tc < nhigh * 2; tc++) hsrc.push_back(highext_base[b] + tc * LSh);
                }
            }
            pb.pack_powers(Ll, nlow, lsrc, ldst);
            pb.pack_powers(Lh, nhigh, hsrc, hdst);
        }
        powers_stage_end_.push_back(powers_prog_.size()); // the tail is the last stage
    }
    const uint32_t powers_top = (uint32_t)arena_.high_water;
    // coefficient-form high powers must survive for get_power(); keep eval scratch above the powers scratch
    // only if they were produced by a mod-switch (otherwise they alias persistent memory).
    const uint32_t eval_base = (Lf != Lh && nhigh) ? powers_top : persistent_top;

    // ---- evaluation program (receiver_ddh.cpp:485-535, bin_bundle.cpp:106-360) ----
    {
        ProgramBuilder pb(*this, eval_prog_);
        result_order_.clear();
        struct BRef {
            BinBundleStore *s;
            uint32_t k; // result slot
        };
        std::vector<BRef> all;
        for (uint32_t b = 0; b < bic; b++)
            for (auto &s : db[b]) {
                all.push_back(BRef{ s.get(), (uint32_t)result_order_.size() });
                result_order_.emplace_back(b, s->cache_idx);
            }
        results_.ensure((size_t)std::max<size_t>(all.size(), 1) * 2 * N);
        uint32_t alpha_max = 0;
        for (auto &v : db) alpha_max = std::max<uint32_t>(alpha_max, (uint32_t)v.size());
        npack_needed_ = alpha_max * bic;

        const uint32_t chunk = eval_chunk_; // measured on B200: larger batches win over L2 locality (launch tails dominate)
        const uint32_t drops = Ll - Lh;
        if (drops > 1) throw std::logic_error("unexpected level gap between low and high powers");

        // one accumulation job; emit_mac regroups them (kKtG to a group, equal lengths together)
        auto add_job = [&](std::vector<KtGroup> &gs, uint32_t p_idx, uint32_t pstride, const u64 *w, uint32_t wstride, uint32_t nterms, uint32_t out) {
            KtGroup g;
            std::memset(&g, 0, sizeof(g));
            g.p_idx = p_idx;
            g.pstride = pstride;
            g.w[0] = w;
            g.wstride[0] = wstride;
            g.nterms[0] = nterms;
            g.out_idx[0] = out;
            g.njobs = 1;
            g.max_terms = nterms;
            gs.push_back(g);
        };

        // ===== stage A (whole DB, one launch): every NTT-domain accumulation job of every BinBundle =====
        struct Job {
            BinBundleStore *s;
            uint32_t bslot; // index into psb
            uint32_t i, nterms;
        };
        struct PSB {
            BRef ref;
            uint32_t job_lo = 0, job_hi = 0;
        };
        std::vector<BRef> direct;
        std::vector<PSB> psb;
        std::vector<Job> jobs; // PS inner polynomials i >= 1, ordered by bundle
        for (auto &r : all) {
            uint32_t degree = r.s->ncoeffs - 1;
            bool using_ps = ps > 1 && ps < degree; // receiver_ddh.cpp:515-517
            if (!using_ps) {
                if (degree > nlow) throw std::logic_error("not enough ciphertext powers available");
                direct.push_back(r);
                continue;
            }
            PSB e;
            e.ref = r;
            e.job_lo = (uint32_t)jobs.size();
            uint32_t H = degree / h, rem = degree % h;
            for (uint32_t i = 1; i < H; i++) jobs.push_back(Job{ r.s, (uint32_t)psb.size(), i, h - 1 });
            if (rem) jobs.push_back(Job{ r.s, (uint32_t)psb.size(), H, rem });
            e.job_hi = (uint32_t)jobs.size();
            psb.push_back(e);
        }
        arena_.top = eval_base;
        const uint32_t nj_all = (uint32_t)jobs.size(), nb_all = (uint32_t)psb.size();
        const uint32_t acc0 = arena_.take((size_t)direct.size() * 2 * Ll); // direct evaluation accumulators
        const uint32_t tin0 = arena_.take((size_t)nj_all * 2 * Ll);        // inner polynomials i >= 1
        const uint32_t r00 = arena_.take((size_t)nb_all * 2 * Ll);        // sum of the i = 0 terms
        const uint32_t stageA_count = (uint32_t)arena_.top - acc0;
        std::vector<KtGroup> groups;
        std::vector<FinalizeJob> fin_direct;
        uint64_t mac_bytes = 0;
        for (size_t k = 0; k < direct.size(); k++) {
            BinBundleStore *s = direct[k].s;
            uint32_t degree = s->ncoeffs - 1, out = acc0 + (uint32_t)k * 2 * Ll;
            // bin_bundle.cpp:106-174; a degree-0 polynomial leaves the zero accumulator
            add_job(groups, lowt_base[s->bundle_idx], nlow, s->ntt_coeffs.p, s->n_ntt, degree, out);
            mac_bytes += (uint64_t)degree * Ll * N * 8;
            FinalizeJob f;
            std::memset(&f, 0, sizeof(f));
            f.src[0] = out;
            f.src[1] = f.src[2] = kNoSrc;
            f.coeff0 = s->plain_coeffs.p;
            f.pack = s->bundle_idx + s->cache_idx * bic;
            f.slot = direct[k].k;
            fin_direct.push_back(f);
        }
        for (uint32_t k = 0; k < nb_all; k++) {
            BinBundleStore *s = psb[k].ref.s;
            for (uint32_t j = psb[k].job_lo; j < psb[k].job_hi; j++) {
                // NTT-form rank of degree i*h + 1 is i*(h-1)
                const u64 *coeff = s->ntt_coeffs.p + (size_t)jobs[j].i * (h - 1) * kKtCols;
                add_job(groups, lowt_base[s->bundle_idx], nlow, coeff, s->n_ntt, jobs[j].nterms, tin0 + j * 2 * Ll);
                mac_bytes += (uint64_t)jobs[j].nterms * Ll * N * 8;
            }
            // i = 0 terms summed in NTT form.  Without a level gap this IS the i = 0 polynomial; with one
            // (per-term mod-switch, bin_bundle.cpp:314-324) it supplies the linear part and only the terms'
            // last-prime residues are handled one by one (k_ms_sum_last).
            add_job(groups, lowt_base[s->bundle_idx], nlow, s->ntt_coeffs.p, s->n_ntt, ps, r00 + k * 2 * Ll);
            mac_bytes += (uint64_t)ps * Ll * N * 8;
        }
        emit_mac(pb, Ll, groups, mac_bytes);
        // ===== stage B (whole DB): back to coefficient form =====
        pb.ntt_run(acc0, stageA_count, ctx.pattern_q(Ll), true);
        emit_finalize(pb, Ll, fin_direct);
        const uint32_t stage_top = (uint32_t)arena_.top;

        // ===== Paterson-Stockmeyer remainder (bin_bundle.cpp:192-360), `chunk` BinBundles at a time =====
        for (uint32_t c0 = 0; c0 < nb_all; c0 += chunk) {
            arena_.top = stage_top;
            const uint32_t c1 = std::min(nb_all, c0 + chunk), nb = c1 - c0;
            const uint32_t j0 = psb[c0].job_lo, j1 = psb[c1 - 1].job_hi, nj = j1 - j0;
            std::vector<FinalizeJob> fin_ps;
            std::vector<KtGroup> k8groups;
            std::vector<MulTermsJob> mulj;
            std::vector<uint32_t> tinh(nj), r0h(nb);
            if (drops) {
                // i = 0 polynomial: every term is mod-switched on its own (:314-324).  Exact shortcut: the
                // rounding of a term depends only on its last-prime residue, so only those ps*2 polynomials
                // are inverse-transformed per term; the other primes use the already inverted sum r00.
                const uint32_t t00 = arena_.take((size_t)nb * ps * 2);
                for (uint32_t k = 0; k < nb; k++) {
                    BinBundleStore *s = psb[c0 + k].ref.s;
                    MulTermsJob mj;
                    std::memset(&mj, 0, sizeof(mj));
                    mj.coeff = s->ntt_coeffs.p;
                    mj.rows = s->n_ntt;
                    mj.out_idx = t00 + k * ps * 2;
                    mj.pow_idx = low_base[s->bundle_idx];
                    mj.pow_term_stride = 2 * Ll;
                    mj.pow_comp_stride = Ll;
                    mj.nterms = ps;
                    mulj.push_back(mj);
                }
                emit_mul_terms(pb, Ll, mulj, ps);
                pb.ntt_run(t00, nb * ps * 2, { Ll - 1 }, true);
                const uint32_t tinh0 = arena_.take((size_t)nj * 2 * Lh);
                const uint32_t r0h0 = arena_.take((size_t)nb * 2 * Lh);
                std::vector<uint32_t> ms_src, ms_dst;
                for (uint32_t j = 0; j < nj; j++)
                    for (uint32_t c = 0; c < 2; c++) {
                        ms_src.push_back(tin0 + ((j0 + j) * 2 + c) * Ll);
                        ms_dst.push_back(tinh0 + (j * 2 + c) * Lh);
                    }
                pb.mod_switch_next(Ll, ms_src, ms_dst);
                {
                    std::vector<uint32_t> sum_idx(nb), last_idx(nb), dst_idx(nb);
                    for (uint32_t k = 0; k < nb; k++) {
                        sum_idx[k] = r00 + (c0 + k) * 2 * Ll;
                        last_idx[k] = t00 + k * ps * 2;
                        dst_idx[k] = r0h0 + k * 2 * Lh;
                    }
                    size_t so = idx_.add(sum_idx), lo = idx_.add(last_idx), dn = idx_.add(dst_idx);
                    pb.step([=] {
                        launch_pdl(k_ms_sum_last, dim3(ctx.N / kEwThreads, 2, nb), dim3(kEwThreads), 0, ctx.stream, arena_.buf.p, (const u32 *)idx_.at(so),
                                   (const u32 *)idx_.at(lo), (const u32 *)idx_.at(dn), ps, ctx.level[Ll], (int)ctx.N);
                        APSU_CUDA_CHECK(cudaGetLastError());
                        ctx.launches++;
                    });
                }
                for (uint32_t j = 0; j < nj; j++) tinh[j] = tinh0 + j * 2 * Lh;
                for (uint32_t k = 0; k < nb; k++) r0h[k] = r0h0 + k * 2 * Lh;
            } else {
                for (uint32_t j = 0; j < nj; j++) tinh[j] = tin0 + (j0 + j) * 2 * Ll;
                for (uint32_t k = 0; k < nb; k++) r0h[k] = r00 + (c0 + k) * 2 * Ll;
            }
            // inner polynomial x high power (:248-304)
            const uint32_t ext0 = arena_.take((size_t)nj * 2 * LSh);
            const uint32_t prod0 = arena_.take((size_t)nj * 3 * Lh);
            const uint32_t mscr = arena_.take(ProgramBuilder::multiply_scratch(ctx, Lh, nj));
            std::vector<uint32_t> exts(nj), hp(nj), prods(nj);
            for (uint32_t j = 0; j < nj; j++) {
                exts[j] = ext0 + j * 2 * LSh;
                prods[j] = prod0 + j * 3 * Lh;
                hp[j] = highext_base[jobs[j0 + j].s->bundle_idx] + (jobs[j0 + j].i - 1) * 2 * LSh;
            }
            pb.extend(Lh, tinh, exts);
            pb.multiply(Lh, exts, hp, prods, mscr);
            // sum of the products per bundle (size 3), one relinearisation per bundle (:308-310)
            const uint32_t res30 = arena_.take((size_t)nb * 3 * Lh), res20 = arena_.take((size_t)nb * 2 * Lh);
            {
                std::vector<std::vector<uint32_t>> terms(nb * 3);
                std::vector<uint32_t> sdst(nb * 3);
                for (uint32_t k = 0; k < nb; k++)
                    for (uint32_t c = 0; c < 3; c++) sdst[k * 3 + c] = res30 + (k * 3 + c) * Lh;
                for (uint32_t j = 0; j < nj; j++)
                    for (uint32_t c = 0; c < 3; c++) terms[(jobs[j0 + j].bslot - c0) * 3 + c].push_back(prods[j] + c * Lh);
                pb.sum_polys(Lh, terms, sdst);
            }
            std::vector<uint32_t> res3(nb), res2(nb);
            for (uint32_t k = 0; k < nb; k++) res3[k] = res30 + k * 3 * Lh, res2[k] = res20 + k * 2 * Lh;
            const uint32_t rscr = arena_.take(ProgramBuilder::relin_scratch(Lh, nb));
            pb.relinearize(Lh, res3, res2, rscr);
            // constant coefficients of the inner polynomials x high powers (:328-337): coefficient-form
            // operands; accumulated in NTT form at the high level, one inverse transform per bundle
            const uint32_t k80 = arena_.take((size_t)nb * 2 * Lh);
            for (uint32_t k = 0; k < nb; k++) {
                BinBundleStore *s = psb[c0 + k].ref.s;
                uint32_t H = (s->ncoeffs - 1) / h;
                add_job(k8groups, hight_base[s->bundle_idx], nhigh, s->plain_high_ntt.p, s->n_plain - 1, H, k80 + k * 2 * Lh);
            }
            emit_mac(pb, Lh, k8groups, 0);
            pb.ntt_run(k80, nb * 2 * Lh, ctx.pattern_q(Lh), true);
            for (uint32_t k = 0; k < nb; k++) {
                BinBundleStore *s = psb[c0 + k].ref.s;
                FinalizeJob f;
                std::memset(&f, 0, sizeof(f));
                f.src[0] = res2[k];
                f.src[1] = r0h[k];
                f.src[2] = k80 + k * 2 * Lh;
                f.coeff0 = s->plain_coeffs.p;
                f.pack = s->bundle_idx + s->cache_idx * bic;
                f.slot = psb[c0 + k].ref.k;
                fin_ps.push_back(f);
            }
            emit_finalize(pb, Lh, fin_ps);
        }
    }

    for (auto &g : fin_groups_) {
        APSU_CUDA_CHECK(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
        APSU_CUDA_CHECK(cudaEventCreateWithFlags(&g.copied, cudaEventDisableTiming));
    }
    arena_.buf.ensure(arena_.high_water * (size_t)N);
    idx_.upload(ctx.stream);
    desc_dev_.upload(desc_host_, ctx.stream);
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    plan_valid_ = true;
    query_loaded_ = powers_done_ = eval_done_ = false;
}

size_t Engine::add_desc(const void *data, size_t bytes)
{
    size_t off = (desc_host_.size() + 15) & ~size_t(15);
    desc_host_.resize(off + bytes);
    std::memcpy(desc_host_.data() + off, data, bytes);
    return off;
}

// cudaEventRecord that also works while the stream is being captured into a graph (an event-record node)
static cudaError_t record_event(cudaEvent_t ev, cudaStream_t st)
{
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaError_t e = cudaStreamIsCapturing(st, &cs);
    if (e != cudaSuccess) return e;
    return cudaEventRecordWithFlags(ev, st, cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault);
}

void Engine::drop_graphs()
{
    for (auto &g : powers_graphs_)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    powers_graphs_.clear();
    if (eval_graph_.exec) cudaGraphExecDestroy(eval_graph_.exec);
    eval_graph_ = ProgGraph();
}

// runs steps [lo, hi) of a program: captured into a graph on first use (and whenever a buffer the kernels were
// given has moved), replayed afterwards
void Engine::run_steps(std::vector<Step> &prog, size_t lo, size_t hi, ProgGraph &g)
{
    if (!use_graphs_ || !ctx.stream) { // the legacy default stream (NULL) cannot be captured: run eagerly on it
        for (size_t k = lo; k < hi; k++) prog[k].run();
        return;
    }
    const void *key[4] = { arena_.buf.p, masks_.p, relin_keys_.p, results_.p };
    const bool fresh = g.exec && g.profiling == profiling && std::equal(key, key + 4, g.key);
    if (!fresh) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = ProgGraph();
        const uint32_t l0 = ctx.launches;
        const size_t e0 = mac_events_used_;
        const uint64_t b0 = timed_mac_bytes_;
        APSU_CUDA_CHECK(cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeRelaxed));
        cudaGraph_t graph = nullptr;
        try {
            for (size_t k = lo; k < hi; k++) prog[k].run();
        } catch (...) {
            cudaStreamEndCapture(ctx.stream, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        APSU_CUDA_CHECK(cudaStreamEndCapture(ctx.stream, &graph));
        cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        APSU_CUDA_CHECK(e);
        g.launches = ctx.launches - l0;
        g.mac_events = mac_events_used_ - e0;
        g.mac_bytes = timed_mac_bytes_ - b0;
        std::copy(key, key + 4, g.key);
        g.profiling = profiling;
        // the capture ran the host side of the steps only; undo its bookkeeping, the launch below redoes it
        ctx.launches = l0;
        mac_events_used_ = e0;
        timed_mac_bytes_ = b0;
    }
    APSU_CUDA_CHECK(cudaGraphLaunch(g.exec, ctx.stream));
    ctx.launches += g.launches;
    mac_events_used_ += g.mac_events;
    timed_mac_bytes_ += g.mac_bytes;
}

void Engine::emit_mac(ProgramBuilder &pb, uint32_t L, std::vector<KtGroup> &groups, uint64_t bytes)
{
    if (groups.empty()) return;
    // regroup: jobs over the same power table, longest first, kKtG to a group; full equal-length groups run the
    // branch-free path of the kernel, the others (marked ragged) the masked one
    {
        struct J {
            const u64 *w;
            uint32_t wstride, nterms, out;
        };
        std::map<std::array<uint32_t, 2>, std::vector<J>> by_pow;
        for (auto &g : groups)
            for (uint32_t k = 0; k < g.njobs; k++) by_pow[{ g.p_idx, g.pstride }].push_back(J{ g.w[k], g.wstride[k], g.nterms[k], g.out_idx[k] });
        std::vector<KtGroup> out;
        for (auto &kv : by_pow) {
            auto &js = kv.second;
            std::stable_sort(js.begin(), js.end(), [](const J &a, const J &b) { return a.nterms > b.nterms; });
            for (size_t i = 0; i < js.size(); i += kKtG) {
                KtGroup g;
                std::memset(&g, 0, sizeof(g));
                g.p_idx = kv.first[0];
                g.pstride = kv.first[1];
                bool same = true;
                for (size_t k = i; k < std::min(js.size(), i + kKtG); k++) {
                    g.w[g.njobs] = js[k].w;
                    g.wstride[g.njobs] = js[k].wstride;
                    g.nterms[g.njobs] = js[k].nterms;
                    g.out_idx[g.njobs] = js[k].out;
                    g.max_terms = std::max(g.max_terms, js[k].nterms);
                    same &= js[k].nterms == js[i].nterms;
                    g.njobs++;
                }
                for (uint32_t k = g.njobs; k < (uint32_t)kKtG; k++) g.w[k] = g.w[0]; // never read: nterms = 0
                g.ragged = (same && g.njobs == (uint32_t)kKtG) ? 0 : 1;
                out.push_back(g);
            }
        }
        groups = out;
    }
    size_t off = add_desc(groups.data(), groups.size() * sizeof(KtGroup));
    uint32_t n = (uint32_t)groups.size();
    groups.clear();
    size_t slot = mac_step_bytes_.size();
    mac_step_bytes_.push_back(bytes);
    pb.step([=] {
        const bool timed = profiling && mac_step_bytes_[slot] > 0;
        cudaEvent_t a = nullptr, b2 = nullptr;
        if (timed) {
            if (mac_events_used_ == mac_events_.size()) {
                cudaEvent_t x, y;
                APSU_CUDA_CHECK(cudaEventCreate(&x));
                APSU_CUDA_CHECK(cudaEventCreate(&y));
                mac_events_.emplace_back(x, y);
            }
            a = mac_events_[mac_events_used_].first;
            b2 = mac_events_[mac_events_used_].second;
            mac_events_used_++;
            APSU_CUDA_CHECK(record_event(a, ctx.stream));
        }
        const KtGroup *gd = reinterpret_cast<const KtGroup *>(desc_dev_.p + off);
        // one persistent launch over all groups (shared-memory opt-in and occupancy: Engine ctor, per device)
        auto kern = k_db_mac_kt<kKtStages, kKtCtasPerSm>;
        constexpr size_t smem = kt_smem_bytes(kKtStages);
        // persistent grid, balanced: with W = ceil(items / resident CTAs) waves every CTA gets W (or W-1) items — a grid
        // of all resident CTAs would leave most of them idle in the last wave (256K-512: 320 items on 296 CTAs = two
        // waves for 8 % more work than one)
        const uint32_t items = n * (L * ctx.N / kKtCols);
        const uint32_t waves = (items + kt_grid_cap_ - 1) / kt_grid_cap_;
        const uint32_t grid = (items + waves - 1) / waves;
        launch_pdl(kern, dim3(grid), dim3(kKtThreads), smem, ctx.stream, arena_.buf.p, gd, n, ctx.level[L], (int)ctx.N, split_, fold_stages_, 0u);
        APSU_CUDA_CHECK(cudaGetLastError());
        ctx.launches++;
        if (timed) {
            APSU_CUDA_CHECK(record_event(b2, ctx.stream));
            timed_mac_bytes_ += mac_step_bytes_[slot];
        }
    });
}

void Engine::emit_mul_terms(ProgramBuilder &pb, uint32_t L, std::vector<MulTermsJob> &jobs, uint32_t nterms)
{
    if (jobs.empty()) return;
    size_t off = add_desc(jobs.data(), jobs.size() * sizeof(MulTermsJob));
    uint32_t n = (uint32_t)jobs.size();
    jobs.clear();
    pb.step([=] {
        launch_pdl(k_db_mul_last, dim3(ctx.N / kMacThreads, nterms, n), dim3(kMacThreads), 0, ctx.stream, arena_.buf.p,
                   reinterpret_cast<const MulTermsJob *>(desc_dev_.p + off), ctx.level[L], (int)ctx.N, split_);
        APSU_CUDA_CHECK(cudaGetLastError());
        ctx.launches++;
    });
}

void Engine::emit_finalize(ProgramBuilder &pb, uint32_t Ls, std::vector<FinalizeJob> &jobs)
{
    if (jobs.empty()) return;
    const apsu_b200_params &p = ctx.params;
    size_t off = add_desc(jobs.data(), jobs.size() * sizeof(FinalizeJob));
    uint32_t n = (uint32_t)jobs.size();
    // the result slots this launch completes: apsu_b200_eval_all_stream delivers them as soon as its event fires
    const size_t group = fin_groups_.size();
    fin_groups_.emplace_back();
    for (auto &j : jobs) fin_groups_.back().slots.push_back(j.slot);
    std::sort(fin_groups_.back().slots.begin(), fin_groups_.back().slots.end());
    jobs.clear();
    // try_clear_irrelevant_bits (bin_bundle.cpp:67-97): the last level always has exactly one prime
    int keep = hm::bit_length(ctx.t) + ((int)ctx.logN + 1) - 1;
    int drop = hm::bit_length(p.coeff_modulus[0]) - keep;
    u64 clear_mask = drop > 0 ? ~((1ull << drop) - 1) : ~0ull;
    pb.step([=] {
        launch_pdl(k_finalize, dim3(ctx.N / kEwThreads, 2, n), dim3(kEwThreads), 0, ctx.stream, arena_.buf.p,
                   reinterpret_cast<const FinalizeJob *>(desc_dev_.p + off), (const LevelConsts *)levels_dev_.p, (const u64 *)masks_.p, results_.p, (int)Ls, ctx.t,
                   clear_mask, (int)ctx.N);
        APSU_CUDA_CHECK(cudaGetLastError());
        ctx.launches++;
        if (fin_groups_[group].done) APSU_CUDA_CHECK(record_event(fin_groups_[group].done, ctx.stream));
    });
}

// ------------------------------------------------------------------------------------------------
// running the programs
// ------------------------------------------------------------------------------------------------
void Engine::compute_powers()
{
    if (powers_part_size_ > 1) throw std::logic_error("the PowersDag is split over several ranks: run compute_powers_stage and exchange between stages");
    const uint32_t n = powers_stage_count();
    for (uint32_t s = 0; s < n; s++) compute_powers_stage(s);
}

uint32_t Engine::powers_stage_count()
{
    if (!plan_valid_) build_plan();
    return (uint32_t)powers_stage_end_.size();
}

// stage s < depth: the products of DAG level s+1 this rank owns; last stage: mod-switches, NTTs, power tables.
// With a split DAG the caller all-gathers the exchange regions of level s+1 between the ranks after stage s.
void Engine::compute_powers_stage(uint32_t stage)
{
    if (!plan_valid_ || !query_loaded_) throw std::logic_error("compute_powers called before query_begin (or the DB changed since)");
    if (ctx.using_keyswitching() && dag.depth() > 0 && !have_keys_) throw std::invalid_argument("relinearization keys have not been set");
    if (stage >= powers_stage_end_.size()) throw std::invalid_argument("powers stage is out of range");
    if (stage == 0) {
        ctx.launches = 0;
        APSU_CUDA_CHECK(cudaEventRecord(ev_[0], ctx.stream));
        powers_done_ = eval_done_ = false;
    }
    const size_t lo = stage ? powers_stage_end_[stage - 1] : 0, hi = powers_stage_end_[stage];
    if (powers_graphs_.size() != powers_stage_end_.size()) {
        drop_graphs();
        powers_graphs_.resize(powers_stage_end_.size());
    }
    run_steps(powers_prog_, lo, hi, powers_graphs_[stage]);
    if (stage + 1 == powers_stage_end_.size()) {
        APSU_CUDA_CHECK(cudaEventRecord(ev_[1], ctx.stream));
        powers_done_ = true;
        powers_launches_ = ctx.launches;
    }
}

void Engine::set_powers_partition(uint32_t rank, uint32_t size)
{
    if (!size || rank >= size) throw std::invalid_argument("powers partition: rank must be below size");
    powers_part_rank_ = rank;
    powers_part_size_ = size;
    invalidate_plan();
}

// exchange regions of one DAG level (one per active bundle index): device pointer of the all-gather buffer
// [size][chunk_bytes] and the chunk size; this rank's chunk is at rank*chunk_bytes
uint32_t Engine::powers_exchange_regions(uint32_t level, void **ptrs, uint64_t *chunk_bytes, uint32_t capacity)
{
    if (!plan_valid_) build_plan();
    uint32_t n = 0;
    for (auto &r : powers_exchange_) {
        if (r.level != level) continue;
        if (n < capacity) {
            if (ptrs) ptrs[n] = arena_.buf.p + (size_t)r.region * ctx.N;
            if (chunk_bytes) chunk_bytes[n] = (uint64_t)r.chunk_polys * ctx.N * 8;
        }
        n++;
    }
    return n;
}

void Engine::eval_all()
{
    if (!plan_valid_ || !powers_done_) throw std::logic_error("eval_all called before compute_powers");
    if (npack_ < npack_needed_) throw std::invalid_argument("mask table is smaller than alpha_max_cache_count * bundle_idx_count");
    ctx.launches = 0;
    mac_events_used_ = 0;
    timed_mac_bytes_ = 0;
    join_masks_upload();
    APSU_CUDA_CHECK(cudaEventRecord(ev_[2], ctx.stream));
    run_steps(eval_prog_, 0, eval_prog_.size(), eval_graph_);
    APSU_CUDA_CHECK(cudaEventRecord(ev_[3], ctx.stream));
    eval_done_ = true;
    eval_launches_ = ctx.launches;
}

void Engine::collect_timings()
{
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    std::memset(&timings, 0, sizeof(timings));
    if (powers_done_) {
        float ms = 0;
        APSU_CUDA_CHECK(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
        timings.compute_powers_ms = ms;
    }
    if (eval_done_) {
        float ms = 0;
        APSU_CUDA_CHECK(cudaEventElapsedTime(&ms, ev_[2], ev_[3]));
        timings.eval_ms = ms;
        float mac = 0;
        for (size_t k = 0; k < mac_events_used_; k++) {
            float x = 0;
            APSU_CUDA_CHECK(cudaEventElapsedTime(&x, mac_events_[k].first, mac_events_[k].second));
            mac += x;
        }
        timings.db_stream_ms = mac;
        timings.db_stream_bytes = timed_mac_bytes_;
        timings.db_stream_launches = (uint32_t)mac_events_used_;
    }
    timings.kernel_launches = powers_launches_ + eval_launches_;
}

void Engine::fetch_results(uint64_t *out, uint32_t *bundle_idx, uint32_t *cache_idx)
{
    if (!eval_done_) throw std::logic_error("fetch_results called before eval_all");
    size_t n = result_order_.size();
    enqueue_flag_read(); // is_valid_for of the query that produced these results: one synchronisation for both
    if (out && n) APSU_CUDA_CHECK(cudaMemcpyAsync(out, results_.p, n * 2 * ctx.N * 8, cudaMemcpyDeviceToHost, ctx.stream));
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    check_flags_after_sync();
    for (size_t k = 0; k < n; k++) {
        if (bundle_idx) bundle_idx[k] = result_order_[k].first;
        if (cache_idx) cache_idx[k] = result_order_[k].second;
    }
}

// ProcessBinBundleCache for every BinBundle with delivery per BinBundle, as the reference hands every ResultPackage to
// the channel the moment its BinBundle is done (receiver_ddh.cpp:527-534): each finalize launch of the evaluation
// records an event; a second stream waits for it and copies that launch's result ciphertexts to the host; the calling
// thread invokes `fn` for them as soon as the copy lands, while the device is still evaluating later chunks.
void Engine::eval_all_stream(uint64_t *out, void (*fn)(void *, uint32_t, uint32_t, const uint64_t *), void *user)
{
    if (!out) throw std::invalid_argument("eval_all_stream: out is null");
    if (!copy_stream_) APSU_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    eval_all();
    const size_t per = (size_t)2 * ctx.N;
    for (auto &g : fin_groups_) {
        APSU_CUDA_CHECK(cudaStreamWaitEvent(copy_stream_, g.done, 0));
        for (size_t a = 0; a < g.slots.size();) { // consecutive slots in one copy
            size_t b = a + 1;
            while (b < g.slots.size() && g.slots[b] == g.slots[b - 1] + 1) b++;
            APSU_CUDA_CHECK(cudaMemcpyAsync(out + g.slots[a] * per, results_.p + g.slots[a] * per, (b - a) * per * 8, cudaMemcpyDeviceToHost, copy_stream_));
            a = b;
        }
        APSU_CUDA_CHECK(cudaEventRecord(g.copied, copy_stream_));
    }
    for (auto &g : fin_groups_) {
        APSU_CUDA_CHECK(cudaEventSynchronize(g.copied));
        if (fn)
            for (uint32_t slot : g.slots) fn(user, result_order_[slot].first, result_order_[slot].second, out + slot * per);
    }
    throw_if_query_invalid();
}

void Engine::results_device(void **ptr, uint64_t *bytes)
{
    if (!plan_valid_) build_plan();
    if (ptr) *ptr = results_.p;
    if (bytes) *bytes = (uint64_t)result_order_.size() * 2 * ctx.N * 8;
}

void Engine::get_power(uint32_t bundle_idx, uint32_t power, uint64_t *out, uint32_t *L, int *is_ntt)
{
    if (!powers_done_) throw std::logic_error("get_power called before compute_powers");
    if (bundle_idx >= final_power_.size() || power >= final_power_[bundle_idx].size() || !final_power_[bundle_idx][power].valid)
        throw std::invalid_argument("power is not available");
    const PowerLoc &pl = final_power_[bundle_idx][power];
    if (L) *L = pl.L;
    if (is_ntt) *is_ntt = pl.ntt ? 1 : 0;
    if (out) {
        APSU_CUDA_CHECK(cudaMemcpyAsync(out, arena_.buf.p + (size_t)pl.idx * ctx.N, (size_t)2 * pl.L * ctx.N * 8, cudaMemcpyDeviceToHost, ctx.stream));
        APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    }
}

// ------------------------------------------------------------------------------------------------
// stand-alone batched evaluator operations (tests, and the "SEAL Evaluator calls" row of the scope)
// ------------------------------------------------------------------------------------------------
void Engine::op_ntt(uint64_t *polys, uint32_t count, const uint32_t *modulus_index, uint32_t pattern_len, bool inverse)
{
    if (!polys || !modulus_index || !pattern_len) throw std::invalid_argument("op_ntt: bad arguments");
    std::vector<uint32_t> pattern(modulus_index, modulus_index + pattern_len);
    DBuf<u64> buf;
    buf.alloc((size_t)count * ctx.N);
    APSU_CUDA_CHECK(cudaMemcpyAsync(buf.p, polys, buf.n * 8, cudaMemcpyHostToDevice, ctx.stream));
    ctx.ntt(buf.p, buf.p, count, pattern, inverse);
    APSU_CUDA_CHECK(cudaMemcpyAsync(polys, buf.p, buf.n * 8, cudaMemcpyDeviceToHost, ctx.stream));
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
}

template <typename Build>
void Engine::run_scratch_program(Build &&build)
{
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    Arena saved_arena = std::move(arena_);
    IdxPool saved_idx = std::move(idx_);
    arena_ = Arena();
    idx_ = IdxPool();
    std::vector<Step> prog;
    try {
        build(prog);
    } catch (...) {
        arena_ = std::move(saved_arena);
        idx_ = std::move(saved_idx);
        throw;
    }
    arena_ = std::move(saved_arena);
    idx_ = std::move(saved_idx);
}

// SEAL's blake2xb generator as a raw stream (refills counter0, counter0+1, ...) and sample_poly_uniform (tests)
void Engine::op_prng_stream(const uint8_t *seed64, uint64_t counter0, uint64_t *out, size_t n_words)
{
    if (!seed64 || !out || !n_words) throw std::invalid_argument("op_prng_stream: bad arguments");
    PrngSeed seed;
    std::memcpy(seed.w, seed64, sizeof(seed.w));
    DBuf<u64> &buf = aux_[0];
    buf.ensure(n_words);
    k_prng_stream<<<(unsigned)((n_words + 511) / 512), 64, 0, ctx.stream>>>(buf.p, seed, counter0, n_words);
    APSU_LAUNCH_CHECK();
    APSU_CUDA_CHECK(cudaMemcpyAsync(out, buf.p, n_words * 8, cudaMemcpyDeviceToHost, ctx.stream));
    APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
}
void Engine::op_expand_seeds(uint32_t L, const uint8_t *seeds64, uint32_t n, uint64_t *out)
{
    if (L < 1 || L > ctx.K || L > (uint32_t)kMaxQ) throw std::invalid_argument("op_expand_seeds: level is out of range");
    if (!seeds64 || !out || !n) throw std::invalid_argument("op_expand_seeds: bad arguments");
    DBuf<u64> &buf = aux_[0];
    buf.ensure((size_t)n * L * ctx.N);
    std::vector<uint32_t> dst(n);
    for (uint32_t k = 0; k < n; k++) dst[k] = k * L;
    expand_seeds(L, dst, seeds64, buf.p, ctx.params.coeff_modulus);
    APSU_CUDA_CHECK(cudaMemcpyAsync(out, buf.p, (size_t)n * L * ctx.N * 8, cudaMemcpyDeviceToHost, ctx.stream));
    throw_if_query_invalid();
}

void Engine::op_multiply(uint32_t L, const uint64_t *a, const uint64_t *b, uint64_t *out, uint32_t n_ops)
{
    if (L < 1 || L > ctx.first_L) throw std::invalid_argument("op_multiply: level is out of range");
    if (!a || !b || !out || !n_ops) throw std::invalid_argument("op_multiply: bad arguments");
    const uint32_t N = ctx.N, LS = L + (uint32_t)ctx.level[L].S;
    run_scratch_program([&](std::vector<Step> &prog) {
        ProgramBuilder pb(*this, prog);
        uint32_t a0 = arena_.take((size_t)n_ops * 2 * L), b0 = arena_.take((size_t)n_ops * 2 * L);
        uint32_t ea = arena_.take((size_t)n_ops * 2 * LS), eb = arena_.take((size_t)n_ops * 2 * LS);
        uint32_t pr = arena_.take((size_t)n_ops * 3 * L);
        uint32_t scr = arena_.take(ProgramBuilder::multiply_scratch(ctx, L, n_ops));
        std::vector<uint32_t> cts, exts, av, bv, pv;
        for (uint32_t o = 0; o < n_ops; o++) {
            cts.push_back(a0 + o * 2 * L);
            exts.push_back(ea + o * 2 * LS);
            av.push_back(ea + o * 2 * LS);
            bv.push_back(eb + o * 2 * LS);
            pv.push_back(pr + o * 3 * L);
        }
        for (uint32_t o = 0; o < n_ops; o++) {
            cts.push_back(b0 + o * 2 * L);
            exts.push_back(eb + o * 2 * LS);
        }
        pb.extend(L, cts, exts);
        pb.multiply(L, av, bv, pv, scr);
        arena_.buf.alloc(arena_.high_water * (size_t)N);
        idx_.upload(ctx.stream);
        size_t w = (size_t)n_ops * 2 * L * N;
        APSU_CUDA_CHECK(cudaMemcpyAsync(arena_.buf.p + (size_t)a0 * N, a, w * 8, cudaMemcpyHostToDevice, ctx.stream));
        APSU_CUDA_CHECK(cudaMemcpyAsync(arena_.buf.p + (size_t)b0 * N, b, w * 8, cudaMemcpyHostToDevice, ctx.stream));
        for (auto &s : prog) s.run();
        APSU_CUDA_CHECK(cudaMemcpyAsync(out, arena_.buf.p + (size_t)pr * N, (size_t)n_ops * 3 * L * N * 8, cudaMemcpyDeviceToHost, ctx.stream));
        APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    });
}

void Engine::op_relinearize(uint32_t L, const uint64_t *in, uint64_t *out, uint32_t n_ops)
{
    if (L < 1 || L > ctx.first_L) throw std::invalid_argument("op_relinearize: level is out of range");
    if (!ctx.using_keyswitching()) throw std::logic_error("parameters do not support key switching");
    if (!have_keys_) throw std::invalid_argument("relinearization keys have not been set");
    if (!in || !out || !n_ops) throw std::invalid_argument("op_relinearize: bad arguments");
    const uint32_t N = ctx.N;
    run_scratch_program([&](std::vector<Step> &prog) {
        ProgramBuilder pb(*this, prog);
        uint32_t i0 = arena_.take((size_t)n_ops * 3 * L), o0 = arena_.take((size_t)n_ops * 2 * L);
        uint32_t scr = arena_.take(ProgramBuilder::relin_scratch(L, n_ops));
        std::vector<uint32_t> iv, ov;
        for (uint32_t o = 0; o < n_ops; o++) iv.push_back(i0 + o * 3 * L), ov.push_back(o0 + o * 2 * L);
        pb.relinearize(L, iv, ov, scr);
        arena_.buf.alloc(arena_.high_water * (size_t)N);
        idx_.upload(ctx.stream);
        APSU_CUDA_CHECK(cudaMemcpyAsync(arena_.buf.p + (size_t)i0 * N, in, (size_t)n_ops * 3 * L * N * 8, cudaMemcpyHostToDevice, ctx.stream));
        for (auto &s : prog) s.run();
        APSU_CUDA_CHECK(cudaMemcpyAsync(out, arena_.buf.p + (size_t)o0 * N, (size_t)n_ops * 2 * L * N * 8, cudaMemcpyDeviceToHost, ctx.stream));
        APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    });
}

void Engine::op_mod_switch_next(uint32_t L, const uint64_t *in, uint64_t *out, uint32_t n_polys)
{
    if (L < 2 || L > ctx.first_L) throw std::invalid_argument("op_mod_switch_next: level is out of range");
    if (!in || !out || !n_polys) throw std::invalid_argument("op_mod_switch_next: bad arguments");
    const uint32_t N = ctx.N;
    run_scratch_program([&](std::vector<Step> &prog) {
        ProgramBuilder pb(*this, prog);
        uint32_t i0 = arena_.take((size_t)n_polys * L), o0 = arena_.take((size_t)n_polys * (L - 1));
        std::vector<uint32_t> iv, ov;
        for (uint32_t k = 0; k < n_polys; k++) iv.push_back(i0 + k * L), ov.push_back(o0 + k * (L - 1));
        pb.mod_switch_next(L, iv, ov);
        arena_.buf.alloc(arena_.high_water * (size_t)N);
        idx_.upload(ctx.stream);
        APSU_CUDA_CHECK(cudaMemcpyAsync(arena_.buf.p + (size_t)i0 * N, in, (size_t)n_polys * L * N * 8, cudaMemcpyHostToDevice, ctx.stream));
        for (auto &s : prog) s.run();
        APSU_CUDA_CHECK(cudaMemcpyAsync(out, arena_.buf.p + (size_t)o0 * N, (size_t)n_polys * (L - 1) * N * 8, cudaMemcpyDeviceToHost, ctx.stream));
        APSU_CUDA_CHECK(cudaStreamSynchronize(ctx.stream));
    });
}

} // namespace apsu_b200
