// Multi-GPU host path of the receiver's query evaluation, in C++ behind the C ABI (include/apsu_b200.h, "multi-GPU"):
// one apsu_b200_mgpu per rank/GPU over one NCCL communicator.  BinBundles are independent units
// (receiver/apsu/receiver_ddh.cpp:340-364), so the DB is sharded by BinBundle and the data path has three exchanges
// (SURVEY.md §2.2 C1-C3, §8e):
//   C1  the rank that received the query (root) uploads it in per-bundle-index chunks and sends every rank ONLY the
//       ciphertexts of the bundle indices it owns (ncclSend/ncclRecv), the relinearisation keys go to all (ncclBroadcast);
//       the upload of chunk b+1 overlaps the sends of chunk b (copy stream + events);
//   C2  ranks that share ONE bundle index split its PowersDag and all-gather every DAG level in place
//       (ncclAllGather on a sub-communicator, ncclCommSplit) — optional, see apsu_b200_mgpu_commit;
//   C3  result ciphertexts are sent to root unpadded (ncclSend/ncclRecv) and leave through one device-to-host copy.
// NCCL is bound at run time (dlopen of libnccl.so.2): the library has no link-time NCCL dependency, a process that
// already loaded an NCCL (e.g. torch's) shares it, and single-GPU users need none.
#include "../../include/apsu_b200.h"
#include "engine.hpp"
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>
#include <algorithm>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <set>
#include <string>

using namespace apsu_b200;

namespace apsu_b200 {
Engine &engine_of(apsu_b200_ctx *ctx); // capi.cu
int guarded_call(const std::function<void()> &f); // capi.cu: exception -> status + last_error
}

namespace {

struct Nccl {
    void *so = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclCommSplit) CommSplit = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};

Nccl &nccl()
{
    static Nccl n;
    static std::once_flag once;
    static std::string err;
    std::call_once(once, [] {
        for (const char *name : { "libnccl.so.2", "libnccl.so" }) {
            n.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.so) break;
        }
        if (!n.so) {
            err = std::string("NCCL is not available (dlopen libnccl.so.2): ") + dlerror();
            return;
        }
#define APSU_NCCL_SYM(field, sym)                                                                                      \
    n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.so, #sym));                                                  \
    if (!n.field) err = "NCCL symbol missing: " #sym;
        APSU_NCCL_SYM(GetUniqueId, ncclGetUniqueId)
        APSU_NCCL_SYM(CommInitRank, ncclCommInitRank)
        APSU_NCCL_SYM(CommDestroy, ncclCommDestroy)
        APSU_NCCL_SYM(CommSplit, ncclCommSplit)
        APSU_NCCL_SYM(Broadcast, ncclBroadcast)
        APSU_NCCL_SYM(AllGather, ncclAllGather)
        APSU_NCCL_SYM(Send, ncclSend)
        APSU_NCCL_SYM(Recv, ncclRecv)
        APSU_NCCL_SYM(GroupStart, ncclGroupStart)
        APSU_NCCL_SYM(GroupEnd, ncclGroupEnd)
        APSU_NCCL_SYM(GetErrorString, ncclGetErrorString)
        APSU_NCCL_SYM(GetVersion, ncclGetVersion)
#undef APSU_NCCL_SYM
    });
    if (!err.empty()) throw std::runtime_error(err);
    return n;
}

#define APSU_NCCL_CHECK(expr)                                                                                          \
    do {                                                                                                               \
        ncclResult_t r__ = (expr);                                                                                     \
        if (r__ != ncclSuccess) throw std::runtime_error(std::string(#expr) + ": " + nccl().GetErrorString(r__));      \
    } while (0)

} // namespace

struct apsu_b200_mgpu {
    Engine *eng = nullptr;
    uint32_t rank = 0, world = 1, root = 0;
    ncclComm_t comm = nullptr, part_comm = nullptr;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> chunk_ev;
    cudaEvent_t ready_ev = nullptr;
    // committed layout
    bool committed = false;
    std::vector<std::vector<uint32_t>> owned;       // [rank] -> bundle indices it holds BinBundles of (ascending)
    std::vector<uint32_t> counts;                   // [rank] -> number of BinBundles
    std::vector<uint32_t> all_bundle_idx, all_cache_idx; // global (bundle_idx, cache_idx) of every result, rank-major
    uint32_t part_rank = 0, part_size = 1;
    // split PowersDag over NVLink peer memory: mapped arenas / flag arrays of the group (own entry = own pointers)
    bool p2p = false;
    std::vector<void *> ipc_opened;
    u64 *arena_at_commit = nullptr;
    // staging
    DBuf<u64> q_stage;  // root: [bundle_idx_count][nsrc][2][L][N]; others: [owned][nsrc][2][L][N]
    DBuf<u64> gathered; // root: [total][2][N]
    DBuf<uint32_t> meta;
};

namespace {

void destroy(apsu_b200_mgpu *m)
{
    if (!m) return;
    if (m->eng) cudaSetDevice(m->eng->ctx.device);
    for (void *q : m->ipc_opened) cudaIpcCloseMemHandle(q);
    if (m->part_comm) nccl().CommDestroy(m->part_comm);
    if (m->comm) nccl().CommDestroy(m->comm);
    for (auto e : m->chunk_ev) cudaEventDestroy(e);
    if (m->ready_ev) cudaEventDestroy(m->ready_ev);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    delete m;
}

// all-gather of a fixed number of 32-bit words per rank through the communicator (setup-time metadata)
std::vector<uint32_t> allgather_u32(apsu_b200_mgpu &m, const std::vector<uint32_t> &mine)
{
    const size_t n = mine.size();
    cudaStream_t st = m.eng->ctx.stream;
    m.meta.ensure(n * (m.world + 1));
    APSU_CUDA_CHECK(cudaMemcpyAsync(m.meta.p, mine.data(), n * 4, cudaMemcpyHostToDevice, st));
    APSU_NCCL_CHECK(nccl().AllGather(m.meta.p, m.meta.p + n, n, ncclUint32, m.comm, st));
    std::vector<uint32_t> all(n * m.world);
    APSU_CUDA_CHECK(cudaMemcpyAsync(all.data(), m.meta.p + n, all.size() * 4, cudaMemcpyDeviceToHost, st));
    APSU_CUDA_CHECK(cudaStreamSynchronize(st));
    return all;
}

// Peer-memory exchange of the split PowersDag: every rank of the group publishes its arena and its barrier flags
// (CUDA IPC handles between processes, plain pointers + cudaDeviceEnablePeerAccess between threads of one process),
// maps the others', and hands the tables to its engine.  Any rank that cannot (no peer access, different arena layout)
// makes the whole group fall back to the ncclAllGather exchange.
struct P2PInfo {
    unsigned long long pid, arena_ptr, flags_ptr, layout;
    int device, pad;
    cudaIpcMemHandle_t arena_h, flags_h;
};
void setup_p2p(apsu_b200_mgpu &m)
{
    Engine &e = *m.eng;
    for (void *q : m.ipc_opened) cudaIpcCloseMemHandle(q);
    m.ipc_opened.clear();
    m.p2p = false;
    e.set_powers_p2p({}, {});
    if (m.part_size < 2 || m.part_size > 8) return;
    if (const char *ev = std::getenv("APSU_B200_NO_P2P"))
        if (atoi(ev)) return;
    cudaStream_t st = e.ctx.stream;
    P2PInfo mine;
    std::memset(&mine, 0, sizeof(mine));
    mine.pid = (unsigned long long)getpid();
    mine.device = e.ctx.device;
    mine.arena_ptr = (unsigned long long)e.arena_base(); // builds the plan: the arena is final until the DB changes
    mine.flags_ptr = (unsigned long long)e.p2p_flags();
    {   // the exchange regions must sit at the same arena offsets on every rank of the group
        std::vector<void *> ptrs(e.ctx.params.bundle_idx_count);
        std::vector<uint64_t> bytes(e.ctx.params.bundle_idx_count);
        unsigned long long h = 1469598103934665603ull;
        for (uint32_t lv = 1; lv <= e.dag.depth(); lv++) {
            const uint32_t n = e.powers_exchange_regions(lv, ptrs.data(), bytes.data(), (uint32_t)ptrs.size());
            for (uint32_t k = 0; k < n; k++)
                for (unsigned long long v : { (unsigned long long)((char *)ptrs[k] - (char *)e.arena_base()), (unsigned long long)bytes[k] }) h = (h ^ v) * 1099511628211ull;
        }
        mine.layout = h;
    }
    bool ok = cudaIpcGetMemHandle(&mine.arena_h, (void *)mine.arena_ptr) == cudaSuccess && cudaIpcGetMemHandle(&mine.flags_h, (void *)mine.flags_ptr) == cudaSuccess;
    cudaGetLastError();
    // all-gather the records inside the group
    const size_t rec = sizeof(P2PInfo);
    DBuf<unsigned char> buf;
    buf.alloc(rec * (m.part_size + 1));
    APSU_CUDA_CHECK(cudaMemcpyAsync(buf.p, &mine, rec, cudaMemcpyHostToDevice, st));
    APSU_NCCL_CHECK(nccl().AllGather(buf.p, buf.p + rec, rec, ncclUint8, m.part_comm, st));
    std::vector<P2PInfo> all(m.part_size);
    APSU_CUDA_CHECK(cudaMemcpyAsync(all.data(), buf.p + rec, rec * m.part_size, cudaMemcpyDeviceToHost, st));
    APSU_CUDA_CHECK(cudaStreamSynchronize(st));
    std::vector<u64 *> arenas(m.part_size, nullptr);
    std::vector<uint32_t *> flags(m.part_size, nullptr);
    for (uint32_t k = 0; k < m.part_size && ok; k++) {
        if (all[k].layout != mine.layout) ok = false;
        if (k == m.part_rank) {
            arenas[k] = (u64 *)mine.arena_ptr;
            flags[k] = (uint32_t *)mine.flags_ptr;
            continue;
        }
        if (all[k].pid == mine.pid) { // a thread of this process: direct peer access
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, e.ctx.device, all[k].device) != cudaSuccess || !can) ok = false;
            else {
                cudaError_t r = cudaDeviceEnablePeerAccess(all[k].device, 0);
                if (r != cudaSuccess && r != cudaErrorPeerAccessAlreadyEnabled) ok = false;
                cudaGetLastError();
                arenas[k] = (u64 *)all[k].arena_ptr;
                flags[k] = (uint32_t *)all[k].flags_ptr;
            }
        } else {
            void *a = nullptr, *f = nullptr;
            if (cudaIpcOpenMemHandle(&a, all[k].arena_h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
            else m.ipc_opened.push_back(a);
            if (ok && cudaIpcOpenMemHandle(&f, all[k].flags_h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
            else if (ok) m.ipc_opened.push_back(f);
            cudaGetLastError();
            arenas[k] = (u64 *)a;
            flags[k] = (uint32_t *)f;
        }
    }
    // unanimous or not at all
    std::vector<uint32_t> votes;
    {
        DBuf<uint32_t> v;
        v.alloc(m.part_size + 1);
        const uint32_t my = ok ? 1u : 0u;
        APSU_CUDA_CHECK(cudaMemcpyAsync(v.p, &my, 4, cudaMemcpyHostToDevice, st));
        APSU_NCCL_CHECK(nccl().AllGather(v.p, v.p + 1, 1, ncclUint32, m.part_comm, st));
        votes.resize(m.part_size);
        APSU_CUDA_CHECK(cudaMemcpyAsync(votes.data(), v.p + 1, 4 * m.part_size, cudaMemcpyDeviceToHost, st));
        APSU_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    for (uint32_t x : votes) ok &= x == 1;
    if (!ok) {
        for (void *q : m.ipc_opened) cudaIpcCloseMemHandle(q);
        m.ipc_opened.clear();
        return;
    }
    e.set_powers_p2p(arenas, flags);
    m.arena_at_commit = (u64 *)mine.arena_ptr;
    m.p2p = true;
}

void commit(apsu_b200_mgpu &m, const uint32_t *global_cache_idx, int dag_split)
{
    Engine &e = *m.eng;
    const auto &order = e.result_order(); // (bundle_idx, local cache_idx), the order of the result buffer
    const uint32_t n_local = (uint32_t)order.size();
    // 1. counts
    std::vector<uint32_t> cnt = allgather_u32(m, { n_local });
    m.counts = cnt;
    uint32_t max_n = 1, total = 0;
    for (uint32_t c : cnt) max_n = std::max(max_n, c), total += c;
    // 2. (bundle_idx, global cache_idx) lists, padded to the longest
    std::vector<uint32_t> mine(2 * max_n, 0xFFFFFFFFu);
    for (uint32_t k = 0; k < n_local; k++) {
        mine[2 * k] = order[k].first;
        mine[2 * k + 1] = global_cache_idx ? global_cache_idx[k] : order[k].second;
    }
    std::vector<uint32_t> all = allgather_u32(m, mine);
    m.owned.assign(m.world, {});
    m.all_bundle_idx.clear();
    m.all_cache_idx.clear();
    for (uint32_t r = 0; r < m.world; r++) {
        std::set<uint32_t> idx;
        for (uint32_t k = 0; k < cnt[r]; k++) {
            idx.insert(all[(size_t)r * 2 * max_n + 2 * k]);
            m.all_bundle_idx.push_back(all[(size_t)r * 2 * max_n + 2 * k]);
            m.all_cache_idx.push_back(all[(size_t)r * 2 * max_n + 2 * k + 1]);
        }
        m.owned[r].assign(idx.begin(), idx.end());
    }
    // 3. PowersDag partition (C2): ranks that each hold exactly one bundle index, the same one, form a group.
    // Worth it only for large DAGs (a level's all-gather costs more than half a small level's products): auto = at
    // least 128 products per bundle index (measured: profiles/README.md).
    if (m.part_comm) {
        nccl().CommDestroy(m.part_comm);
        m.part_comm = nullptr;
    }
    bool eligible = true;
    for (auto &o : m.owned) eligible &= o.size() == 1;
    std::vector<uint32_t> group;
    if (eligible)
        for (uint32_t r = 0; r < m.world; r++)
            if (m.owned[r][0] == m.owned[m.rank][0]) group.push_back(r);
    size_t products = 0;
    for (auto &lv : e.dag.levels()) products += lv.size();
    products -= e.dag.levels().empty() ? 0 : e.dag.levels()[0].size();
    // auto (-1): split when the levels can be exchanged through peer memory (cheap: measured 0.62 -> 0.53 ms per bundle
    // index of 16M-4096 on 2 GPUs) or when the DAG is large enough to pay for ncclAllGather per level (>= 128 products)
    const bool want = dag_split != 0;
    bool split = eligible && want && e.dag.depth() > 0;
    // every rank must take the same decision about calling ncclCommSplit: `eligible` and `want` are global facts
    bool any_group = false;
    if (split) {
        std::map<uint32_t, uint32_t> sizes;
        for (auto &o : m.owned) sizes[o[0]]++;
        for (auto &kv : sizes) any_group |= kv.second > 1;
    }
    if (split && any_group) {
        const int color = group.size() > 1 ? (int)m.owned[m.rank][0] : NCCL_SPLIT_NOCOLOR;
        APSU_NCCL_CHECK(nccl().CommSplit(m.comm, color, (int)m.rank, &m.part_comm, nullptr));
    }
    if (m.part_comm && group.size() > 1) {
        m.part_size = (uint32_t)group.size();
        m.part_rank = (uint32_t)(std::find(group.begin(), group.end(), m.rank) - group.begin());
    } else {
        m.part_size = 1;
        m.part_rank = 0;
    }
    e.set_powers_partition(m.part_rank, m.part_size);
    setup_p2p(m);
    if (dag_split < 0 && m.part_size > 1 && !m.p2p && products < 128) {
        // no peer memory and a small DAG: recomputing is faster than all-gathering (every rank of the group decides alike:
        // the peer-memory vote is unanimous)
        nccl().CommDestroy(m.part_comm);
        m.part_comm = nullptr;
        m.part_size = 1;
        m.part_rank = 0;
        e.set_powers_partition(0, 1);
    }
    // 4. staging
    const apsu_b200_params &p = e.ctx.params;
    const size_t ct_words = (size_t)2 * e.ctx.first_L * e.ctx.N, idx_words = (size_t)p.query_power_count * ct_words;
    m.q_stage.ensure((m.rank == m.root ? p.bundle_idx_count : std::max<size_t>(m.owned[m.rank].size(), 1)) * idx_words);
    if (m.rank == m.root) m.gathered.ensure((size_t)std::max<uint32_t>(total, 1) * 2 * e.ctx.N);
    while (m.chunk_ev.size() < p.bundle_idx_count + 1) {
        cudaEvent_t ev;
        APSU_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        m.chunk_ev.push_back(ev);
    }
    m.committed = true;
}

// ComputePowers on this rank; with a split PowersDag every DAG level is followed by the in-place all-gather of its
// exchange regions between the ranks of the group
void compute_powers(apsu_b200_mgpu &m)
{
    Engine &e = *m.eng;
    if (m.part_size == 1) {
        e.compute_powers();
        return;
    }
    const uint32_t stages = e.powers_stage_count();
    if (m.p2p) {
        // the levels are exchanged by the kernels themselves (mirrored stores + flag barrier on the stream)
        if (e.arena_base() != m.arena_at_commit) throw std::logic_error("the DB changed since apsu_b200_mgpu_commit: commit again");
        for (uint32_t s = 0; s < stages; s++) e.compute_powers_stage(s);
        return;
    }
    std::vector<void *> ptrs(e.ctx.params.bundle_idx_count);
    std::vector<uint64_t> bytes(e.ctx.params.bundle_idx_count);
    for (uint32_t s = 0; s < stages; s++) {
        e.compute_powers_stage(s);
        if (s + 1 == stages) break;
        const uint32_t n = e.powers_exchange_regions(s + 1, ptrs.data(), bytes.data(), (uint32_t)ptrs.size());
        if (n > 1) APSU_NCCL_CHECK(nccl().GroupStart());
        for (uint32_t k = 0; k < n; k++) {
            char *base = static_cast<char *>(ptrs[k]);
            APSU_NCCL_CHECK(nccl().AllGather(base + (size_t)m.part_rank * bytes[k], base, bytes[k], ncclUint8, m.part_comm, e.ctx.stream));
        }
        if (n > 1) APSU_NCCL_CHECK(nccl().GroupEnd());
    }
}

// shared_query: every rank was handed the query (the same host memory: threads of one process, or processes mapping
// one shared segment) and uploads its own part over its own PCIe link — no scatter, no broadcast
void run_query(apsu_b200_mgpu &m, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *cts, const uint64_t *relin_keys, const uint64_t *masks_local,
               uint32_t npack_local, uint64_t *out, uint32_t *bundle_idx, uint32_t *cache_idx, bool shared_query, bool local_results = false)
{
    if (!m.committed) throw std::logic_error("apsu_b200_mgpu_commit has not been called");
    Engine &e = *m.eng;
    Nccl &nc = nccl();
    const apsu_b200_params &p = e.ctx.params;
    cudaStream_t st = e.ctx.stream;
    const bool is_root = m.rank == m.root;
    const uint32_t bic = p.bundle_idx_count;
    if (nsrc != p.query_power_count) throw std::invalid_argument("query powers do not match the query_powers of the parameters");
    const size_t ct_words = (size_t)2 * e.ctx.first_L * e.ctx.N, ct_bytes = ct_words * 8, idx_words = (size_t)nsrc * ct_words;
    const size_t key_words = e.ctx.using_keyswitching() ? (size_t)(e.ctx.K - 1) * 2 * e.ctx.K * e.ctx.N : 0;
    if (local_results && !shared_query) throw std::invalid_argument("local result delivery needs the shared query");
    if (local_results && m.counts[m.rank] && !out) throw std::invalid_argument("local result delivery: this rank needs an output buffer");
    if (!local_results && is_root && (!cts || (key_words && !relin_keys) || !out)) throw std::invalid_argument("root rank needs the query, the keys and the output buffer");
    if (shared_query && (!cts || (key_words && !relin_keys))) throw std::invalid_argument("shared query: every rank needs the query and the keys");
    e.query_begin_partial(src_powers, nsrc);
    if (shared_query) {
        if (key_words) {
            e.reserve_relin_keys();
            APSU_CUDA_CHECK(cudaMemcpyAsync(e.relin_keys_device(), relin_keys, key_words * 8, cudaMemcpyHostToDevice, st));
            e.relin_keys_loaded();
        }
        uint32_t slot = 0;
        for (uint32_t b : m.owned[m.rank]) { // host [nsrc][bic][ct] -> device [nsrc][ct]
            u64 *buf = m.q_stage.p + (size_t)(is_root ? b : slot++) * idx_words;
            APSU_CUDA_CHECK(cudaMemcpy2DAsync(buf, ct_bytes, cts + (size_t)b * ct_words, (size_t)bic * ct_bytes, ct_bytes, nsrc, cudaMemcpyHostToDevice, st));
            e.query_load_index(b, buf);
        }
    }

    // ---- C1: keys to everyone, ciphertexts of bundle index b to the ranks that own BinBundles of b ----
    if (!shared_query && is_root) {
        // uploads run on the copy stream, chunk by chunk; the sends of chunk b wait for its event only
        APSU_CUDA_CHECK(cudaEventRecord(m.ready_ev, st)); // staging buffers are free once earlier work on st is done
        APSU_CUDA_CHECK(cudaStreamWaitEvent(m.copy_stream, m.ready_ev, 0));
        if (key_words) {
            e.reserve_relin_keys();
            APSU_CUDA_CHECK(cudaMemcpyAsync(e.relin_keys_device(), relin_keys, key_words * 8, cudaMemcpyHostToDevice, m.copy_stream));
            APSU_CUDA_CHECK(cudaEventRecord(m.chunk_ev[bic], m.copy_stream));
        }
        for (uint32_t b = 0; b < bic; b++) {
            bool needed = false;
            for (auto &o : m.owned) needed |= std::find(o.begin(), o.end(), b) != o.end();
            if (!needed) continue;
            // host [nsrc][bic][ct] -> device [b][nsrc][ct]
            APSU_CUDA_CHECK(cudaMemcpy2DAsync(m.q_stage.p + (size_t)b * idx_words, ct_bytes, cts + (size_t)b * ct_words, (size_t)bic * ct_bytes, ct_bytes, nsrc,
                                              cudaMemcpyHostToDevice, m.copy_stream));
            APSU_CUDA_CHECK(cudaEventRecord(m.chunk_ev[b], m.copy_stream));
        }
    }
    if (key_words && !shared_query) {
        if (is_root) APSU_CUDA_CHECK(cudaStreamWaitEvent(st, m.chunk_ev[bic], 0));
        else e.reserve_relin_keys();
        if (m.world > 1) APSU_NCCL_CHECK(nc.Broadcast(e.relin_keys_device(), e.relin_keys_device(), key_words, ncclUint64, (int)m.root, m.comm, st));
        e.relin_keys_loaded();
    }
    if (shared_query) {
        // nothing to exchange
    } else if (is_root) {
        for (uint32_t b = 0; b < bic; b++) {
            std::vector<uint32_t> dst;
            bool mine = false;
            for (uint32_t r = 0; r < m.world; r++) {
                if (std::find(m.owned[r].begin(), m.owned[r].end(), b) == m.owned[r].end()) continue;
                if (r == m.rank) mine = true;
                else dst.push_back(r);
            }
            if (dst.empty() && !mine) continue;
            APSU_CUDA_CHECK(cudaStreamWaitEvent(st, m.chunk_ev[b], 0));
            if (!dst.empty()) {
                APSU_NCCL_CHECK(nc.GroupStart());
                for (uint32_t r : dst) APSU_NCCL_CHECK(nc.Send(m.q_stage.p + (size_t)b * idx_words, idx_words, ncclUint64, (int)r, m.comm, st));
                APSU_NCCL_CHECK(nc.GroupEnd());
            }
            if (mine) e.query_load_index(b, m.q_stage.p + (size_t)b * idx_words);
        }
    } else {
        uint32_t slot = 0;
        for (uint32_t b : m.owned[m.rank]) {
            u64 *buf = m.q_stage.p + (size_t)slot++ * idx_words;
            APSU_NCCL_CHECK(nc.Recv(buf, idx_words, ncclUint64, (int)m.root, m.comm, st));
            e.query_load_index(b, buf);
        }
    }
    // ---- masks are this rank's own (RunQuery draws them, receiver_ddh.cpp:218-289) ----
    if (masks_local) e.set_masks_overlapped(masks_local, npack_local); // behind ComputePowers: read by the last kernel only

    // ---- the evaluation ----
    if (!e.result_order().empty()) {
        compute_powers(m);
        e.eval_all();
    }

    // ---- C3: results to root, unpadded ----
    void *res = nullptr;
    uint64_t res_bytes = 0;
    e.results_device(&res, &res_bytes);
    const size_t per = (size_t)2 * e.ctx.N;
    if (local_results) {
        // every rank hands ITS BinBundles' results to its own host: no gather, the device-to-host copies of all ranks
        // run in parallel (the reference sends every ResultPackage from the worker that finished it, receiver_ddh.cpp:527-534)
        const size_t mine = m.counts[m.rank];
        if (mine) APSU_CUDA_CHECK(cudaMemcpyAsync(out, res, mine * per * 8, cudaMemcpyDeviceToHost, st));
        e.throw_if_query_invalid(); // synchronises
        APSU_CUDA_CHECK(cudaStreamSynchronize(st));
        size_t first = 0;
        for (uint32_t r = 0; r < m.rank; r++) first += m.counts[r];
        for (size_t k = 0; k < mine; k++) {
            if (bundle_idx) bundle_idx[k] = m.all_bundle_idx[first + k];
            if (cache_idx) cache_idx[k] = m.all_cache_idx[first + k];
        }
        return;
    }
    if (!is_root) {
        if (m.counts[m.rank]) APSU_NCCL_CHECK(nc.Send(res, m.counts[m.rank] * per, ncclUint64, (int)m.root, m.comm, st));
        e.throw_if_query_invalid();
        APSU_CUDA_CHECK(cudaStreamSynchronize(st));
        return;
    }
    size_t off = 0, total = 0;
    for (uint32_t c : m.counts) total += c;
    if (m.world > 1) APSU_NCCL_CHECK(nc.GroupStart());
    for (uint32_t r = 0; r < m.world; r++) {
        if (r == m.rank) {
            if (m.counts[r]) APSU_CUDA_CHECK(cudaMemcpyAsync(out + off * per, res, m.counts[r] * per * 8, cudaMemcpyDeviceToHost, st));
        } else if (m.counts[r]) {
            APSU_NCCL_CHECK(nc.Recv(m.gathered.p + off * per, m.counts[r] * per, ncclUint64, (int)r, m.comm, st));
        }
        off += m.counts[r];
    }
    if (m.world > 1) APSU_NCCL_CHECK(nc.GroupEnd());
    off = 0;
    for (uint32_t r = 0; r < m.world; r++) {
        if (r != m.rank && m.counts[r]) APSU_CUDA_CHECK(cudaMemcpyAsync(out + off * per, m.gathered.p + off * per, m.counts[r] * per * 8, cudaMemcpyDeviceToHost, st));
        off += m.counts[r];
    }
    e.throw_if_query_invalid(); // synchronises
    APSU_CUDA_CHECK(cudaStreamSynchronize(st));
    for (size_t k = 0; k < total; k++) {
        if (bundle_idx) bundle_idx[k] = m.all_bundle_idx[k];
        if (cache_idx) cache_idx[k] = m.all_cache_idx[k];
    }
}

} // namespace

extern "C" {

int apsu_b200_mgpu_unique_id(uint8_t *id)
{
    return guarded_call([&] {
        if (!id) throw std::invalid_argument("id is null");
        static_assert(sizeof(ncclUniqueId) == APSU_B200_MGPU_ID_BYTES, "ncclUniqueId size");
        ncclUniqueId u;
        APSU_NCCL_CHECK(nccl().GetUniqueId(&u));
        std::memcpy(id, &u, sizeof(u));
    });
}

int apsu_b200_mgpu_create(apsu_b200_ctx *ctx, const uint8_t *id, uint32_t rank, uint32_t world, apsu_b200_mgpu **out)
{
    return guarded_call([&] {
        if (!id || !out || !world || rank >= world) throw std::invalid_argument("mgpu_create: bad arguments");
        Engine &e = engine_of(ctx);
        auto m = std::make_unique<apsu_b200_mgpu>();
        m->eng = &e;
        m->rank = rank;
        m->world = world;
        ncclUniqueId u;
        std::memcpy(&u, id, sizeof(u));
        try {
            APSU_NCCL_CHECK(nccl().CommInitRank(&m->comm, (int)world, u, (int)rank));
            APSU_CUDA_CHECK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
            APSU_CUDA_CHECK(cudaEventCreateWithFlags(&m->ready_ev, cudaEventDisableTiming));
        } catch (...) {
            destroy(m.release());
            throw;
        }
        *out = m.release();
    });
}

void apsu_b200_mgpu_destroy(apsu_b200_mgpu *m)
{
    if (m && m->eng) {
        cudaSetDevice(m->eng->ctx.device);
        cudaStreamSynchronize(m->eng->ctx.stream);
    }
    destroy(m);
}

int apsu_b200_mgpu_commit(apsu_b200_mgpu *m, const uint32_t *global_cache_idx, int dag_split)
{
    return guarded_call([&] {
        if (!m) throw std::invalid_argument("mgpu is null");
        APSU_CUDA_CHECK(cudaSetDevice(m->eng->ctx.device));
        commit(*m, global_cache_idx, dag_split);
    });
}

int apsu_b200_mgpu_info(const apsu_b200_mgpu *m, uint32_t *total_bin_bundles, uint32_t *dag_group_size, int *dag_exchange, int *nccl_version)
{
    return guarded_call([&] {
        if (!m || !m->committed) throw std::logic_error("apsu_b200_mgpu_commit has not been called");
        uint32_t total = 0;
        for (uint32_t c : m->counts) total += c;
        if (total_bin_bundles) *total_bin_bundles = total;
        if (dag_group_size) *dag_group_size = m->part_size;
        if (dag_exchange) *dag_exchange = m->part_size < 2 ? 0 : (m->p2p ? 2 : 1);
        if (nccl_version) APSU_NCCL_CHECK(nccl().GetVersion(nccl_version));
    });
}

int apsu_b200_mgpu_compute_powers(apsu_b200_mgpu *m)
{
    return guarded_call([&] {
        if (!m || !m->committed) throw std::logic_error("apsu_b200_mgpu_commit has not been called");
        APSU_CUDA_CHECK(cudaSetDevice(m->eng->ctx.device));
        compute_powers(*m);
    });
}

int apsu_b200_mgpu_run_query(
    apsu_b200_mgpu *m, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *cts, const uint64_t *relin_keys, const uint64_t *masks_local,
    uint32_t npack_local, uint64_t *out, uint32_t *bundle_idx, uint32_t *cache_idx)
{
    return guarded_call([&] {
        if (!m) throw std::invalid_argument("mgpu is null");
        APSU_CUDA_CHECK(cudaSetDevice(m->eng->ctx.device));
        run_query(*m, src_powers, nsrc, cts, relin_keys, masks_local, npack_local, out, bundle_idx, cache_idx, false);
    });
}

int apsu_b200_mgpu_run_query_shared(
    apsu_b200_mgpu *m, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *cts, const uint64_t *relin_keys, const uint64_t *masks_local,
    uint32_t npack_local, uint64_t *out, uint32_t *bundle_idx, uint32_t *cache_idx)
{
    return guarded_call([&] {
        if (!m) throw std::invalid_argument("mgpu is null");
        APSU_CUDA_CHECK(cudaSetDevice(m->eng->ctx.device));
        run_query(*m, src_powers, nsrc, cts, relin_keys, masks_local, npack_local, out, bundle_idx, cache_idx, true);
    });
}

int apsu_b200_mgpu_run_query_local(
    apsu_b200_mgpu *m, const uint32_t *src_powers, uint32_t nsrc, const uint64_t *cts, const uint64_t *relin_keys, const uint64_t *masks_local,
    uint32_t npack_local, uint64_t *out_local, uint32_t *bundle_idx_local, uint32_t *cache_idx_local)
{
    return guarded_call([&] {
        if (!m) throw std::invalid_argument("mgpu is null");
        APSU_CUDA_CHECK(cudaSetDevice(m->eng->ctx.device));
        run_query(*m, src_powers, nsrc, cts, relin_keys, masks_local, npack_local, out_local, bundle_idx_local, cache_idx_local, true, true);
    });
}

int apsu_b200_mgpu_local_count(const apsu_b200_mgpu *m, uint32_t *count)
{
    return guarded_call([&] {
        if (!m || !m->committed) throw std::logic_error("apsu_b200_mgpu_commit has not been called");
        if (!count) throw std::invalid_argument("count is null");
        *count = m->counts[m->rank];
    });
}

} // extern "C"
