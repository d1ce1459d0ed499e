// Row f1, first half: ReceiverDB::insert_or_assign on an empty DB (set_data) on the device —
// receiver/apsu/receiver_db.cpp:330-438 (insert_or_assign_worker), :446 (dispatch per bundle index), :966.
//
// The reference walks the algebraised items (`data_with_indices`: felts of one item + its cuckoo index, i.e. the
// first bin of the table slot the item hashed to) IN ORDER, one worker per bundle index, and puts each item into the
// NEWEST BinBundle of that bundle index whose bins for the item would stay below max_items_per_bin (bundles are scanned
// from the back, :370-393); when none has room a new BinBundle is appended (:407-432).  Which BinBundle an item lands in
// therefore depends on the arrival order of everything before it — a sequential process over 50 M insertions at 2^24.
//
// Restated so that it parallelises, per bundle index:
//   * all felts_per_item bins of a table slot always grow together, so the state is one counter per (BinBundle, slot);
//   * between two BinBundle creations the set of BinBundles is fixed and a slot's arrivals simply fill them newest
//     first: with cap = max_items_per_bin - 1 items per bin, the arrivals of a slot in that window go to the newest
//     BinBundle until it holds cap, then to the next older one with room, ...: pure arithmetic on the counters;
//   * a creation happens at the first arrival (in global order) that finds all BinBundles full for its slot: for
//     every slot that is its (free capacity + 1)-th arrival of the window, and the creation time is the minimum over
//     the slots.  That arrival opens the new BinBundle; the next window starts right after it.
// So the build is one stable sort of the items by slot (arrival order kept inside a slot), then one tiny kernel per
// BinBundle creation (a thread per slot: binary search + counter arithmetic, an atomicMin for the next creation time),
// then one pass that replays the windows per slot and writes every item's (BinBundle, position in its bins), then one
// scatter of the felts into per-BinBundle bin lists, and the cache build of every BinBundle (engine.cu:
// add_binbundle_from_bins_device) — all on the context stream; the host only reads back one word per creation.
// The result is the reference's DB bit for bit (tests/test_gpu_dbbuild.py against the sequential restatement).
#include "engine.hpp"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>

namespace apsu_b200 {

namespace {

constexpr unsigned long long kNoCreation = ~0ull;

// location (table slot) of every item + validation: the cuckoo index is the first bin of a slot
__global__ void k_ff_keys(const u64 *__restrict__ cuckoo_idx, size_t n, u32 felts_per_item, u32 table_size, u32 *__restrict__ keys, u32 *__restrict__ vals,
                          int *__restrict__ bad)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 c = cuckoo_idx[i];
    u64 loc = c / felts_per_item;
    if (c % felts_per_item || loc >= table_size) {
        atomicExch(bad, 1);
        loc = 0;
    }
    keys[i] = (u32)loc;
    vals[i] = (u32)i;
}
// slot_first[k] = number of items in slots below k = lower bound of k in the sorted keys (k = 0 .. table_size)
__global__ void k_ff_slot_first(const u32 *__restrict__ sorted_keys, u32 n, u32 table_size, u32 *__restrict__ slot_first)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > table_size) return;
    u32 lo = 0, hi = n;
    while (lo < hi) {
        const u32 mid = lo + ((hi - lo) >> 1);
        if (sorted_keys[mid] < k) lo = mid + 1;
        else hi = mid;
    }
    slot_first[k] = lo;
}

struct FfState {
    const u32 *arrivals;    // item indices sorted by slot, arrival order inside a slot
    const u32 *slot_first;  // [table_size + 1] exclusive prefix of the slot populations
    u32 *start;             // [ipb] arrivals of the slot already placed
    u32 *counts;            // [max_bundles][ipb] items per bin of (BinBundle, slot)
    unsigned long long *creation; // [max_bundles + 1] packed (arrival index << 16 | slot) of every creation; [e] = next one
    u32 slot0, ipb, cap, max_bundles;
};

// fills newest-first: `n` arrivals into BinBundles B-1 .. 0 with room; returns what did not fit (0 by construction)
__device__ __forceinline__ u32 ff_distribute(u32 *counts, u32 ipb, u32 s, u32 B, u32 cap, u32 n)
{
    for (int c = (int)B - 1; c >= 0 && n; c--) {
        const u32 have = counts[(size_t)c * ipb + s], take = min(cap - have, n);
        counts[(size_t)c * ipb + s] = have + take;
        n -= take;
    }
    return n;
}

// One window.  B = BinBundles that exist during the window; prev = the creation that opened BinBundle B-1 (none for
// the first window).  Thread s: place the arrivals of the window that just ended, then find this slot's next overflow.
__global__ void k_ff_window(FfState st, u32 B)
{
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= st.ipb) return;
    const u32 *A = st.arrivals + st.slot_first[st.slot0 + s];
    const u32 len = st.slot_first[st.slot0 + s + 1] - st.slot_first[st.slot0 + s];
    u32 start = st.start[s];
    if (B >= 1) { // (not "creation[0] is set": other threads are writing creation[B] of THIS window right now)
        // window that ended with creation[B-1]: arrivals before it were placed among BinBundles 0..B-2
        const unsigned long long prev = st.creation[B - 1];
        const u32 g = (u32)(prev >> 16), trig = (u32)(prev & 0xFFFF);
        u32 lo = start, hi = len; // first arrival >= g
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (A[mid] < g) lo = mid + 1;
            else hi = mid;
        }
        ff_distribute(st.counts, st.ipb, s, B - 1, st.cap, lo - start);
        start = lo;
        if (s == trig) { // the arrival that found everything full opens BinBundle B-1
            st.counts[(size_t)(B - 1) * st.ipb + s] = 1;
            start++;
        }
        st.start[s] = start;
    }
    u32 free_total = 0;
    for (u32 c = 0; c < B; c++) free_total += st.cap - st.counts[(size_t)c * st.ipb + s];
    const unsigned long long fail_pos = (unsigned long long)start + free_total;
    if (fail_pos < len) atomicMin(&st.creation[B], ((unsigned long long)A[fail_pos] << 16) | s);
}

// Replays every window for one slot (one warp per slot) and writes, for every arrival, the BinBundle it lands in and
// its position inside the bins of that BinBundle.  n_created = number of creations (BinBundles = n_created, the first
// arrival of the bundle index being creation 0 of an empty DB).
__global__ void k_ff_assign(FfState st, u32 n_bundles, u32 *__restrict__ item_bundle, u32 *__restrict__ item_pos)
{
    const u32 s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= st.ipb) return;
    const size_t base = st.slot_first[st.slot0 + s];
    const u32 *A = st.arrivals + base;
    const u32 len = st.slot_first[st.slot0 + s + 1] - st.slot_first[st.slot0 + s];
    u32 start = 0;
    // window w runs with BinBundles 0..w-1 ... the first window (w = 0) has none: its only event is creation 0
    for (u32 w = 0; w <= n_bundles; w++) {
        // arrivals of this window: [start, end) with end = first arrival >= creation[w] (all remaining in the last one)
        u32 end = len;
        u32 trig = 0xFFFFFFFFu;
        if (w < n_bundles) {
            const unsigned long long cr = st.creation[w];
            const u32 g = (u32)(cr >> 16);
            trig = (u32)(cr & 0xFFFF);
            u32 lo = start, hi = len;
            while (lo < hi) {
                const u32 mid = (lo + hi) >> 1;
                if (A[mid] < g) lo = mid + 1;
                else hi = mid;
            }
            end = lo;
        }
        // newest first over BinBundles w-1 .. 0; counts = st.counts rebuilt on the fly in registers is not possible for
        // many bundles, so the per-(bundle, slot) totals are re-accumulated in st.counts (zeroed before this kernel)
        u32 at = start;
        for (int c = (int)w - 1; c >= 0 && at < end; c--) {
            const u32 have = st.counts[(size_t)c * st.ipb + s], take = min(st.cap - have, end - at);
            for (u32 k = lane; k < take; k += 32) {
                item_bundle[base + at + k] = (u32)c;
                item_pos[base + at + k] = have + k;
            }
            __syncwarp();
            if (lane == 0) st.counts[(size_t)c * st.ipb + s] = have + take;
            __syncwarp();
            at += take;
        }
        start = end;
        if (s == trig) {
            if (lane == 0) {
                item_bundle[base + start] = w;
                item_pos[base + start] = 0;
                st.counts[(size_t)w * st.ipb + s] = 1;
            }
            __syncwarp();
            start++;
        }
    }
}

// roots of every bin of every BinBundle of one bundle index: bundle c's bins are [bundle_base[c] + bin_first[c][bin] ..),
// bin = slot * F + f holds felt f of the items of that slot
__global__ void k_ff_scatter(const u64 *__restrict__ felts, const u32 *__restrict__ arrivals, const u32 *__restrict__ item_bundle, const u32 *__restrict__ item_pos,
                             const u32 *__restrict__ slot_first, u32 slot0, u32 ipb, u32 F, const u32 *__restrict__ bin_first /*[n_bundles][ipb*F]*/,
                             const u64 *__restrict__ bundle_base, u64 *__restrict__ roots)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; // position among the sorted items of this bundle index
    const size_t lo = slot_first[slot0], hi = slot_first[slot0 + ipb];
    if (lo + k >= hi) return;
    const size_t at = lo + k;
    // slot of this position: binary search in slot_first
    u32 a = slot0, b = slot0 + ipb;
    while (b - a > 1) {
        const u32 mid = (a + b) >> 1;
        if (slot_first[mid] <= at) a = mid;
        else b = mid;
    }
    const u32 s = a - slot0, c = item_bundle[at], pos = item_pos[at];
    const u64 *x = felts + (size_t)arrivals[at] * F;
    for (u32 f = 0; f < F; f++) roots[bundle_base[c] + bin_first[(size_t)c * ipb * F + (size_t)s * F + f] + pos] = x[f];
}

} // namespace

void Engine::set_data(const uint64_t *felts, const uint64_t *cuckoo_idx, size_t n, bool on_device, uint32_t *bundle_counts)
{
    const apsu_b200_params &p = ctx.params;
    const uint32_t F = p.felts_per_item, ipb = p.items_per_bundle, bic = p.bundle_idx_count, table = p.table_size;
    const uint32_t cap = p.max_items_per_bin - 1;
    if (!felts || !cuckoo_idx) throw std::invalid_argument("set_data: items are null");
    if (n >= (1ull << 31)) throw std::invalid_argument("set_data: too many items (item indices are 32-bit, the sort takes an int count)");
    if (p.max_items_per_bin < 2) throw std::invalid_argument("max_items_per_bin must be at least 2 to hold an item");
    if (ipb > 0xFFFF) throw std::invalid_argument("set_data: too many slots per bundle");
    // APSU_B200_BUILD_TIMING=1: host-timed phases (with a stream synchronisation each) on stderr
    const bool timing = std::getenv("APSU_B200_BUILD_TIMING") && atoi(std::getenv("APSU_B200_BUILD_TIMING"));
    auto t_last = std::chrono::steady_clock::now();
    auto phase = [&](const char *what) {
        if (!timing) return;
        cudaStreamSynchronize(ctx.stream);
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[set_data] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    clear_db();
    phase("clear_db (frees)");
    cudaStream_t st = ctx.stream;
    if (!n) {
        if (bundle_counts) std::fill(bundle_counts, bundle_counts + bic, 0u);
        return;
    }
    // ---- items on the device ----
    DBuf<u64> d_felts_own, d_cidx_own;
    const u64 *d_felts = (const u64 *)felts, *d_cidx = (const u64 *)cuckoo_idx;
    if (!on_device) {
        d_felts_own.alloc(n * F);
        d_cidx_own.alloc(n);
        APSU_CUDA_CHECK(cudaMemcpyAsync(d_felts_own.p, felts, n * F * 8, cudaMemcpyHostToDevice, st));
        APSU_CUDA_CHECK(cudaMemcpyAsync(d_cidx_own.p, cuckoo_idx, n * 8, cudaMemcpyHostToDevice, st));
        d_felts = d_felts_own.p;
        d_cidx = d_cidx_own.p;
    }
    // ---- stable sort by slot ----
    DBuf<uint32_t> keys, vals, keys2, arrivals, slot_first;
    DBuf<int> bad;
    keys.alloc(n), vals.alloc(n), keys2.alloc(n), arrivals.alloc(n), slot_first.alloc(table + 1), bad.alloc(1);
    APSU_CUDA_CHECK(cudaMemsetAsync(bad.p, 0, 4, st));
    k_ff_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_cidx, n, F, table, keys.p, vals.p, bad.p);
    APSU_CUDA_CHECK(cudaGetLastError());
    int bits = 1;
    while ((1u << bits) < table) bits++;
    size_t tmp_bytes = 0;
    APSU_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys2.p, vals.p, arrivals.p, (int)n, 0, bits, st));
    DBuf<unsigned char> tmp;
    tmp.alloc(tmp_bytes);
    APSU_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys2.p, vals.p, arrivals.p, (int)n, 0, bits, st));
    k_ff_slot_first<<<(table + 1 + 255) / 256, 256, 0, st>>>(keys2.p, (u32)n, table, slot_first.p);
    APSU_CUDA_CHECK(cudaGetLastError());
    std::vector<uint32_t> h_count(table + 1, 0), h_first(table + 1);
    int h_bad = 0;
    APSU_CUDA_CHECK(cudaMemcpyAsync(h_first.data(), slot_first.p, (table + 1) * 4, cudaMemcpyDeviceToHost, st));
    APSU_CUDA_CHECK(cudaMemcpyAsync(&h_bad, bad.p, 4, cudaMemcpyDeviceToHost, st));
    APSU_CUDA_CHECK(cudaStreamSynchronize(st));
    if (h_bad) throw std::invalid_argument("set_data: a cuckoo index is not the first bin of a table slot");
    for (uint32_t k = 0; k < table; k++) h_count[k] = h_first[k + 1] - h_first[k];
    keys.release(), vals.release(), keys2.release(), tmp.release();
    phase("upload + sort by slot");

    DBuf<uint32_t> item_bundle, item_pos;
    item_bundle.alloc(n), item_pos.alloc(n);

    for (uint32_t b = 0; b < bic; b++) {
        const uint32_t slot0 = b * ipb;
        const size_t n_b = (size_t)h_first[slot0 + ipb] - h_first[slot0];
        if (bundle_counts) bundle_counts[b] = 0;
        if (!n_b) continue;
        // an upper bound on the BinBundles of this bundle index: every creation is preceded by a slot holding a
        // multiple of cap items, and first-fit never leaves more than the fullest slot needs plus the windows' slack
        uint32_t max_slot = 0;
        for (uint32_t s = 0; s < ipb; s++) max_slot = std::max(max_slot, h_count[slot0 + s]);
        uint32_t max_bundles = (max_slot + cap - 1) / cap + 1;
        DBuf<uint32_t> start, counts;
        DBuf<unsigned long long> creation;
        std::vector<unsigned long long> h_creation;
        uint32_t B = 0;
        for (;;) {
            // (re)start with a larger table if the bound was too small (adversarial arrival orders)
            start.alloc(ipb), counts.alloc((size_t)max_bundles * ipb), creation.alloc(max_bundles + 1);
            APSU_CUDA_CHECK(cudaMemsetAsync(start.p, 0, ipb * 4, st));
            APSU_CUDA_CHECK(cudaMemsetAsync(counts.p, 0, (size_t)max_bundles * ipb * 4, st));
            APSU_CUDA_CHECK(cudaMemsetAsync(creation.p, 0xFF, (max_bundles + 1) * 8, st));
            FfState fs{ arrivals.p, slot_first.p, start.p, counts.p, creation.p, slot0, ipb, cap, max_bundles };
            h_creation.clear();
            B = 0;
            bool overflow = false;
            for (;;) {
                // window with B BinBundles: places the previous window, finds creation[B]
                k_ff_window<<<(ipb + 127) / 128, 128, 0, st>>>(fs, B);
                APSU_CUDA_CHECK(cudaGetLastError());
                unsigned long long next = 0;
                APSU_CUDA_CHECK(cudaMemcpyAsync(&next, creation.p + B, 8, cudaMemcpyDeviceToHost, st));
                APSU_CUDA_CHECK(cudaStreamSynchronize(st));
                if (next == kNoCreation) break;
                h_creation.push_back(next);
                B++;
                if (B >= max_bundles) {
                    overflow = true;
                    break;
                }
            }
            if (!overflow) break;
            max_bundles *= 2;
        }
        phase("first-fit windows");
        // ---- every item's BinBundle and position; counts are rebuilt by the replay ----
        APSU_CUDA_CHECK(cudaMemsetAsync(counts.p, 0, (size_t)max_bundles * ipb * 4, st));
        FfState fs{ arrivals.p, slot_first.p, start.p, counts.p, creation.p, slot0, ipb, cap, max_bundles };
        k_ff_assign<<<(ipb * 32 + 255) / 256, 256, 0, st>>>(fs, B, item_bundle.p, item_pos.p);
        APSU_CUDA_CHECK(cudaGetLastError());
        std::vector<uint32_t> h_counts((size_t)B * ipb);
        APSU_CUDA_CHECK(cudaMemcpyAsync(h_counts.data(), counts.p, h_counts.size() * 4, cudaMemcpyDeviceToHost, st));
        APSU_CUDA_CHECK(cudaStreamSynchronize(st));
        // ---- bin tables of the B BinBundles ----
        const uint32_t nbins = ipb * F; // == bins_per_bundle
        std::vector<uint32_t> h_bin_first((size_t)B * nbins), h_bin_size((size_t)B * nbins), max_deg(B, 0);
        std::vector<u64> h_base(B + 1, 0);
        for (uint32_t c = 0; c < B; c++) {
            uint32_t acc = 0;
            for (uint32_t s = 0; s < ipb; s++)
                for (uint32_t f = 0; f < F; f++) {
                    const uint32_t sz = h_counts[(size_t)c * ipb + s];
                    h_bin_first[(size_t)c * nbins + s * F + f] = acc;
                    h_bin_size[(size_t)c * nbins + s * F + f] = sz;
                    acc += sz;
                    max_deg[c] = std::max(max_deg[c], sz);
                }
            h_base[c + 1] = h_base[c] + acc;
        }
        DBuf<uint32_t> bin_first, bin_size;
        DBuf<u64> bundle_base, roots;
        bin_first.upload(h_bin_first, st), bin_size.upload(h_bin_size, st), bundle_base.upload(h_base, st);
        roots.alloc(std::max<u64>(h_base[B], 1));
        k_ff_scatter<<<(unsigned)((n_b + 255) / 256), 256, 0, st>>>(d_felts, arrivals.p, item_bundle.p, item_pos.p, slot_first.p, slot0, ipb, F, bin_first.p,
                                                                   bundle_base.p, roots.p);
        APSU_CUDA_CHECK(cudaGetLastError());
        phase("assign + scatter");
        // ---- BinBundle::regen_cache for each of them (engine.cu), asynchronous ----
        for (uint32_t c = 0; c < B; c++)
            add_binbundle_from_bins_device(b, bin_first.p + (size_t)c * nbins, bin_size.p + (size_t)c * nbins, roots.p + h_base[c], max_deg[c]);
        if (bundle_counts) bundle_counts[b] = B;
        APSU_CUDA_CHECK(cudaStreamSynchronize(st)); // the tables above go out of scope
        phase("BinBundle caches");
    }
    throw_if_build_invalid();
}

} // namespace apsu_b200
