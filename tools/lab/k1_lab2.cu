// Lab 2 (not product): MAC-form / software-pipelining / layout variants of the DB-stream kernel, plus
// IMAD.WIDE throughput as a function of resident warps and independent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I apsu_b200/csrc -o tools/lab/k1_lab2 tools/lab/k1_lab2.cu
#include "modarith.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace apsu_b200;

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e = (x);                                                                       \
        if (e != cudaSuccess) {                                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);         \
            exit(1);                                                                               \
        }                                                                                          \
    } while (0)

// ---- IMAD.WIDE rate vs warps / chains ----
template <int NACC, int FORM>
__global__ void k_imad2(u64 *out, int iters, u32 seed)
{
    u32 a = seed + threadIdx.x, b = a ^ 0x55555;
    u64 acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; k++) acc[k] = k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 32 / NACC; r++) {
#pragma unroll
            for (int k = 0; k < NACC; k++) {
                if (FORM == 0)
                    acc[k] += (u64)(a + k) * b;
                else
                    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b));
            }
            b += 0x9E37u;
        }
    }
    u64 x = 0;
#pragma unroll
    for (int k = 0; k < NACC; k++) x ^= acc[k];
    if (x == 0x12345) out[0] = x;
}

// ---- stream kernel ----
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "LAB_DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ u64 pol_evict_first()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 pol_evict_last()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar, u64 policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
                 : "memory");
}

// Both W and P are stored split at bit `split` (low part in the low half of the u64, rest in the high half).
// lanes: FORM 0/1: ll, mid, hh (4 products); FORM 2/3: ll, kk, hh (Karatsuba, 3 products)
struct Acc3 {
    u64 ll, mid, hh;
};
template <int FORM>
__device__ __forceinline__ void mac(Acc3 &a, u32 wl, u32 wh, u32 ws, u32 pl, u32 ph, u32 ps)
{
    if (FORM == 0) {
        a.ll += (u64)wl * pl;
        a.mid += (u64)wl * ph;
        a.mid += (u64)wh * pl;
        a.hh += (u64)wh * ph;
    } else if (FORM == 1) {
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.ll) : "r"(wl), "r"(pl));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(wl), "r"(ph));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(wh), "r"(pl));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.hh) : "r"(wh), "r"(ph));
    } else if (FORM == 2) {
        // carry-chain form: ptxas folds each pair into one IMAD.WIDE.U32 with the 64-bit addend
        u32 lo, hi;
#define APSU_MADW(acc, x, y)                                                                                   \
    lo = (u32)(acc);                                                                                           \
    hi = (u32)((acc) >> 32);                                                                                   \
    asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(x), "r"(y)); \
    (acc) = ((u64)hi << 32) | lo;
        APSU_MADW(a.ll, wl, pl)
        APSU_MADW(a.mid, ws, ps)
        APSU_MADW(a.hh, wh, ph)
    } else if (FORM == 4) {
        a.ll += (u64)wl * pl;
        a.mid += (u64)ws * ps;
        a.hh += (u64)wh * ph;
    } else {
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.ll) : "r"(wl), "r"(pl));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(ws), "r"(ps));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.hh) : "r"(wh), "r"(ph));
    }
}
// value of the lanes modulo q; no renormalisation inside the lab (T is small enough)
template <int FORM>
__device__ __forceinline__ u64 reduce(const Acc3 &a, const DMod &m, int split)
{
    u64 mid = FORM >= 2 ? a.mid - a.ll - a.hh : a.mid;
    // ll + mid*2^s + hh*2^2s as 128 bits
    u64 lo = a.ll, hi = 0;
    u64 t = mid << split;
    lo += t;
    hi += (mid >> (64 - split)) + (lo < t);
    t = a.hh << (2 * split);
    lo += t;
    hi += (a.hh >> (64 - 2 * split)) + (lo < t);
    return barrett128(lo, hi, m);
}

struct LabArgs {
    const u64 *W;
    const u64 *P;
    u64 *out; // [job][2][LN]
    u32 njobs, T, L, N;
    DMod q[4];
    int split;
};

template <int G, int TS>
struct Regs {
    u64 p[TS][2], w[G][TS];
};

template <int G, int TS>
__device__ __forceinline__ void load_stage(Regs<G, TS> &r, const u64 *sb)
{
    constexpr int COLS = 128;
#pragma unroll
    for (int h = 0; h < TS; h++) {
        r.p[h][0] = sb[(h * 2) * COLS];
        r.p[h][1] = sb[(h * 2 + 1) * COLS];
#pragma unroll
        for (int k = 0; k < G; k++) r.w[k][h] = sb[(TS * 2 + k * TS + h) * COLS];
    }
}
// terms [h0, h0+SUB) of a TS-term stage
template <int G, int TS, int SUB>
__device__ __forceinline__ void load_sub(Regs<G, SUB> &r, const u64 *sb, int h0)
{
    constexpr int COLS = 128;
#pragma unroll
    for (int h = 0; h < SUB; h++) {
        r.p[h][0] = sb[((h0 + h) * 2) * COLS];
        r.p[h][1] = sb[((h0 + h) * 2 + 1) * COLS];
#pragma unroll
        for (int k = 0; k < G; k++) r.w[k][h] = sb[(TS * 2 + k * TS + h0 + h) * COLS];
    }
}
template <int G, int TS, int FORM>
__device__ __forceinline__ void mac_stage(const Regs<G, TS> &r, Acc3 (&acc)[G][2])
{
#pragma unroll
    for (int h = 0; h < TS; h++) {
        const u32 p0l = (u32)r.p[h][0], p0h = (u32)(r.p[h][0] >> 32), p1l = (u32)r.p[h][1], p1h = (u32)(r.p[h][1] >> 32);
        const u32 p0s = p0l + p0h, p1s = p1l + p1h;
#pragma unroll
        for (int k = 0; k < G; k++) {
            const u32 wl = (u32)r.w[k][h], wh = (u32)(r.w[k][h] >> 32), ws = wl + wh;
            mac<FORM>(acc[k][0], wl, wh, ws, p0l, p0h, p0s);
            mac<FORM>(acc[k][1], wl, wh, ws, p1l, p1h, p1s);
        }
    }
}

// PIPE 0: wait / load / release / MAC per stage.  PIPE 1: the registers of stage i+1 are loaded before the MACs of
// stage i (two register sets, loop unrolled by two stages).
template <int G, int TS, int STAGES, int FORM, int LAYOUT, int CPS, int PIPE, int UN, int SUB>
__global__ void __launch_bounds__(160, CPS) k_stream(LabArgs a)
{
    constexpr int COLS = 128;
    constexpr int term_words = (2 + G) * COLS;
    constexpr int stage_words = TS * term_words;
    extern __shared__ __align__(128) u64 smem[];
    u64 *ring = smem;
    u64 *full = smem + (size_t)STAGES * stage_words;
    u64 *empty = full + STAGES;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const u32 LN = a.L * a.N, ntiles = LN / COLS, slices = a.njobs / G, T = a.T;
    const u32 n_items = slices * ntiles;
    const u32 nst = T / TS; // lab: T is a multiple of 2*TS
    u32 it = 0;
    if (tid >= COLS) {
        if (tid != COLS) return;
        const u64 pol_s = pol_evict_first(), pol_k = pol_evict_last();
        for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
            const u32 tile = item % ntiles, job0 = (item / ntiles) * G;
            for (u32 st = 0; st < nst; st++, it++) {
                const int s = it % STAGES;
                const u32 use = it / STAGES;
                if (use) mbar_wait(&empty[s], (use - 1) & 1);
                const u32 t0 = st * TS;
                mbar_expect_tx(&full[s], stage_words * 8);
                u64 *sb = ring + (size_t)s * stage_words;
                if (LAYOUT == 0) {
#pragma unroll
                    for (u32 h = 0; h < TS; h++) {
                        bulk_g2s(sb + (h * 2) * COLS, a.P + ((size_t)(t0 + h) * 2) * LN + tile * COLS, COLS * 8, &full[s], pol_k);
                        bulk_g2s(sb + (h * 2 + 1) * COLS, a.P + ((size_t)(t0 + h) * 2 + 1) * LN + tile * COLS, COLS * 8, &full[s], pol_k);
                    }
#pragma unroll
                    for (int k = 0; k < G; k++)
#pragma unroll
                        for (u32 h = 0; h < TS; h++)
                            bulk_g2s(sb + (TS * 2 + k * TS + h) * COLS, a.W + ((size_t)(job0 + k) * T + t0 + h) * LN + tile * COLS, COLS * 8, &full[s], pol_s);
                } else {
                    bulk_g2s(sb, a.P + ((size_t)tile * T + t0) * 2 * COLS, TS * 2 * COLS * 8, &full[s], pol_k);
#pragma unroll
                    for (int k = 0; k < G; k++)
                        bulk_g2s(sb + (TS * 2 + k * TS) * COLS, a.W + (((size_t)(job0 + k) * ntiles + tile) * T + t0) * COLS, TS * COLS * 8, &full[s], pol_s);
                }
            }
        }
        return;
    }
    const bool lane0 = (tid & 31) == 0;
    for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
        const u32 tile = item % ntiles, job0 = (item / ntiles) * G;
        const u32 col0 = tile * COLS;
        const DMod m = a.q[col0 / a.N];
        Acc3 acc[G][2];
#pragma unroll
        for (int k = 0; k < G; k++) acc[k][0] = acc[k][1] = Acc3{ 0, 0, 0 };
        if (PIPE == 0) {
#pragma unroll UN
            for (u32 st = 0; st < nst; st++, it++) {
                const int s = it % STAGES;
                mbar_wait(&full[s], (it / STAGES) & 1);
                if (SUB == TS) {
                    Regs<G, TS> r;
                    load_stage<G, TS>(r, ring + (size_t)s * stage_words + tid);
                    __syncwarp();
                    if (lane0) mbar_arrive(&empty[s]);
                    mac_stage<G, TS, FORM>(r, acc);
                } else {
#pragma unroll
                    for (int h0 = 0; h0 < TS; h0 += SUB) {
                        Regs<G, SUB> r;
                        load_sub<G, TS, SUB>(r, ring + (size_t)s * stage_words + tid, h0);
                        if (h0 + SUB == TS) {
                            __syncwarp();
                            if (lane0) mbar_arrive(&empty[s]);
                        }
                        mac_stage<G, SUB, FORM>(r, acc);
                    }
                }
            }
        } else {
            Regs<G, TS> ra, rb;
            {
                const int s = it % STAGES;
                mbar_wait(&full[s], (it / STAGES) & 1);
                load_stage<G, TS>(ra, ring + (size_t)s * stage_words + tid);
                __syncwarp();
                if (lane0) mbar_arrive(&empty[s]);
                it++;
            }
            for (u32 st = 0; st < nst; st += 2) {
                {
                    const int s = it % STAGES;
                    mbar_wait(&full[s], (it / STAGES) & 1);
                    load_stage<G, TS>(rb, ring + (size_t)s * stage_words + tid);
                    __syncwarp();
                    if (lane0) mbar_arrive(&empty[s]);
                    it++;
                }
                mac_stage<G, TS, FORM>(ra, acc);
                if (st + 2 < nst) {
                    const int s = it % STAGES;
                    mbar_wait(&full[s], (it / STAGES) & 1);
                    load_stage<G, TS>(ra, ring + (size_t)s * stage_words + tid);
                    __syncwarp();
                    if (lane0) mbar_arrive(&empty[s]);
                    it++;
                }
                mac_stage<G, TS, FORM>(rb, acc);
            }
        }
#pragma unroll
        for (int k = 0; k < G; k++) {
            u64 *o = a.out + ((size_t)(job0 + k) * 2) * LN + col0 + tid;
            o[0] = reduce<FORM>(acc[k][0], m, a.split);
            o[LN] = reduce<FORM>(acc[k][1], m, a.split);
        }
    }
}

// ---- data ----
__host__ __device__ inline u64 splitmix(u64 x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline u64 val_w(u32 job, u32 term, u32 col, u64 q) { return splitmix(((u64)job << 40) ^ ((u64)term << 24) ^ col ^ 0xABCDull << 50) % q; }
__host__ __device__ inline u64 val_p(u32 term, u32 comp, u32 col, u64 q) { return splitmix(((u64)term << 24) ^ ((u64)comp << 60) ^ col ^ 0x77ull << 52) % q; }
__host__ __device__ inline u64 packs(u64 w, int s) { return (w & ((1ull << s) - 1)) | ((w >> s) << 32); }

template <int LAYOUT>
__global__ void k_fill_w(u64 *W, u32 njobs, u32 T, u32 L, u32 N, u64 q0, u64 q1, u64 q2, int split)
{
    const u32 LN = L * N, ntiles = LN / 128;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)njobs * T * LN;
    if (i >= total) return;
    u32 col = i % LN, term = (i / LN) % T, job = i / ((size_t)LN * T);
    u32 l = col / N;
    u64 q = l == 0 ? q0 : l == 1 ? q1 : q2;
    size_t dst = LAYOUT == 0 ? i : ((((size_t)job * ntiles + col / 128) * T + term) * 128 + col % 128);
    W[dst] = packs(val_w(job, term, col, q), split);
}
template <int LAYOUT>
__global__ void k_fill_p(u64 *P, u32 T, u32 L, u32 N, u64 q0, u64 q1, u64 q2, int split)
{
    const u32 LN = L * N;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)T * 2 * LN;
    if (i >= total) return;
    u32 col = i % LN, comp = (i / LN) % 2, term = i / ((size_t)LN * 2);
    u32 l = col / N;
    u64 q = l == 0 ? q0 : l == 1 ? q1 : q2;
    size_t dst = LAYOUT == 0 ? i : ((((size_t)(col / 128) * T + term) * 2 + comp) * 128 + col % 128);
    P[dst] = packs(val_p(term, comp, col, q), split);
}

static DMod make_mod(u64 q)
{
    DMod m;
    m.q = q;
    unsigned __int128 r = (~(unsigned __int128)0) / q;
    m.r0 = (u64)r;
    m.r1 = (u64)(r >> 64);
    return m;
}

struct Timer {
    cudaEvent_t a, b;
    Timer()
    {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    template <typename F>
    float run(F f, int warm = 2, int reps = 5)
    {
        for (int i = 0; i < warm; i++) f();
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int i = 0; i < reps; i++) {
            cudaEventRecord(a);
            f();
            cudaEventRecord(b);
            CK(cudaEventSynchronize(b));
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            best = ms < best ? ms : best;
        }
        CK(cudaGetLastError());
        return best;
    }
};

static u32 g_njobs, g_T, g_L, g_N;
static u64 g_q[4];
static u64 *g_out;
static int g_sms;

static bool check(const char *name)
{
    const u32 LN = g_L * g_N;
    std::vector<u64> h((size_t)g_njobs * 2 * LN);
    CK(cudaMemcpy(h.data(), g_out, h.size() * 8, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (u32 job : { 0u, 1u, g_njobs / 2 + 1, g_njobs - 1 })
        for (u32 col : { 0u, 1u, 127u, 128u, g_N - 1, g_N, LN / 2 + 77, LN - 1 })
            for (u32 comp = 0; comp < 2; comp++) {
                u64 q = g_q[col / g_N];
                unsigned __int128 s = 0;
                for (u32 t = 0; t < g_T; t++) s = (s + (unsigned __int128)val_w(job, t, col, q) * val_p(t, comp, col, q)) % q;
                u64 got = h[((size_t)job * 2 + comp) * LN + col];
                if (got != (u64)s) {
                    if (bad < 2) printf("   MISMATCH %s job %u col %u comp %u: got %llx want %llx\n", name, job, col, comp, got, (u64)s);
                    bad++;
                }
            }
    return bad == 0;
}

struct Data {
    u64 *W[2], *P[2]; // by layout
    int split;
};

template <int G, int TS, int STAGES, int FORM, int LAYOUT, int CPS, int PIPE, int UN = 1, int SUB = TS>
static void run_stream(Timer &tm, const char *name, LabArgs a, const Data &d30, const Data &d28)
{
    constexpr size_t smem = (size_t)STAGES * TS * (2 + G) * 128 * 8 + 2 * STAGES * 8 + 16;
    auto kern = k_stream<G, TS, STAGES, FORM, LAYOUT, CPS, PIPE, UN, SUB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 160, smem));
    if (per_sm > CPS) per_sm = CPS;
    const Data &d = FORM >= 2 ? d28 : d30;
    a.W = d.W[LAYOUT];
    a.P = d.P[LAYOUT];
    a.split = d.split;
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    CK(cudaMemset(g_out, 0, (size_t)g_njobs * 2 * g_L * g_N * 8));
    float ms = tm.run([&] { kern<<<g_sms * per_sm, 160, smem>>>(a); });
    double bytes = (double)g_njobs * g_T * g_L * g_N * 8;
    bool ok = check(name);
    printf("stream %-36s G=%d TS=%d ST=%d form=%d layout=%d pipe=%d un=%d sub=%d cta/sm=%d regs=%d smem=%zuK : %7.3f ms  %7.1f GB/s  %s\n", name, G, TS, STAGES, FORM, LAYOUT, PIPE, UN, SUB,
           per_sm, fa.numRegs, smem / 1024, ms, bytes / ms / 1e6, ok ? "OK" : "WRONG");
    fflush(stdout);
}

template <int NACC, int FORM>
static void run_imad(Timer &tm, int warps_per_sm)
{
    u64 *o;
    CK(cudaMalloc(&o, 64));
    const int iters = 2048;
    // one block per SM with warps_per_sm warps
    float ms = tm.run([&] { k_imad2<NACC, FORM><<<g_sms, warps_per_sm * 32>>>(o, iters, 1); });
    double macs = (double)g_sms * warps_per_sm * 32 * iters * 32;
    printf("imad2 form %d chains/thread %2d warps/SM %2d : %.3f ms  %.2f MAC/clk/SM @1965MHz\n", FORM, NACC, warps_per_sm, ms, macs / ms / 1e6 * 1e9 / g_sms / 1.965e9);
    cudaFree(o);
}

int main(int argc, char **argv)
{
    double gib = argc > 1 ? atof(argv[1]) : 4.0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, g_sms);
    Timer tm;

    for (int w : std::vector<int>{}) {
        run_imad<1, 0>(tm, w);
        run_imad<2, 0>(tm, w);
        run_imad<4, 0>(tm, w);
        run_imad<8, 0>(tm, w);
        run_imad<16, 0>(tm, w);
        run_imad<8, 1>(tm, w);
        run_imad<16, 1>(tm, w);
    }

    g_L = 3, g_N = 8192, g_T = 44;
    g_q[0] = 0xfffffffff70001ull, g_q[1] = 0xfffffffff78001ull, g_q[2] = 0xfffffffffb4001ull;
    const u32 LN = g_L * g_N;
    g_njobs = (u32)(gib * 1073741824.0 / ((double)g_T * LN * 8));
    g_njobs -= g_njobs % 24;
    size_t wwords = (size_t)g_njobs * g_T * LN, pwords = (size_t)g_T * 2 * LN;
    printf("DB %u jobs x %u terms x %u cols = %.2f GiB\n", g_njobs, g_T, LN, wwords * 8 / 1073741824.0);
    CK(cudaMalloc(&g_out, (size_t)g_njobs * 2 * LN * 8));
    Data d30, d28;
    d30.split = 30, d28.split = 28;
    for (Data *d : { &d30, &d28 }) {
        for (int lay = 0; lay < 2; lay++) {
            CK(cudaMalloc(&d->W[lay], wwords * 8));
            CK(cudaMalloc(&d->P[lay], pwords * 8));
        }
        k_fill_w<0><<<(unsigned)((wwords + 255) / 256), 256>>>(d->W[0], g_njobs, g_T, g_L, g_N, g_q[0], g_q[1], g_q[2], d->split);
        k_fill_w<1><<<(unsigned)((wwords + 255) / 256), 256>>>(d->W[1], g_njobs, g_T, g_L, g_N, g_q[0], g_q[1], g_q[2], d->split);
        k_fill_p<0><<<(unsigned)((pwords + 255) / 256), 256>>>(d->P[0], g_T, g_L, g_N, g_q[0], g_q[1], g_q[2], d->split);
        k_fill_p<1><<<(unsigned)((pwords + 255) / 256), 256>>>(d->P[1], g_T, g_L, g_N, g_q[0], g_q[1], g_q[2], d->split);
    }
    CK(cudaDeviceSynchronize());

    LabArgs a;
    memset(&a, 0, sizeof(a));
    a.out = g_out;
    a.njobs = g_njobs, a.T = g_T, a.L = g_L, a.N = g_N;
    for (int i = 0; i < 3; i++) a.q[i] = make_mod(g_q[i]);
    //          G TS ST FORM LAYOUT CPS PIPE UN SUB
    run_stream<4, 4, 3, 2, 1, 3, 0, 2, 4>(tm, "tile TS4 un2 3cta (best so far)", a, d30, d28);
    run_stream<4, 4, 3, 2, 1, 3, 0, 1, 4>(tm, "tile TS4 un1 3cta", a, d30, d28);
    run_stream<4, 4, 3, 2, 1, 3, 0, 4, 4>(tm, "tile TS4 un4 3cta", a, d30, d28);
    run_stream<4, 4, 3, 2, 0, 3, 0, 2, 4>(tm, "row  TS4 un2 3cta", a, d30, d28);
    run_stream<4, 4, 3, 3, 1, 3, 0, 2, 4>(tm, "tile TS4 un2 3cta asm-split", a, d30, d28);
    run_stream<4, 4, 3, 2, 1, 3, 0, 2, 2>(tm, "tile TS4 sub2 un2 3cta", a, d30, d28);
    run_stream<4, 4, 3, 2, 1, 3, 0, 1, 2>(tm, "tile TS4 sub2 un1 3cta", a, d30, d28);
    run_stream<4, 4, 3, 2, 1, 3, 0, 1, 1>(tm, "tile TS4 sub1 un1 3cta", a, d30, d28);
    run_stream<4, 4, 2, 2, 1, 4, 0, 2, 2>(tm, "tile TS4 sub2 un2 2st 4cta", a, d30, d28);
    run_stream<4, 4, 4, 2, 1, 2, 0, 2, 4>(tm, "tile TS4 un2 4st 2cta", a, d30, d28);
    run_stream<4, 8, 2, 2, 1, 2, 0, 1, 2>(tm, "tile TS8 sub2 un1 2st 2cta", a, d30, d28);
    run_stream<4, 8, 2, 2, 1, 2, 0, 1, 4>(tm, "tile TS8 sub4 un1 2st 2cta", a, d30, d28);
    run_stream<4, 8, 1, 2, 1, 4, 0, 1, 2>(tm, "tile TS8 sub2 un1 1st 4cta", a, d30, d28);
    run_stream<4, 8, 1, 2, 1, 4, 0, 1, 4>(tm, "tile TS8 sub4 un1 1st 4cta", a, d30, d28);
    run_stream<8, 4, 2, 2, 1, 2, 0, 1, 2>(tm, "tile G8 TS4 sub2 un1 2st 2cta", a, d30, d28);
    run_stream<8, 4, 2, 2, 1, 2, 0, 1, 1>(tm, "tile G8 TS4 sub1 un1 2st 2cta", a, d30, d28);
    run_stream<6, 4, 2, 2, 1, 3, 0, 1, 2>(tm, "tile G6 TS4 sub2 un1 2st 3cta", a, d30, d28);
    run_stream<2, 4, 4, 2, 1, 4, 0, 2, 4>(tm, "tile G2 TS4 un2 4st 4cta", a, d30, d28);
    run_stream<2, 8, 3, 2, 1, 4, 0, 1, 4>(tm, "tile G2 TS8 sub4 3st 4cta", a, d30, d28);
    printf("done\n");
    return 0;
}
