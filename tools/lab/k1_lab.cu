// Lab (not product): micro-benchmarks that decide the design of the DB-stream kernel K1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I apsu_b200/csrc -o tools/lab/k1_lab tools/lab/k1_lab.cu
//   gpurun -- ./tools/lab/k1_lab [GiB of DB]
// Sections: (1) integer-pipe rates of the MAC idioms, (2) pure-read ceilings (LDG / bulk-copy ring at several
// copy sizes), (3) stream-MAC variants checked against a host 128-bit reference.
#include "modarith.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>

using namespace apsu_b200;

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e = (x);                                                                       \
        if (e != cudaSuccess) {                                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);         \
            exit(1);                                                                               \
        }                                                                                          \
    } while (0)

// ------------------------------------------------------------------------------------------------
// (1) MAC idiom rates
// ------------------------------------------------------------------------------------------------
template <int FORM>
__global__ void __launch_bounds__(256) k_imad(u64 *out, int iters, u32 seed)
{
    u32 a0 = seed + threadIdx.x, a1 = a0 * 3 + 1, b0 = a0 ^ 0x55555, b1 = a1 ^ 0x33333;
    u64 acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
            if (FORM == 0) {
                acc[k] += (u64)a0 * b0;
                acc[k + 1] += (u64)a0 * b1;
                acc[k + 2] += (u64)a1 * b0;
                acc[k + 3] += (u64)a1 * b1;
            } else {
                asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a0), "r"(b0));
                asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k + 1]) : "r"(a0), "r"(b1));
                asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k + 2]) : "r"(a1), "r"(b0));
                asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k + 3]) : "r"(a1), "r"(b1));
            }
            a0 += 0x9E37u; // keep the operands changing (cheap ALU op)
            b1 ^= a0;
        }
    }
    u64 x = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) x ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// ------------------------------------------------------------------------------------------------
// (2)/(3) stream kernels
// ------------------------------------------------------------------------------------------------
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

// pure read, LDG.128 grid-stride, UNROLL independent loads in flight per thread
template <int UNROLL>
__global__ void __launch_bounds__(256) k_read_ldg(const ulonglong2 *__restrict__ p, size_t n, u64 *out)
{
    u64 x = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
        ulonglong2 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) x ^= v[u].x ^ v[u].y;
    }
    if (x == 0x1234567) out[0] = x;
}

// pure read through a bulk-copy ring: each CTA streams contiguous chunks of CHUNK bytes, COPY bytes per bulk op
template <int STAGES>
__global__ void __launch_bounds__(160) k_read_ring(const u64 *__restrict__ p, size_t n_chunks, int chunk_bytes, int copy_bytes, u64 *out)
{
    extern __shared__ __align__(128) u64 smem[];
    u64 *full = smem + (size_t)STAGES * chunk_bytes / 8;
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
    u32 it = 0;
    if (tid >= 128) {
        if (tid != 128) return;
        const u64 pol = pol_evict_first();
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, it++) {
            const int s = it % STAGES;
            const u32 use = it / STAGES;
            if (use) mbar_wait(&empty[s], (use - 1) & 1);
            mbar_expect_tx(&full[s], chunk_bytes);
            const char *src = (const char *)p + c * (size_t)chunk_bytes;
            char *dst = (char *)smem + (size_t)s * chunk_bytes;
            for (int o = 0; o < chunk_bytes; o += copy_bytes) bulk_g2s(dst + o, src + o, copy_bytes, &full[s], pol);
        }
        return;
    }
    u64 x = 0;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, it++) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        const u64 *sp = smem + (size_t)s * chunk_bytes / 8;
        for (int o = tid; o < chunk_bytes / 8; o += 128) x ^= sp[o];
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }
    if (x == 0x1234567) out[0] = x;
}

// ---- stream MAC lab kernel ----
// logical data: W[job][term][col] (col in [0, LN)), P[term][comp][col].
// LAYOUT 0 (row-major, as the product today): W at ((job*T + term)*LN + col), P at ((term*2+comp)*LN + col)
// LAYOUT 1 (tile-major): W at (((job*ntiles + tile)*T + term)*COLS + c), P at (((tile*T + term)*2 + comp)*COLS + c)
struct Acc3 {
    u64 ll, mid, hh, m2;
};
template <int FORM>
__device__ __forceinline__ void mac3(Acc3 &a, u32 wl, u32 wh, u32 pl, u32 ph)
{
    if (FORM == 0) {
        a.ll += (u64)wl * pl;
        a.mid += (u64)wl * ph;
        a.mid += (u64)wh * pl;
        a.hh += (u64)wh * ph;
    } else if (FORM == 2) { // four independent lanes: every product is one fused IMAD.WIDE
        a.ll += (u64)wl * pl;
        a.mid += (u64)wl * ph;
        a.m2 += (u64)wh * pl;
        a.hh += (u64)wh * ph;
    } else {
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.ll) : "r"(wl), "r"(pl));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(wl), "r"(ph));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(wh), "r"(pl));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.hh) : "r"(wh), "r"(ph));
    }
}
__device__ __forceinline__ void normalize3(Acc3 &a)
{
    a.mid += a.m2;
    a.m2 = 0;
    a.mid += a.ll >> 30;
    a.ll &= 0x3FFFFFFFull;
    a.hh += a.mid >> 30;
    a.mid &= 0x3FFFFFFFull;
}
__device__ __forceinline__ u64 reduce3(const Acc3 &a, const DMod &m)
{
    u64 lo = a.ll, hi = 0;
    u64 t = a.mid << 30;
    lo += t;
    hi += (a.mid >> 34) + (lo < t);
    t = a.hh << 60;
    lo += t;
    hi += (a.hh >> 4) + (lo < t);
    return barrett128(lo, hi, m);
}

struct LabArgs {
    const u64 *W;
    const u64 *P;
    u64 *out; // [job][2][LN]
    u32 njobs, T, L, N;
    DMod q[4];
    u32 norm_terms; // terms between lane renormalisations
};

// MODE 0: full MAC; 1: loads from smem only (xor), no MAC; 2: barrier handshake only (no smem reads)
template <int G, int TS, int STAGES, int FORM, int MODE, int LAYOUT, int CPS>
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
    const u32 nst = (T + TS - 1) / TS;
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
                const u32 t0 = st * TS, nt = min((u32)TS, T - t0);
                mbar_expect_tx(&full[s], nt * term_words * 8);
                u64 *sb = ring + (size_t)s * stage_words;
                // stage layout: [P: TS*2*COLS][W job0: TS*COLS] ... [W job G-1]
                if (LAYOUT == 0) {
                    for (u32 h = 0; h < nt; h++) {
                        bulk_g2s(sb + (h * 2) * COLS, a.P + ((size_t)(t0 + h) * 2) * LN + tile * COLS, COLS * 8, &full[s], pol_k);
                        bulk_g2s(sb + (h * 2 + 1) * COLS, a.P + ((size_t)(t0 + h) * 2 + 1) * LN + tile * COLS, COLS * 8, &full[s], pol_k);
                    }
                    for (int k = 0; k < G; k++)
                        for (u32 h = 0; h < nt; h++)
                            bulk_g2s(sb + (TS * 2 + k * TS + h) * COLS, a.W + ((size_t)(job0 + k) * T + t0 + h) * LN + tile * COLS, COLS * 8, &full[s], pol_s);
                } else {
                    bulk_g2s(sb, a.P + ((size_t)tile * T + t0) * 2 * COLS, nt * 2 * COLS * 8, &full[s], pol_k);
                    for (int k = 0; k < G; k++)
                        bulk_g2s(sb + (TS * 2 + k * TS) * COLS, a.W + (((size_t)(job0 + k) * ntiles + tile) * T + t0) * COLS, nt * COLS * 8, &full[s], pol_s);
                }
            }
        }
        return;
    }
    for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
        const u32 tile = item % ntiles, job0 = (item / ntiles) * G;
        const u32 col0 = tile * COLS;
        const DMod m = a.q[col0 / a.N];
        Acc3 acc[G][2];
#pragma unroll
        for (int k = 0; k < G; k++) acc[k][0] = acc[k][1] = Acc3{ 0, 0, 0, 0 };
        u64 x = 0;
        u32 since = 0;
        for (u32 st = 0; st < nst; st++, it++) {
            const int s = it % STAGES;
            mbar_wait(&full[s], (it / STAGES) & 1);
            const u64 *sb = ring + (size_t)s * stage_words + tid;
            const u32 t0 = st * TS, nt = min((u32)TS, T - t0);
            if (MODE == 2) {
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty[s]);
                continue;
            }
            u64 p[TS][2], w[G][TS];
#pragma unroll
            for (int h = 0; h < TS; h++) {
                p[h][0] = sb[(h * 2) * COLS];
                p[h][1] = sb[(h * 2 + 1) * COLS];
#pragma unroll
                for (int k = 0; k < G; k++) w[k][h] = sb[(TS * 2 + k * TS + h) * COLS];
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
            if (MODE == 1) {
#pragma unroll
                for (int h = 0; h < TS; h++) {
                    x ^= p[h][0] ^ p[h][1];
#pragma unroll
                    for (int k = 0; k < G; k++) x ^= w[k][h];
                }
                continue;
            }
#pragma unroll
            for (int h = 0; h < TS; h++) {
                const bool live = (u32)h < nt;
                const u32 p0l = (u32)p[h][0] & 0x3FFFFFFFu, p0h = (u32)(p[h][0] >> 30);
                const u32 p1l = (u32)p[h][1] & 0x3FFFFFFFu, p1h = (u32)(p[h][1] >> 30);
#pragma unroll
                for (int k = 0; k < G; k++) {
                    const u64 ww = live ? w[k][h] : 0ull;
                    const u32 wl = (u32)ww, wh = (u32)(ww >> 32);
                    mac3<FORM>(acc[k][0], wl, wh, p0l, p0h);
                    mac3<FORM>(acc[k][1], wl, wh, p1l, p1h);
                }
            }
            since += TS;
            if (since + TS > a.norm_terms) {
                since = 0;
#pragma unroll
                for (int k = 0; k < G; k++) {
                    normalize3(acc[k][0]);
                    normalize3(acc[k][1]);
                }
            }
        }
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < G; k++) {
                normalize3(acc[k][0]);
                normalize3(acc[k][1]);
                u64 *o = a.out + ((size_t)(job0 + k) * 2) * LN + col0 + tid;
                o[0] = reduce3(acc[k][0], m);
                o[LN] = reduce3(acc[k][1], m);
            }
        } else if (x == 0x1234567) {
            a.out[0] = x;
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
__host__ __device__ inline u64 pack30(u64 w) { return (w & 0x3FFFFFFFull) | ((w >> 30) << 32); }

template <int LAYOUT>
__global__ void k_fill_w(u64 *W, u32 njobs, u32 T, u32 L, u32 N, DMod q0, DMod q1, DMod q2, DMod q3)
{
    const u32 LN = L * N, ntiles = LN / 128;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)njobs * T * LN;
    if (i >= total) return;
    u32 col = i % LN, term = (i / LN) % T, job = i / ((size_t)LN * T);
    u32 l = col / N;
    u64 q = l == 0 ? q0.q : l == 1 ? q1.q : l == 2 ? q2.q : q3.q;
    u64 v = pack30(val_w(job, term, col, q));
    size_t dst = LAYOUT == 0 ? i : ((((size_t)job * ntiles + col / 128) * T + term) * 128 + col % 128);
    W[dst] = v;
}
template <int LAYOUT>
__global__ void k_fill_p(u64 *P, u32 T, u32 L, u32 N, DMod q0, DMod q1, DMod q2, DMod q3)
{
    const u32 LN = L * N;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)T * 2 * LN;
    if (i >= total) return;
    u32 col = i % LN, comp = (i / LN) % 2, term = i / ((size_t)LN * 2);
    u32 l = col / N;
    u64 q = l == 0 ? q0.q : l == 1 ? q1.q : l == 2 ? q2.q : q3.q;
    size_t dst = LAYOUT == 0 ? i : ((((size_t)(col / 128) * T + term) * 2 + comp) * 128 + col % 128);
    P[dst] = val_p(term, comp, col, q);
}

static DMod make_mod(u64 q)
{
    DMod m;
    m.q = q;
    unsigned __int128 num = ~(unsigned __int128)0; // floor((2^128-1)/q) == floor(2^128/q) for odd q > 1
    unsigned __int128 r = num / q;
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
    // compare a sample of outputs with a host 128-bit reference
    const u32 LN = g_L * g_N;
    std::vector<u64> h((size_t)g_njobs * 2 * LN);
    CK(cudaMemcpy(h.data(), g_out, h.size() * 8, cudaMemcpyDeviceToHost));
    int bad = 0, n = 0;
    for (u32 job : { 0u, 1u, g_njobs / 2 + 1, g_njobs - 1 })
        for (u32 col : { 0u, 1u, 127u, 128u, g_N - 1, g_N, LN / 2 + 77, LN - 1 })
            for (u32 comp = 0; comp < 2; comp++) {
                u64 q = g_q[col / g_N];
                unsigned __int128 s = 0;
                for (u32 t = 0; t < g_T; t++) s = (s + (unsigned __int128)val_w(job, t, col, q) * val_p(t, comp, col, q)) % q;
                u64 got = h[((size_t)job * 2 + comp) * LN + col];
                n++;
                if (got != (u64)s) {
                    if (bad < 3) printf("   MISMATCH %s job %u col %u comp %u: got %llx want %llx\n", name, job, col, comp, got, (u64)s);
                    bad++;
                }
            }
    return bad == 0;
}

template <int G, int TS, int STAGES, int FORM, int MODE, int LAYOUT, int CPS>
static void run_stream(Timer &tm, const char *name, LabArgs a, const u64 *W0, const u64 *P0, const u64 *W1, const u64 *P1)
{
    constexpr size_t smem = (size_t)STAGES * TS * (2 + G) * 128 * 8 + 2 * STAGES * 8 + 16;
    auto kern = k_stream<G, TS, STAGES, FORM, MODE, LAYOUT, CPS>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 160, smem));
    if (per_sm > CPS) per_sm = CPS;
    a.W = LAYOUT ? W1 : W0;
    a.P = LAYOUT ? P1 : P0;
    CK(cudaMemset(g_out, 0, (size_t)g_njobs * 2 * g_L * g_N * 8));
    float ms = tm.run([&] { kern<<<g_sms * per_sm, 160, smem>>>(a); });
    double bytes = (double)g_njobs * g_T * g_L * g_N * 8;
    bool ok = MODE != 0 || check(name);
    printf("stream %-44s G=%d TS=%d ST=%d form=%d mode=%d layout=%d cta/sm=%d smem=%zuK : %7.3f ms  %7.1f GB/s  %s\n", name, G, TS, STAGES, FORM, MODE, LAYOUT,
           per_sm, smem / 1024, ms, bytes / ms / 1e6, MODE ? "-" : ok ? "OK" : "WRONG");
    fflush(stdout);
}

int main(int argc, char **argv)
{
    double gib = argc > 1 ? atof(argv[1]) : 4.0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, g_sms, prop.clockRate);
    Timer tm;

    // (1) integer rates
    {
        u64 *o;
        CK(cudaMalloc(&o, (size_t)g_sms * 8 * 256 * 8));
        const int iters = 4096;
        for (int form = 0; form < 2; form++) {
            float ms = tm.run([&] {
                if (form == 0)
                    k_imad<0><<<g_sms * 8, 256>>>(o, iters, 1);
                else
                    k_imad<1><<<g_sms * 8, 256>>>(o, iters, 1);
            });
            double macs = (double)g_sms * 8 * 256 * iters * 16;
            printf("imad form %d (%s): %.3f ms, %.1f G 32x32+64 MAC/s, %.2f MAC/clk/SM @1965MHz\n", form, form ? "asm mad.wide (split by ptxas)" : "C (fused IMAD.WIDE)",
                   ms, macs / ms / 1e6, macs / ms / 1e6 * 1e9 / g_sms / 1.965e9);
        }
        cudaFree(o);
    }

    // data
    g_L = 3, g_N = 8192, g_T = 44;
    g_q[0] = 0xfffffffff70001ull, g_q[1] = 0xfffffffff78001ull, g_q[2] = 0xfffffffffb4001ull, g_q[3] = 0x3ffffffffc001ull;
    const u32 LN = g_L * g_N;
    g_njobs = (u32)(gib * 1073741824.0 / ((double)g_T * LN * 8));
    g_njobs -= g_njobs % 8;
    size_t wwords = (size_t)g_njobs * g_T * LN, pwords = (size_t)g_T * 2 * LN;
    printf("DB %u jobs x %u terms x %u cols = %.2f GiB\n", g_njobs, g_T, LN, wwords * 8 / 1073741824.0);
    u64 *W0, *W1, *P0, *P1;
    CK(cudaMalloc(&W0, wwords * 8));
    CK(cudaMalloc(&W1, wwords * 8));
    CK(cudaMalloc(&P0, pwords * 8));
    CK(cudaMalloc(&P1, pwords * 8));
    CK(cudaMalloc(&g_out, (size_t)g_njobs * 2 * LN * 8));
    DMod q[4];
    for (int i = 0; i < 4; i++) q[i] = make_mod(g_q[i]);
    k_fill_w<0><<<(unsigned)((wwords + 255) / 256), 256>>>(W0, g_njobs, g_T, g_L, g_N, q[0], q[1], q[2], q[3]);
    k_fill_w<1><<<(unsigned)((wwords + 255) / 256), 256>>>(W1, g_njobs, g_T, g_L, g_N, q[0], q[1], q[2], q[3]);
    k_fill_p<0><<<(unsigned)((pwords + 255) / 256), 256>>>(P0, g_T, g_L, g_N, q[0], q[1], q[2], q[3]);
    k_fill_p<1><<<(unsigned)((pwords + 255) / 256), 256>>>(P1, g_T, g_L, g_N, q[0], q[1], q[2], q[3]);
    CK(cudaDeviceSynchronize());

    // (2) read ceilings
    {
        double bytes = (double)wwords * 8;
        float ms = tm.run([&] { k_read_ldg<4><<<g_sms * 8, 256>>>((const ulonglong2 *)W0, wwords / 2, g_out); });
        printf("read LDG.128 x4  grid 8/SM : %.3f ms %.1f GB/s\n", ms, bytes / ms / 1e6);
        ms = tm.run([&] { k_read_ldg<8><<<g_sms * 8, 256>>>((const ulonglong2 *)W0, wwords / 2, g_out); });
        printf("read LDG.128 x8  grid 8/SM : %.3f ms %.1f GB/s\n", ms, bytes / ms / 1e6);
        ms = tm.run([&] { k_read_ldg<8><<<g_sms * 4, 256>>>((const ulonglong2 *)W0, wwords / 2, g_out); });
        printf("read LDG.128 x8  grid 4/SM : %.3f ms %.1f GB/s\n", ms, bytes / ms / 1e6);
        for (int chunk : { 8192, 16384, 32768 })
            for (int copy : { 1024, 4096, chunk }) {
                for (int cps : { 2, 4 }) {
                    size_t smem = (size_t)4 * chunk + 64 + 16;
                    if (smem * cps > 220 * 1024) continue;
                    CK(cudaFuncSetAttribute(k_read_ring<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    size_t n_chunks = wwords * 8 / chunk;
                    ms = tm.run([&] { k_read_ring<4><<<g_sms * cps, 160, smem>>>(W0, n_chunks, chunk, copy, g_out); });
                    printf("read ring 4 stages chunk %5d copy %5d cta/sm %d : %.3f ms %.1f GB/s\n", chunk, copy, cps, ms, bytes / ms / 1e6);
                }
            }
    }

    // (3) stream variants
    LabArgs a;
    memset(&a, 0, sizeof(a));
    a.out = g_out;
    a.njobs = g_njobs, a.T = g_T, a.L = g_L, a.N = g_N;
    for (int i = 0; i < 4; i++) a.q[i] = q[i];
    a.norm_terms = 14;
    //          G TS ST FORM MODE LAYOUT CPS
    run_stream<4, 2, 4, 1, 0, 0, 4>(tm, "today: asm MAC, row-major 1KB copies", a, W0, P0, W1, P1);
    run_stream<4, 2, 4, 0, 0, 0, 4>(tm, "fused MAC, row-major", a, W0, P0, W1, P1);
    run_stream<4, 2, 4, 0, 1, 0, 4>(tm, "no MAC (smem reads), row-major", a, W0, P0, W1, P1);
    run_stream<4, 2, 4, 0, 2, 0, 4>(tm, "handshake only, row-major", a, W0, P0, W1, P1);
    run_stream<4, 2, 4, 0, 2, 1, 4>(tm, "handshake only, tile-major", a, W0, P0, W1, P1);
    run_stream<4, 4, 4, 0, 2, 1, 2>(tm, "handshake only, tile-major TS4", a, W0, P0, W1, P1);
    run_stream<4, 2, 4, 0, 1, 1, 4>(tm, "no MAC, tile-major", a, W0, P0, W1, P1);
    run_stream<4, 2, 4, 0, 0, 1, 4>(tm, "fused MAC, tile-major", a, W0, P0, W1, P1);
    run_stream<4, 2, 4, 2, 0, 1, 4>(tm, "4-lane fused MAC, tile-major", a, W0, P0, W1, P1);
    run_stream<4, 2, 4, 2, 0, 1, 3>(tm, "4-lane fused MAC, tile-major 3cta", a, W0, P0, W1, P1);
    run_stream<4, 2, 4, 2, 0, 0, 3>(tm, "4-lane fused MAC, row-major 3cta", a, W0, P0, W1, P1);
    run_stream<4, 2, 4, 1, 0, 1, 4>(tm, "asm MAC, tile-major", a, W0, P0, W1, P1);
    run_stream<4, 4, 4, 0, 0, 1, 2>(tm, "fused MAC, tile-major TS4", a, W0, P0, W1, P1);
    run_stream<4, 4, 3, 0, 0, 1, 3>(tm, "fused MAC, tile-major TS4 3 stages", a, W0, P0, W1, P1);
    run_stream<8, 2, 4, 0, 0, 1, 2>(tm, "fused MAC, tile-major G8", a, W0, P0, W1, P1);
    run_stream<8, 2, 3, 0, 0, 1, 3>(tm, "fused MAC, tile-major G8 3 stages", a, W0, P0, W1, P1);
    run_stream<8, 2, 4, 0, 0, 0, 2>(tm, "fused MAC, row-major G8", a, W0, P0, W1, P1);
    run_stream<2, 4, 4, 0, 0, 1, 4>(tm, "fused MAC, tile-major G2 TS4", a, W0, P0, W1, P1);
    run_stream<4, 1, 8, 0, 0, 1, 4>(tm, "fused MAC, tile-major TS1 8 stages", a, W0, P0, W1, P1);
    printf("done\n");
    return 0;
}
