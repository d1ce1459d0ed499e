// Lab 3 (not product): the PRODUCTION DB-stream kernel (apsu_b200/csrc/db_stream.cuh) on synthetic data, checked
// against a host 128-bit reference, in several launch shapes and job mixes (uniform, partial last stage, ragged,
// long sums that need lane folds, 60-bit primes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I apsu_b200/csrc -o tools/lab/k1_lab3 tools/lab/k1_lab3.cu
#include "db_stream.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

using namespace apsu_b200;

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e = (x);                                                                       \
        if (e != cudaSuccess) {                                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);         \
            exit(1);                                                                               \
        }                                                                                          \
    } while (0)

__host__ __device__ inline u64 splitmix(u64 x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline u64 val_w(u32 job, u32 term, u32 col, u64 q) { return splitmix(((u64)job << 40) ^ ((u64)term << 24) ^ col ^ 0xABCDull << 50) % q; }
__host__ __device__ inline u64 val_p(u32 term, u32 comp, u32 col, u64 q) { return splitmix(((u64)term << 24) ^ ((u64)comp << 60) ^ col ^ 0x77ull << 52) % q; }

// W: every job owns a buffer of T rows, tile-major: ((job*ntiles + tile)*T + term)*128 + c
__global__ void k_fill_w(u64 *W, u32 njobs, u32 T, u32 L, u32 N, u64 q0, u64 q1, u64 q2, int split)
{
    const u32 LN = L * N, ntiles = LN / 128;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)njobs * T * LN;
    if (i >= total) return;
    u32 c = i % 128, term = (i / 128) % T, tile = (i / ((size_t)128 * T)) % ntiles, job = i / ((size_t)128 * T * ntiles);
    u32 col = tile * 128 + c, l = col / N;
    u64 q = l == 0 ? q0 : l == 1 ? q1 : q2;
    W[i] = split_word(val_w(job, term, col, q), split);
}
__global__ void k_fill_p(u64 *P, u32 T, u32 L, u32 N, u64 q0, u64 q1, u64 q2, int split)
{
    const u32 LN = L * N;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)T * 2 * LN;
    if (i >= total) return;
    u32 c = i % 128, comp = (i / 128) % 2, term = (i / 256) % T, tile = i / ((size_t)256 * T);
    u32 col = tile * 128 + c, l = col / N;
    u64 q = l == 0 ? q0 : l == 1 ? q1 : q2;
    P[i] = split_word(val_p(term, comp, col, q), split);
}

static DMod make_mod(u64 q)
{
    DMod m;
    m.q = q;
    unsigned __int128 r = (~(unsigned __int128)0) / q;
    m.r0 = (u64)r;
    m.r1 = (u64)(r >> 64);
    int k = 0;
    while ((q >> k) != 0) k++;
    m.sh = (u32)(k - 1);
    m.mu = (u64)((((unsigned __int128)1) << (64 + m.sh)) / q);
    m.pad_ = 0;
    return m;
}
static int bits(u64 v)
{
    int b = 0;
    while (v) b++, v >>= 1;
    return b;
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

struct Scenario {
    const char *name;
    u32 T;                  // rows per job buffer and terms in the power table
    std::vector<u32> lens;  // job lengths, cycled over the jobs
    u64 q[3];
    double gib;
};

static int g_sms;

template <int STAGES, int MINB>
static void run(Timer &tm, const Scenario &sc)
{
    const u32 L = 3, N = 8192, LN = L * N, ntiles = LN / 128, T = sc.T;
    u32 njobs = (u32)(sc.gib * 1073741824.0 / ((double)T * LN * 8));
    njobs = std::max(8u, njobs - njobs % 8);
    int b = 0;
    for (int i = 0; i < 3; i++) b = std::max(b, bits(sc.q[i]));
    const int split = (b + 1) / 2;
    const unsigned __int128 one = 1;
    const unsigned __int128 sum_max = ((one << split) - 1) + ((one << (b - split)) - 1);
    const unsigned __int128 cap = (((one << 64) - 1) - (one << (split + 1))) / (sum_max * sum_max);
    const u32 fold_stages = (u32)std::min<unsigned __int128>(cap / kKtTS, 0x7FFFFFFF);

    // arena A = [P table: T*2 polys-of-N... in words][out: njobs*2*LN]
    const size_t pwords = (size_t)T * 2 * LN, owords = (size_t)njobs * 2 * LN, wwords = (size_t)njobs * T * LN;
    u64 *A, *W;
    CK(cudaMalloc(&A, (pwords + owords) * 8));
    CK(cudaMalloc(&W, wwords * 8));
    k_fill_w<<<(unsigned)((wwords + 255) / 256), 256>>>(W, njobs, T, L, N, sc.q[0], sc.q[1], sc.q[2], split);
    k_fill_p<<<(unsigned)((pwords + 255) / 256), 256>>>(A, T, L, N, sc.q[0], sc.q[1], sc.q[2], split);
    CK(cudaMemset(A + pwords, 0xFF, owords * 8));

    std::vector<u32> len(njobs);
    for (u32 j = 0; j < njobs; j++) len[j] = sc.lens[j % sc.lens.size()];
    // group like the engine: longest first, four to a group
    std::vector<u32> order(njobs);
    for (u32 j = 0; j < njobs; j++) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](u32 x, u32 y) { return len[x] > len[y]; });
    std::vector<KtGroup> groups;
    double bytes = 0;
    for (u32 i = 0; i < njobs; i += kKtG) {
        KtGroup g;
        memset(&g, 0, sizeof(g));
        g.p_idx = 0;
        g.pstride = T;
        bool same = true;
        for (u32 k = i; k < std::min(njobs, i + kKtG); k++) {
            u32 j = order[k];
            g.w[g.njobs] = W + (size_t)j * ntiles * T * 128;
            g.wstride[g.njobs] = T;
            g.nterms[g.njobs] = len[j];
            g.out_idx[g.njobs] = (u32)((pwords + (size_t)j * 2 * LN) / N);
            g.max_terms = std::max(g.max_terms, len[j]);
            same &= len[j] == len[order[i]];
            bytes += (double)len[j] * LN * 8;
            g.njobs++;
        }
        for (u32 k = g.njobs; k < (u32)kKtG; k++) g.w[k] = g.w[0];
        g.ragged = (same && g.njobs == (u32)kKtG) ? 0 : 1;
        groups.push_back(g);
    }
    KtGroup *gd;
    CK(cudaMalloc(&gd, groups.size() * sizeof(KtGroup)));
    CK(cudaMemcpy(gd, groups.data(), groups.size() * sizeof(KtGroup), cudaMemcpyHostToDevice));
    LevelConsts c;
    memset(&c, 0, sizeof(c));
    c.L = L;
    for (int i = 0; i < 3; i++) c.q[i] = make_mod(sc.q[i]);

    auto kern = k_db_mac_kt<STAGES, MINB>;
    constexpr size_t smem = kt_smem_bytes(STAGES);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kKtThreads, smem));
    per_sm = std::min(per_sm, MINB);
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    const u32 items = (u32)groups.size() * ntiles;
    const u32 grid = std::min<u32>(items, g_sms * per_sm);
    float ms = tm.run([&] { kern<<<grid, kKtThreads, smem>>>(A, gd, (u32)groups.size(), c, (int)N, split, fold_stages, 0u); });

    std::vector<u64> h(owords);
    CK(cudaMemcpy(h.data(), A + pwords, owords * 8, cudaMemcpyDeviceToHost));
    int bad = 0, n = 0;
    for (u32 job : { 0u, 1u, 2u, 3u, njobs / 2 + 1, njobs - 2, njobs - 1 })
        for (u32 col : { 0u, 1u, 127u, 128u, N - 1, N, LN / 2 + 77, LN - 1 })
            for (u32 comp = 0; comp < 2; comp++) {
                u64 q = sc.q[col / N];
                unsigned __int128 s = 0;
                for (u32 t = 0; t < len[job]; t++) s = (s + (unsigned __int128)val_w(job, t, col, q) * val_p(t, comp, col, q)) % q;
                u64 got = h[((size_t)job * 2 + comp) * LN + col];
                n++;
                if (got != (u64)s) {
                    if (bad < 2) printf("   MISMATCH job %u (len %u) col %u comp %u: got %llx want %llx\n", job, len[job], col, comp, got, (u64)s);
                    bad++;
                }
            }
    printf("%-34s ST=%d cta/sm=%d regs=%d split=%d fold=%u groups=%zu : %7.3f ms  %7.1f GB/s  %s\n", sc.name, STAGES, per_sm, fa.numRegs, split, fold_stages,
           groups.size(), ms, bytes / ms / 1e6, bad ? "WRONG" : "OK");
    fflush(stdout);
    cudaFree(A);
    cudaFree(W);
    cudaFree(gd);
}

int main(int argc, char **argv)
{
    double gib = argc > 1 ? atof(argv[1]) : 4.0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, g_sms);
    Timer tm;
    const u64 q56[3] = { 0xfffffffff70001ull, 0xfffffffff78001ull, 0xfffffffffb4001ull };
    const u64 q50[3] = { 0x3ffffffef4001ull, 0x3fffffffcc001ull, 0x3ffffffffc001ull };
    const u64 q60[3] = { 0xffffffffffc0001ull, 0xfffffffff840001ull, 0xfffffffff6c0001ull }; // 60-bit, = 1 mod 2^18
    std::vector<Scenario> scs;
    scs.push_back({ "uniform T=44 56-bit (16M-4096)", 44, { 44 }, { q56[0], q56[1], q56[2] }, gib });
    scs.push_back({ "uniform T=43 (partial last stage)", 44, { 43 }, { q56[0], q56[1], q56[2] }, gib });
    scs.push_back({ "16M mix 28x44 + 43", 44, { 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 43, 44 },
                    { q56[0], q56[1], q56[2] }, gib });
    scs.push_back({ "ragged 44/18/7/1/0", 44, { 44, 18, 7, 1, 0, 33 }, { q56[0], q56[1], q56[2] }, gib / 4 });
    scs.push_back({ "uniform T=310 50-bit (256M-4096)", 310, { 310 }, { q50[0], q50[1], q50[2] }, gib });
    scs.push_back({ "long T=300 56-bit (folds)", 300, { 300, 299, 130 }, { q56[0], q56[1], q56[2] }, gib / 2 });
    scs.push_back({ "60-bit primes T=44 (fold/stage)", 44, { 44, 43 }, { q60[0], q60[1], q60[2] }, gib / 4 });
    for (auto &sc : scs) {
        run<4, 2>(tm, sc);
        run<3, 3>(tm, sc);
        run<3, 2>(tm, sc);
        run<2, 4>(tm, sc);
    }
    printf("done\n");
    return 0;
}
