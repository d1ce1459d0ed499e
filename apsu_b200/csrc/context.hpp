// DeviceContext: the B200-resident equivalent of APSU's CryptoContext (SEALContext with
// expand_mod_chain=true + Evaluator), common/apsu/crypto_context.h:28-125.
#pragma once
#include "device_ctx.hpp"
#include "params.hpp"
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace apsu_b200 {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define APSU_CUDA_CHECK(expr)                                                                                          \
    do {                                                                                                               \
        cudaError_t e__ = (expr);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
            throw ::apsu_b200::CudaError(std::string(#expr) + ": " + cudaGetErrorString(e__));                         \
    } while (0)

// Kernel launch with a programmatic (PDL) edge to the preceding kernel of the stream: the kernel may be scheduled
// while its predecessor is still running and waits for it in pdl_enter() (modarith.cuh) — ONLY for kernels that call
// pdl_enter() before touching memory.  Inside stream capture the edge becomes a programmatic graph dependency.
// Off unless APSU_B200_PDL=1 (measured: no gain under graph replay, context.cu pdl_enabled).
bool pdl_enabled();
template <typename... P, typename... A>
inline void launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A &&...args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    APSU_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...));
}

// simple owning device buffer
template <typename T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    DBuf(DBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr, o.n = 0; }
    DBuf &operator=(DBuf &&o) noexcept
    {
        if (this != &o) {
            release();
            p = o.p, n = o.n;
            o.p = nullptr, o.n = 0;
        }
        return *this;
    }
    ~DBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr, n = 0;
    }
    void alloc(size_t count)
    {
        release();
        if (count) APSU_CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void ensure(size_t count)
    {
        if (count > n) alloc(count);
    }
    void upload(const std::vector<T> &h, cudaStream_t st)
    {
        ensure(h.size());
        if (!h.empty()) APSU_CUDA_CHECK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    }
};

class DeviceContext {
public:
    DeviceContext(const apsu_b200_params &params, int device);
    ~DeviceContext();

    apsu_b200_params params;
    int device = 0;
    int sms = 0; // SM count of `device`
    uint32_t N = 0, logN = 0, K = 0;
    uint64_t t = 0;
    uint32_t first_L = 0; // primes at the first data level
    uint32_t low_L = 0;   // DB plaintexts / low powers (chain index 2 with PS, 1 without; clamped)
    uint32_t high_L = 0;  // high powers (chain index 1, clamped)
    bool using_keyswitching() const { return K > 1; }

    // modulus table: [coeff_modulus 0..K-1][m_sk][B_0..B_{nB-1}][t]
    std::vector<uint64_t> mod_values;
    uint32_t idx_msk = 0, idx_B0 = 0, idx_t = 0, nB = 0;
    std::vector<DMod> mod_host;
    std::vector<DShoup> inv_n_host, inv_n_w_host;
    DBuf<ulonglong2> twiddles; // [modulus][2][N]
    DBuf<uint32_t> slot_map;   // BatchEncoder index map

    std::vector<LevelConsts> level;      // index by L (1..first_L)
    std::vector<KeySwitchConsts> ks;     // index by L

    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    uint32_t launches = 0; // kernels launched since last reset

    // ---- NTT launches (K2/K3) ----
    // pattern: modulus-table indices; polynomial p uses pattern[p % len]; src/dst index arrays (device,
    // units of one polynomial) may be null for the identity mapping; reduce_input reduces every input
    // word modulo the target modulus first.
    void ntt(const u64 *in, u64 *out, uint32_t count, const std::vector<uint32_t> &pattern, bool inverse,
             const uint32_t *src_idx = nullptr, const uint32_t *dst_idx = nullptr, bool reduce_input = false);
    // transform with a fused element-wise prologue (ntt.cuh: kNttExtend / kNttTensor / kNttKsMac); fused inputs are read
    // from `arena`, polynomial p is written to out_base + (dst_idx ? dst_idx[p] : p) * N
    void ntt_fused(int mode, u64 *arena, uint32_t count, const std::vector<uint32_t> &pattern, const uint32_t *src_idx, const uint32_t *dst_idx, u64 *out_base,
                   const NttFuse &f);
    std::vector<uint32_t> pattern_q(uint32_t L) const;   // q_0..q_{L-1}
    std::vector<uint32_t> pattern_bsk(uint32_t L) const; // B_0..B_{|B|-1}, m_sk for level L
    std::vector<uint32_t> pattern_ks(uint32_t L) const;  // q_0..q_{L-1}, P
    std::vector<uint32_t> pattern_ext(uint32_t L) const; // q_0..q_{L-1}, B_0.., m_sk

private:
    void build_moduli();
    void build_levels();
    NttArgs make_args(const std::vector<uint32_t> &pattern) const;
};

} // namespace apsu_b200
