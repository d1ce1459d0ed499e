#include "context.hpp"
#include "hostmath.hpp"
#include "ntt.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace apsu_b200 {

using hm::invm;
using hm::mulm;

template <int LOGN>
static void configure_ntt(); // below, with the launch shapes

static DMod make_mod(uint64_t q)
{
    DMod m;
    m.q = q;
    // floor(2^128 / q) in two words
    hm::u128 top = (hm::u128)1 << 64;
    uint64_t hi = (uint64_t)(top / q);
    hm::u128 rem = top % q;
    uint64_t lo = (uint64_t)((rem << 64) / q);
    m.r0 = lo;
    m.r1 = hi;
    m.sh = (u32)(hm::bit_length(q) - 1);
    m.mu = (uint64_t)(((hm::u128)1 << (64 + m.sh)) / q); // < 2^64 because q > 2^sh (q is not a power of two)
    m.pad_ = 0;
    return m;
}
static DShoup make_shoup(uint64_t v, uint64_t q)
{
    v %= q;
    DShoup s;
    s.op = v;
    s.quot = (uint64_t)(((hm::u128)v << 64) / q);
    return s;
}

DeviceContext::DeviceContext(const apsu_b200_params &p, int dev) : params(p), device(dev)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw CudaError(std::string("no CUDA device available (apsu_b200 has no CPU fallback): ") + cudaGetErrorString(e));
    if (dev < 0 || dev >= count) throw std::invalid_argument("device ordinal is out of range");
    APSU_CUDA_CHECK(cudaSetDevice(dev));
    N = p.poly_modulus_degree;
    logN = 0;
    while ((1u << logN) < N) logN++;
    if (logN < 11 || logN > 14) throw std::invalid_argument("poly_modulus_degree must be between 2048 and 16384");
    K = p.coeff_modulus_count;
    t = p.plain_modulus;
    if (K > (uint32_t)kMaxKey) throw std::invalid_argument("too many coeff_modulus primes for this build");
    first_L = K > 1 ? K - 1 : 1;
    // get_parms_id_for_chain_idx clamps to the first data level; a data level with chain index c has c+1 primes
    auto level_for_chain = [&](uint32_t c) { return std::min(first_L, c + 1); };
    low_L = level_for_chain(std::min<uint32_t>(first_L - 1, p.ps_low_degree ? 2 : 1));
    high_L = level_for_chain(1);
    APSU_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    owns_stream = true;
    APSU_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    switch (logN) {
    case 11: configure_ntt<11>(); break;
    case 12: configure_ntt<12>(); break;
    case 13: configure_ntt<13>(); break;
    default: configure_ntt<14>(); break;
    }
    build_moduli();
    build_levels();
    APSU_CUDA_CHECK(cudaStreamSynchronize(stream));
}

DeviceContext::~DeviceContext()
{
    if (owns_stream && stream) cudaStreamDestroy(stream);
}

void DeviceContext::build_moduli()
{
    // BEHZ auxiliary bases (RNSTool::initialize): |B| = L (+1 if 32 + bits(t) + bits(q) >= 61 L + 61)
    uint32_t maxB = 0;
    for (uint32_t L = 1; L <= first_L; L++) {
        hm::Wide q;
        for (uint32_t i = 0; i < L; i++) q.mul(params.coeff_modulus[i]);
        uint32_t nb = L + ((32 + hm::bit_length(t) + q.bits() >= 61 * (int)L + 61) ? 1 : 0);
        maxB = std::max(maxB, nb);
    }
    if (maxB + 1 > (uint32_t)kMaxBsk || first_L > (uint32_t)kMaxQ) throw std::invalid_argument("modulus chain too long for this build");
    nB = maxB;
    auto aux = hm::primes_below_pow2(2ull * N, 61, maxB + 2); // m_sk, gamma, B_0, B_1, ...
    mod_values.assign(params.coeff_modulus, params.coeff_modulus + K);
    idx_msk = (uint32_t)mod_values.size();
    mod_values.push_back(aux[0]);
    idx_B0 = (uint32_t)mod_values.size();
    for (uint32_t i = 0; i < maxB; i++) mod_values.push_back(aux[2 + i]);
    idx_t = (uint32_t)mod_values.size();
    mod_values.push_back(t);

    size_t M = mod_values.size();
    std::vector<ulonglong2> tw(M * 2 * N);
    mod_host.resize(M);
    inv_n_host.resize(M);
    inv_n_w_host.resize(M);
    for (size_t m = 0; m < M; m++) {
        uint64_t q = mod_values[m];
        mod_host[m] = make_mod(q);
        uint64_t psi = hm::min_primitive_root(2ull * N, q), ipsi = invm(psi, q);
        uint64_t f = 1, b = 1;
        for (uint32_t i = 0; i < N; i++) {
            unsigned r = hm::bitrev(i, (int)logN);
            DShoup sf = make_shoup(f, q), sb = make_shoup(b, q);
            tw[(m * 2 + 0) * N + r] = make_ulonglong2(sf.op, sf.quot);
            tw[(m * 2 + 1) * N + r] = make_ulonglong2(sb.op, sb.quot);
            f = mulm(f, psi, q);
            b = mulm(b, ipsi, q);
        }
        inv_n_host[m] = make_shoup(invm(N % q, q), q);
        // inverse table entry 1 = psi^-bitrev(1) = psi^-(N/2): the only twiddle of the last Gentleman-Sande stage
        inv_n_w_host[m] = make_shoup(mulm(tw[(m * 2 + 1) * N + 1].x, invm(N % q, q), q), q);
    }
    twiddles.upload(tw, stream);

    // BatchEncoder slot -> coefficient-index map
    std::vector<uint32_t> map(N);
    uint32_t row = N >> 1, m2 = N << 1;
    uint64_t pos = 1;
    for (uint32_t i = 0; i < row; i++) {
        map[i] = hm::bitrev((unsigned)((pos - 1) >> 1), (int)logN);
        map[row | i] = hm::bitrev((unsigned)((m2 - pos - 1) >> 1), (int)logN);
        pos = (pos * 3) & (m2 - 1);
    }
    slot_map.upload(map, stream);
    APSU_CUDA_CHECK(cudaStreamSynchronize(stream)); // host vectors go out of scope
}

void DeviceContext::build_levels()
{
    level.assign(first_L + 1, LevelConsts());
    ks.assign(first_L + 1, KeySwitchConsts());
    const uint64_t mt = 1ull << 32;
    for (uint32_t L = 1; L <= first_L; L++) {
        LevelConsts &c = level[L];
        std::memset(&c, 0, sizeof(c));
        std::vector<uint64_t> q(params.coeff_modulus, params.coeff_modulus + L);
        hm::Wide qw;
        for (uint64_t x : q) qw.mul(x);
        uint32_t nb = L + ((32 + hm::bit_length(t) + qw.bits() >= 61 * (int)L + 61) ? 1 : 0);
        std::vector<uint64_t> B(mod_values.begin() + idx_B0, mod_values.begin() + idx_B0 + nb);
        uint64_t msk = mod_values[idx_msk];
        std::vector<uint64_t> bsk = B;
        bsk.push_back(msk);
        c.L = (int)L;
        c.S = (int)bsk.size();
        hm::Wide q_div_t = qw.div(t);
        c.q_mod_t = qw.mod(t);
        for (uint32_t i = 0; i < L; i++) {
            c.q[i] = make_mod(q[i]);
            c.coeff_div_plain[i] = q_div_t.mod(q[i]);
            if (i + 1 < L) {
                c.inv_qlast[i] = make_shoup(invm(q[L - 1] % q[i], q[i]), q[i]);
                c.half_mod[i] = (q[L - 1] >> 1) % q[i];
                c.last_kind[i] = q[L - 1] <= q[i] ? 0u : (q[L - 1] <= 2 * q[i] ? 1u : 2u);
            }
            uint64_t punct = hm::prod_mod(q, q[i], (int)i);
            uint64_t inv_punct = invm(punct, q[i]);
            c.inv_punct_q[i] = make_shoup(inv_punct, q[i]);
            c.t_inv_punct_q[i] = make_shoup(mulm(t % q[i], inv_punct, q[i]), q[i]);
            c.mtilde_inv_punct_q[i] = make_shoup(mulm(mt % q[i], inv_punct, q[i]), q[i]);
            c.q_punct_mod_mtilde[i] = (u32)hm::prod_mod(q, mt, (int)i);
            c.t_mod_q[i] = make_shoup(t, q[i]);
            uint64_t Bq = hm::prod_mod(B, q[i]);
            c.B_mod_q[i] = make_shoup(Bq, q[i]);
            c.neg_B_mod_q[i] = make_shoup((q[i] - Bq) % q[i], q[i]);
            for (uint32_t k = 0; k < nb; k++) c.B_punct_mod_q[i][k] = make_shoup(hm::prod_mod(B, q[i], (int)k), q[i]);
        }
        c.neg_inv_q_mod_mtilde = (u32)((mt - invm(hm::prod_mod(q, mt), mt)) % mt);
        for (uint32_t j = 0; j < bsk.size(); j++) {
            uint64_t p = bsk[j];
            c.bsk[j] = make_mod(p);
            uint64_t qp = hm::prod_mod(q, p);
            const uint64_t inv_mt = invm(mt % p, p);
            // fast floor (step 7) constants carry q^-1 and, for the primes of B (j < nb), also the (B/B_j)^-1 of the
            // Shenoy-Kumaresan conversion that follows (step 8): the kernel then gets g_j = f_j * (B/B_j)^-1 directly
            uint64_t inv_q = invm(qp, p);
            if (j < nb) inv_q = mulm(inv_q, invm(hm::prod_mod(B, p, (int)j), p), p);
            else inv_q = mulm(inv_q, invm(hm::prod_mod(B, p), p), p); // m_sk: times B^-1, see B_punct_mod_msk below
            for (uint32_t i = 0; i < L; i++) {
                const uint64_t punct = hm::prod_mod(q, p, (int)i);
                // extension (steps 1-2): the m_tilde^-1 of the Montgomery reduction folded into the conversion
                c.ext_punct_bsk[j][i] = make_shoup(mulm(punct, inv_mt, p), p);
                // fast floor (step 7): -(q/q_i) * q^-1
                c.floor_punct_bsk[j][i] = make_shoup((p - mulm(punct, inv_q, p)) % p, p);
            }
            c.ext_q_bsk[j] = make_shoup(mulm(qp, inv_mt, p), p);
            c.floor_t_bsk[j] = make_shoup(mulm(t % p, inv_q, p), p);
        }
        for (uint32_t k = 0; k < nb; k++) {
            c.inv_punct_B[k] = make_shoup(invm(hm::prod_mod(B, B[k], (int)k), B[k]), B[k]);
            // alpha_sk = (sum_k g_k * (B/B_k) - f_msk) * B^-1 mod m_sk with the B^-1 folded into both sides
            c.B_punct_mod_msk[k] = make_shoup(mulm(hm::prod_mod(B, msk, (int)k), invm(hm::prod_mod(B, msk), msk), msk), msk);
        }
        c.inv_B_mod_msk = make_shoup(invm(hm::prod_mod(B, msk), msk), msk);

        KeySwitchConsts &kc = ks[L];
        std::memset(&kc, 0, sizeof(kc));
        kc.L = (int)L;
        kc.K = (int)K;
        if (K > 1) {
            uint64_t P = params.coeff_modulus[K - 1];
            for (uint32_t i = 0; i < L; i++) {
                kc.key_mod[i] = make_mod(q[i]);
                kc.inv_P[i] = make_shoup(invm(P % q[i], q[i]), q[i]);
                kc.half_P_mod[i] = (P >> 1) % q[i];
                kc.P_kind[i] = P <= q[i] ? 0u : (P <= 2 * q[i] ? 1u : 2u);
            }
            kc.key_mod[L] = make_mod(P);
            kc.half_P = P >> 1;
        }
    }
}

std::vector<uint32_t> DeviceContext::pattern_q(uint32_t L) const
{
    std::vector<uint32_t> p(L);
    for (uint32_t i = 0; i < L; i++) p[i] = i;
    return p;
}
std::vector<uint32_t> DeviceContext::pattern_bsk(uint32_t L) const
{
    std::vector<uint32_t> p;
    for (int k = 0; k + 1 < level[L].S; k++) p.push_back(idx_B0 + (uint32_t)k);
    p.push_back(idx_msk);
    return p;
}
std::vector<uint32_t> DeviceContext::pattern_ext(uint32_t L) const
{
    std::vector<uint32_t> p = pattern_q(L), b = pattern_bsk(L);
    p.insert(p.end(), b.begin(), b.end());
    return p;
}
std::vector<uint32_t> DeviceContext::pattern_ks(uint32_t L) const
{
    std::vector<uint32_t> p = pattern_q(L);
    p.push_back(K - 1);
    return p;
}

NttArgs DeviceContext::make_args(const std::vector<uint32_t> &pattern) const
{
    if (pattern.empty() || pattern.size() > (size_t)kMaxPattern) throw std::invalid_argument("NTT modulus pattern length is invalid");
    NttArgs a;
    std::memset(&a, 0, sizeof(a));
    a.tw = twiddles.p;
    a.pattern_len = (int)pattern.size();
    for (size_t i = 0; i < pattern.size(); i++) {
        if (pattern[i] >= mod_values.size()) throw std::invalid_argument("NTT modulus index is out of range");
        a.mod[i] = mod_host[pattern[i]];
        a.inv_n[i] = inv_n_host[pattern[i]];
        a.inv_n_w[i] = inv_n_w_host[pattern[i]];
        a.table[i] = (int)pattern[i];
    }
    return a;
}

template <int LOGN, int DIV, int MODE, bool TWS = false>
static void launch_ntt_shape(const u64 *in, u64 *out, uint32_t count, const NttArgs &a, const NttSrc &s, const NttFuse &f, bool inverse, cudaStream_t st)
{
    constexpr int threads = (1 << LOGN) / DIV;
    // polynomial + one pad word per 16 (+ the modulus' twiddle table, 16 bytes per entry, when staged: ntt.cuh TWS)
    constexpr size_t smem = (sizeof(u64) << LOGN) + (sizeof(u64) << (LOGN - 4)) + (TWS ? (size_t)16 << LOGN : 0);
    // the forward transform fuses kNttExtend, the inverse one kNttTensor / kNttKsMac (ntt.cuh)
    if (inverse) {
        if constexpr (MODE != kNttExtend) launch_pdl(ntt_kernel<LOGN, false, DIV, MODE, TWS>, dim3(count), dim3(threads), smem, st, in, out, a, s, f);
    } else {
        if constexpr (MODE == kNttPlain || MODE == kNttExtend) launch_pdl(ntt_kernel<LOGN, true, DIV, MODE, TWS>, dim3(count), dim3(threads), smem, st, in, out, a, s, f);
    }
}

// Function attributes are per device: every DeviceContext opts its device's kernel instances into the large
// dynamic shared-memory size once, in its constructor (two contexts on different GPUs of one process both work).
template <int LOGN, int DIV>
static void configure_ntt_shape()
{
    constexpr size_t smem = (sizeof(u64) << LOGN) + (sizeof(u64) << (LOGN - 4));
    APSU_CUDA_CHECK(cudaFuncSetAttribute(ntt_kernel<LOGN, true, DIV, kNttPlain>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    APSU_CUDA_CHECK(cudaFuncSetAttribute(ntt_kernel<LOGN, true, DIV, kNttExtend>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    APSU_CUDA_CHECK(cudaFuncSetAttribute(ntt_kernel<LOGN, false, DIV, kNttPlain>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    APSU_CUDA_CHECK(cudaFuncSetAttribute(ntt_kernel<LOGN, false, DIV, kNttTensor>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    APSU_CUDA_CHECK(cudaFuncSetAttribute(ntt_kernel<LOGN, false, DIV, kNttKsMac>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}
// split transforms (ntt.cuh: ntt_split_kernel): 2^LC CTAs (one cluster) per polynomial
template <int LOGN, int LC>
static void launch_ntt_split(const u64 *in, u64 *out, uint32_t count, const NttArgs &a, const NttSrc &s, bool inverse, cudaStream_t st)
{
    constexpr int LOGM = LOGN - LC, DIV = 8, threads = (1 << LOGM) / DIV;
    constexpr size_t smem = (sizeof(u64) << LOGM) + (sizeof(u64) << (LOGM - 4));
    if (inverse)
        launch_pdl(ntt_split_kernel<LOGN, LC, false, DIV>, dim3(count << LC), dim3(threads), smem, st, in, out, a, s);
    else
        launch_pdl(ntt_split_kernel<LOGN, LC, true, DIV>, dim3(count << LC), dim3(threads), smem, st, in, out, a, s);
}
template <int LOGN, int LC>
static void configure_ntt_split()
{
    constexpr int LOGM = LOGN - LC;
    constexpr size_t smem = (sizeof(u64) << LOGM) + (sizeof(u64) << (LOGM - 4));
    APSU_CUDA_CHECK(cudaFuncSetAttribute(ntt_split_kernel<LOGN, LC, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    APSU_CUDA_CHECK(cudaFuncSetAttribute(ntt_split_kernel<LOGN, LC, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}
// APSU_B200_NTT_SPLIT: unset = by batch size (launch_ntt); 0 = never; 2 / 4 = that many CTAs per polynomial whatever
// the batch size (A/B runs, tools/bench_ntt.py)
static int ntt_split_mode()
{
    static const int mode = [] {
        const char *ev = std::getenv("APSU_B200_NTT_SPLIT");
        return ev ? atoi(ev) : -1;
    }();
    return mode;
}

constexpr bool ntt_tws_fits(int logn) { return logn <= 13; } // 69.6 + 128 KB at N = 8192
template <int LOGN>
static void configure_ntt()
{
    constexpr int kLatDiv = (LOGN == 12 || LOGN == 13) ? 8 : 16;
    configure_ntt_shape<LOGN, kLatDiv>();
    configure_ntt_shape<LOGN, 32>();
    configure_ntt_split<LOGN, 1>();
    configure_ntt_split<LOGN, 2>();
    if constexpr (ntt_tws_fits(LOGN)) {
        constexpr size_t smem = (sizeof(u64) << LOGN) + (sizeof(u64) << (LOGN - 4)) + ((size_t)16 << LOGN);
        APSU_CUDA_CHECK(cudaFuncSetAttribute(ntt_kernel<LOGN, true, kLatDiv, kNttPlain, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        APSU_CUDA_CHECK(cudaFuncSetAttribute(ntt_kernel<LOGN, false, kLatDiv, kNttPlain, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
}

bool pdl_enabled()
{
    // measured on B200 with the launch sequences replayed as CUDA graphs (profiles/bench_r02_pdl_ab.json): 16M-4096
    // 5.91 -> 5.93 ms, 1M-4096-com 1.66 -> 1.70, 256K-512 0.213 -> 0.227, 1M-1024-cmp 0.729 -> 0.689: a graph's
    // kernel-to-kernel edges leave nothing for the programmatic edge to hide, so it is off by default
    // (APSU_B200_PDL=1 for A/B runs; results are bit-identical either way)
    static const bool on = [] {
        const char *ev = std::getenv("APSU_B200_PDL");
        return ev ? atoi(ev) != 0 : false;
    }();
    return on;
}

static int ntt_tws_mode()
{
    static const int mode = [] {
        // measured on B200 (tools/bench_ntt.py, 24..273 polynomials): no gain (N = 8192, 148 polynomials: 16.6 us either
        // way; N = 4096, 84: 10.0 vs 10.4 us) — a lone transform is bound by its own multiplier instructions (ntt.cuh),
        // not by twiddle latency; kept for A/B runs, off by default
        const char *ev = std::getenv("APSU_B200_NTT_TWS");
        return ev ? atoi(ev) : 0;
    }();
    return mode;
}

// picks the launch shape by batch size (ntt.cuh): more threads per polynomial while the batch leaves SMs idle
template <int LOGN, int MODE>
static void launch_ntt(const u64 *in, u64 *out, uint32_t count, const NttArgs &a, const NttSrc &s, const NttFuse &f, bool inverse, cudaStream_t st, int sms)
{
    // measured (N = 8192, one bundle index): 24..112 polynomials take 16 us with N/8 threads and 20-22 us with N/32;
    // N/16 threads for 150..300 polynomials made no difference
    constexpr int kLatDiv = (LOGN == 12 || LOGN == 13) ? 8 : 16; // at most 1024 threads per CTA
    if constexpr (MODE == kNttPlain) {
        const int split = ntt_split_mode();
        // measured (tools/bench_ntt.py, profiles/ntt_split_r02.jsonl; N = 8192 forward, 24 / 42 / 84 polynomials: 16.4 us
        // unsplit, 10.3 / 10.3 / 16.4 us two ways, 8.3 / 10.3 / 14.4 us four ways): a slice costs its share of ONE SM's
        // multiplier time, so splitting pays only while the slices still find idle SMs.  The inverse is never split by
        // default: its stride-C stores are partial-sector writes (15.5 -> 14.4 us at 24 polynomials, slower from 42 on).
        const bool four = split == 4 || (split < 0 && !inverse && LOGN >= 13 && (uint64_t)count * 4 <= (uint64_t)sms);
        const bool two = split == 2 || (split < 0 && !inverse && (uint64_t)count * 2 <= (uint64_t)sms);
        if (four) {
            launch_ntt_split<LOGN, 2>(in, out, count, a, s, inverse, st);
            return;
        }
        if (two) {
            launch_ntt_split<LOGN, 1>(in, out, count, a, s, inverse, st);
            return;
        }
    }
    if (count <= (uint32_t)sms * ntt_min_blocks(LOGN, kLatDiv)) {
        if constexpr (MODE == kNttPlain && ntt_tws_fits(LOGN)) {
            // a batch of at most one CTA per SM: twiddles staged in shared memory (one CTA per SM at N = 8192)
            if (ntt_tws_mode() && count <= (uint32_t)sms * (LOGN == 13 ? 1 : 2)) {
                launch_ntt_shape<LOGN, kLatDiv, MODE, true>(in, out, count, a, s, f, inverse, st);
                return;
            }
        }
        launch_ntt_shape<LOGN, kLatDiv, MODE>(in, out, count, a, s, f, inverse, st);
    } else
        launch_ntt_shape<LOGN, 32, MODE>(in, out, count, a, s, f, inverse, st);
}

template <int MODE>
static void launch_ntt_logn(uint32_t logN, const u64 *in, u64 *out, uint32_t count, const NttArgs &a, const NttSrc &s, const NttFuse &f, bool inverse, cudaStream_t st,
                            int sms)
{
    switch (logN) {
    case 11: launch_ntt<11, MODE>(in, out, count, a, s, f, inverse, st, sms); break;
    case 12: launch_ntt<12, MODE>(in, out, count, a, s, f, inverse, st, sms); break;
    case 13: launch_ntt<13, MODE>(in, out, count, a, s, f, inverse, st, sms); break;
    case 14: launch_ntt<14, MODE>(in, out, count, a, s, f, inverse, st, sms); break;
    default: throw std::invalid_argument("unsupported poly_modulus_degree");
    }
}

void DeviceContext::ntt(const u64 *in, u64 *out, uint32_t count, const std::vector<uint32_t> &pattern, bool inverse,
                        const uint32_t *src_idx, const uint32_t *dst_idx, bool reduce_input)
{
    if (!count) return;
    NttArgs a = make_args(pattern);
    NttSrc s{ src_idx, dst_idx, reduce_input ? 1 : 0 };
    launch_ntt_logn<kNttPlain>(logN, in, out, count, a, s, NttFuse(), inverse, stream, sms);
    APSU_CUDA_CHECK(cudaGetLastError());
    launches++;
}

// transforms with a fused element-wise prologue (ntt.cuh: NttFuse); `arena` is both the source of the fused inputs and
// the destination
void DeviceContext::ntt_fused(int mode, u64 *arena, uint32_t count, const std::vector<uint32_t> &pattern, const uint32_t *src_idx, const uint32_t *dst_idx,
                              u64 *out_base, const NttFuse &f)
{
    if (!count) return;
    NttArgs a = make_args(pattern);
    NttSrc s{ src_idx, dst_idx, 0 };
    switch (mode) {
    case kNttExtend: launch_ntt_logn<kNttExtend>(logN, arena, out_base, count, a, s, f, false, stream, sms); break;
    case kNttTensor: launch_ntt_logn<kNttTensor>(logN, arena, out_base, count, a, s, f, true, stream, sms); break;
    case kNttKsMac: launch_ntt_logn<kNttKsMac>(logN, arena, out_base, count, a, s, f, true, stream, sms); break;
    default: throw std::invalid_argument("unknown fused transform");
    }
    APSU_CUDA_CHECK(cudaGetLastError());
    launches++;
}

} // namespace apsu_b200
