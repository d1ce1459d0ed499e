// C++ facade test.  `test_facade host <json>`: parameter loading, PowersDag, error mapping (no GPU needed).
// `test_facade gpu <json> <ncoeffs> <seed>`: synthetic DB + synthetic query through Receiver::RunQuery; prints
// an FNV-1a checksum per BinBundle result so that the Python test can compare it with the C-ABI path.
// `test_facade mgpu <json> <ncoeffs> <seed> <world>`: the same DB spread over `world` GPUs (bundle index b on GPU
// b % world) through MultiGpuReceiver (apsu_b200_mgpu_*, NCCL): must print the same lines.
#include "../../apsu_b200/host/apsu_b200.hpp"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

static std::string slurp(const char *path)
{
    std::ifstream f(path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
static uint64_t splitmix(uint64_t &s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint64_t fnv(const uint64_t *d, size_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) {
        h ^= d[i];
        h *= 1099511628211ull;
    }
    return h;
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    std::string mode = argv[1];
    apsu::PSUParams params = apsu::PSUParams::Load(slurp(argv[2]));
    if (mode == "host") {
        apsu::PowersDag pd;
        if (!pd.configure(params)) return 1;
        std::printf("N=%zu t=%llu K=%zu bundle_idx_count=%u items_per_bundle=%u depth=%u sources=%u targets=%zu\n",
                    params.seal_params().poly_modulus_degree, (unsigned long long)params.seal_params().plain_modulus,
                    params.seal_params().coeff_modulus.size(), params.bundle_idx_count(), params.items_per_bundle(), pd.depth(),
                    pd.source_count(), pd.target_powers().size());
        // the reference's exception types
        int caught = 0;
        try {
            apsu::PSUParams::Load("{\"table_params\": {}}");
        } catch (const std::runtime_error &) {
            caught |= 1;
        }
        try {
            auto tp = params.table_params();
            tp.table_size += 1; // not a multiple of items_per_bundle
            apsu::PSUParams bad(params.item_params(), tp, params.query_params(), params.seal_params());
        } catch (const std::invalid_argument &) {
            caught |= 2;
        }
        try {
            apsu::PowersDag unconfigured;
            unconfigured.depth();
        } catch (const std::logic_error &) {
            caught |= 4;
        }
        std::printf("exceptions=%d\n", caught);
        return caught == 7 ? 0 : 1;
    }
    // gpu / mgpu mode
    uint32_t ncoeffs = (uint32_t)std::atoi(argv[3]);
    uint64_t seed = std::strtoull(argv[4], nullptr, 10);
    const size_t world = mode == "mgpu" ? (size_t)std::atoi(argv[5]) : 1;
    std::unique_ptr<apsu::receiver::MultiGpuReceiver> mg;
    std::shared_ptr<apsu::receiver::ReceiverDB> db;
    if (mode == "mgpu") {
        std::vector<int> devices;
        for (size_t r = 0; r < world; r++) devices.push_back((int)r);
        mg = std::make_unique<apsu::receiver::MultiGpuReceiver>(params, devices);
        for (uint32_t b = 0; b < params.bundle_idx_count(); b++) mg->add_bin_bundle_synthetic(b % world, b, ncoeffs, seed + b);
        mg->commit();
        db = mg->db(0);
    } else {
        db = std::make_shared<apsu::receiver::ReceiverDB>(params, 0);
        for (uint32_t b = 0; b < params.bundle_idx_count(); b++) db->add_bin_bundle_synthetic(b, ncoeffs, seed + b);
    }
    auto sp = params.seal_params();
    size_t N = sp.poly_modulus_degree, K = sp.coeff_modulus.size(), L = K > 1 ? K - 1 : 1;
    uint64_t s = seed;
    std::unordered_map<uint32_t, std::vector<std::vector<uint64_t>>> data;
    for (uint32_t e : params.query_params().query_powers) {
        auto &row = data[e];
        for (uint32_t b = 0; b < params.bundle_idx_count(); b++) {
            std::vector<uint64_t> ct(2 * L * N);
            for (size_t c = 0; c < 2; c++)
                for (size_t j = 0; j < L; j++)
                    for (size_t n = 0; n < N; n++) ct[(c * L + j) * N + n] = splitmix(s) % sp.coeff_modulus[j];
            row.push_back(std::move(ct));
        }
    }
    std::vector<uint64_t> relin;
    if (K > 1) {
        relin.resize((K - 1) * 2 * K * N);
        for (size_t i = 0; i < relin.size(); i++) relin[i] = splitmix(s) % sp.coeff_modulus[(i / N) % K];
    }
    std::vector<uint64_t> masks(params.bundle_idx_count() * N);
    for (auto &m : masks) m = splitmix(s) % sp.plain_modulus;
    apsu::receiver::Query query(db, std::move(data), std::move(relin));
    if (!query) return 1;
    auto print = [&](apsu::receiver::ResultPart rp) {
        std::printf("bundle_idx=%u cache_idx=%u fnv=%016llx\n", rp->bundle_idx, rp->cache_idx,
                    (unsigned long long)fnv(rp->psu_result.data(), rp->psu_result.size()));
    };
    if (mg)
        mg->RunQuery(query, std::vector<std::vector<uint64_t>>(world, masks), print); // every bundle index has one BinBundle: local cache 0 everywhere
    else
        apsu::receiver::Receiver::RunQuery(query, masks, print);
    return 0;
}
