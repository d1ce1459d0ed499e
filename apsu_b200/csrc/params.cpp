#include "params.hpp"
#include "hostmath.hpp"
#include <algorithm>
#include <cctype>
#include <cstring>
#include <memory>

namespace apsu_b200 {

// ------------------------------------------------------------------------------------------------
// Minimal JSON reader (objects, arrays, integers, strings, true/false/null) — jsoncpp is what the
// reference uses (psu_params.cpp:20-93); only the subset the parameter files need is implemented.
// ------------------------------------------------------------------------------------------------
namespace {
struct JValue {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    bool negative = false, integral = true;
    uint64_t num = 0;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;
    const JValue *find(const std::string &k) const
    {
        for (auto &kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

class JParser {
public:
    explicit JParser(const std::string &s) : s_(s) {}
    JValue parse()
    {
        JValue v = value();
        ws();
        if (i_ != s_.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string &s_;
    size_t i_ = 0;
    [[noreturn]] void fail(const char *what) const
    {
        throw std::runtime_error(std::string("JSON parse error at offset ") + std::to_string(i_) + ": " + what);
    }
    void ws()
    {
        while (i_ < s_.size() && std::isspace((unsigned char)s_[i_])) i_++;
    }
    bool eat(char c)
    {
        ws();
        if (i_ < s_.size() && s_[i_] == c) {
            i_++;
            return true;
        }
        return false;
    }
    JValue value()
    {
        ws();
        if (i_ >= s_.size()) fail("unexpected end");
        char c = s_[i_];
        JValue v;
        if (c == '{') {
            i_++;
            v.kind = JValue::Obj;
            if (eat('}')) return v;
            do {
                ws();
                JValue k = string_();
                if (!eat(':')) fail("expected ':'");
                v.obj.emplace_back(k.str, value());
            } while (eat(','));
            if (!eat('}')) fail("expected '}'");
        } else if (c == '[') {
            i_++;
            v.kind = JValue::Arr;
            if (eat(']')) return v;
            do v.arr.push_back(value());
            while (eat(','));
            if (!eat(']')) fail("expected ']'");
        } else if (c == '"') {
            v = string_();
        } else if (c == '-' || std::isdigit((unsigned char)c)) {
            v.kind = JValue::Num;
            if (c == '-') v.negative = true, i_++;
            if (i_ >= s_.size() || !std::isdigit((unsigned char)s_[i_])) fail("bad number");
            while (i_ < s_.size() && std::isdigit((unsigned char)s_[i_])) v.num = v.num * 10 + (uint64_t)(s_[i_++] - '0');
            if (i_ < s_.size() && (s_[i_] == '.' || s_[i_] == 'e' || s_[i_] == 'E')) {
                v.integral = false;
                while (i_ < s_.size() && (std::isdigit((unsigned char)s_[i_]) || std::strchr(".eE+-", s_[i_]))) i_++;
            }
        } else if (!s_.compare(i_, 4, "true")) {
            v.kind = JValue::Bool, v.b = true, i_ += 4;
        } else if (!s_.compare(i_, 5, "false")) {
            v.kind = JValue::Bool, i_ += 5;
        } else if (!s_.compare(i_, 4, "null")) {
            i_ += 4;
        } else {
            fail("unexpected character");
        }
        return v;
    }
    JValue string_()
    {
        if (i_ >= s_.size() || s_[i_] != '"') fail("expected string");
        i_++;
        JValue v;
        v.kind = JValue::Str;
        while (i_ < s_.size() && s_[i_] != '"') {
            if (s_[i_] == '\\' && i_ + 1 < s_.size()) i_++;
            v.str.push_back(s_[i_++]);
        }
        if (i_ >= s_.size()) fail("unterminated string");
        i_++;
        return v;
    }
};

const JValue &member(const JValue &parent, const char *name)
{
    const JValue *v = parent.kind == JValue::Obj ? parent.find(name) : nullptr;
    if (!v || v->kind == JValue::Null) throw std::runtime_error(std::string("JSON is missing required member: ") + name);
    return *v;
}
uint64_t as_u64(const JValue &v, const char *what)
{
    if (v.kind != JValue::Num || v.negative || !v.integral) throw std::runtime_error(std::string(what) + " should be an unsigned integer");
    return v.num;
}
uint32_t as_u32(const JValue &v, const char *what)
{
    uint64_t x = as_u64(v, what);
    if (x > 0xFFFFFFFFull) throw std::runtime_error(std::string(what) + " is out of range");
    return (uint32_t)x;
}
} // namespace

std::vector<uint64_t> coeff_modulus_create(uint32_t N, const std::vector<int> &bit_sizes)
{
    if (N < 2 || (N & (N - 1))) throw std::invalid_argument("poly_modulus_degree is invalid");
    if (bit_sizes.size() > APSU_B200_MAX_COEFF_MODULUS) throw std::invalid_argument("bit_sizes is invalid");
    // one descending candidate list per distinct size; equal sizes are handed out smallest-first
    std::map<int, std::vector<uint64_t>> pool;
    for (int b : bit_sizes)
        if (!pool.count(b)) pool[b] = hm::primes_below_pow2(2ull * N, b, (size_t)std::count(bit_sizes.begin(), bit_sizes.end(), b));
    std::vector<uint64_t> out;
    for (int b : bit_sizes) {
        out.push_back(pool[b].back());
        pool[b].pop_back();
    }
    return out;
}

uint64_t plain_modulus_batching(uint32_t N, int bits) { return hm::primes_below_pow2(2ull * N, bits, 1)[0]; }

void params_load_json(const std::string &text, apsu_b200_params &out)
{
    JValue root = JParser(text).parse();
    std::memset(&out, 0, sizeof(out));

    const JValue &tp = member(root, "table_params");
    out.hash_func_count = as_u32(member(tp, "hash_func_count"), "hash_func_count");
    out.table_size = as_u32(member(tp, "table_size"), "table_size");
    out.max_items_per_bin = as_u32(member(tp, "max_items_per_bin"), "max_items_per_bin");

    out.felts_per_item = as_u32(member(member(root, "item_params"), "felts_per_item"), "felts_per_item");

    const JValue &qp = member(root, "query_params");
    out.ps_low_degree = as_u32(member(qp, "ps_low_degree"), "ps_low_degree");
    std::set<uint32_t> powers{ 1 }; // "Should always contain 1" (psu_params.cpp:328)
    const JValue &qpow = member(qp, "query_powers");
    if (qpow.kind != JValue::Arr) throw std::runtime_error("query_powers should be an array");
    for (auto &v : qpow.arr) powers.insert(as_u32(v, "query_powers element"));
    if (powers.size() > APSU_B200_MAX_QUERY_POWERS) throw std::runtime_error("too many query_powers");
    for (uint32_t p : powers) out.query_powers[out.query_power_count++] = p;

    const JValue &sp = member(root, "seal_params");
    const JValue &bits = member(sp, "coeff_modulus_bits");
    uint64_t N = as_u64(member(sp, "poly_modulus_degree"), "poly_modulus_degree");
    if (N > (1u << 17)) throw std::runtime_error("poly_modulus_degree is out of range");
    out.poly_modulus_degree = (uint32_t)N;
    const JValue *pm = sp.find("plain_modulus"), *pmb = sp.find("plain_modulus_bits");
    if (pm && pmb) throw std::runtime_error("only one of plain_modulus and plain_modulus_bits must be specified");
    if (pm)
        out.plain_modulus = as_u64(*pm, "plain_modulus");
    else if (pmb)
        out.plain_modulus = plain_modulus_batching(out.poly_modulus_degree, (int)as_u32(*pmb, "plain_modulus_bits"));
    else
        throw std::runtime_error("neither plain_modulus nor plain_modulus_bits was specified");
    if (bits.kind != JValue::Arr) throw std::runtime_error("coeff_modulus_bits should be an array");
    std::vector<int> sizes;
    for (auto &v : bits.arr) sizes.push_back((int)as_u32(v, "coeff_modulus_bits element"));
    auto primes = coeff_modulus_create(out.poly_modulus_degree, sizes);
    out.coeff_modulus_count = (uint32_t)primes.size();
    for (size_t i = 0; i < primes.size(); i++) out.coeff_modulus[i] = primes[i];

    params_validate(out);
}

void params_validate(apsu_b200_params &p)
{
    if (!p.table_size) throw std::invalid_argument("table_size cannot be zero");
    if (!p.max_items_per_bin) throw std::invalid_argument("max_items_per_bin cannot be zero");
    if (p.hash_func_count < 1 || p.hash_func_count > 8) throw std::invalid_argument("hash_func_count is too large or too small");
    if (p.felts_per_item < 2 || p.felts_per_item > 32) throw std::invalid_argument("felts_per_item is too large or too small");
    if (p.ps_low_degree > p.max_items_per_bin) throw std::invalid_argument("ps_low_degree cannot be larger than max_items_per_bin");
    if (p.query_power_count > APSU_B200_MAX_QUERY_POWERS) throw std::invalid_argument("too many query_powers");
    std::set<uint32_t> qp(p.query_powers, p.query_powers + p.query_power_count);
    if (qp.count(0) || !qp.count(1)) throw std::invalid_argument("query_powers cannot contain 0 and must contain 1");
    if (qp.size() > p.max_items_per_bin) throw std::invalid_argument("query_powers cannot be larger than max_items_per_bin");
    for (uint32_t q : qp) {
        if (q > p.max_items_per_bin) throw std::invalid_argument("query_powers cannot contain values larger than max_items_per_bin");
        if (q > p.ps_low_degree && q % (p.ps_low_degree + 1))
            throw std::invalid_argument("query_powers cannot contain values larger than ps_low_degree that are not multiples ps_low_degree + 1");
    }
    // keep the canonical (sorted, unique) form
    p.query_power_count = 0;
    for (uint32_t q : qp) p.query_powers[p.query_power_count++] = q;

    // what SEALContext would reject (seal/context.cpp) for the parameter shapes the path supports
    uint32_t N = p.poly_modulus_degree;
    if (N < 1024 || N > 32768 || (N & (N - 1))) throw std::invalid_argument("Microsoft SEAL parameters are invalid: poly_modulus_degree");
    if (!p.coeff_modulus_count || p.coeff_modulus_count > APSU_B200_MAX_COEFF_MODULUS)
        throw std::invalid_argument("Microsoft SEAL parameters are invalid: coeff_modulus size");
    for (uint32_t i = 0; i < p.coeff_modulus_count; i++) {
        uint64_t q = p.coeff_modulus[i];
        if (hm::bit_length(q) > 60 || q < 2 || (q - 1) % (2ull * N) || !hm::miller_rabin(q))
            throw std::invalid_argument("Microsoft SEAL parameters are invalid: coeff_modulus primes must be NTT-friendly primes of at most 60 bits");
        for (uint32_t j = 0; j < i; j++)
            if (p.coeff_modulus[j] == q) throw std::invalid_argument("Microsoft SEAL parameters are invalid: coeff_modulus primes must be distinct");
        if (q <= p.plain_modulus) throw std::invalid_argument("plain_modulus must be smaller than every coeff_modulus prime");
    }
    if (p.plain_modulus < 2 || (p.plain_modulus - 1) % (2ull * N) || !hm::miller_rabin(p.plain_modulus))
        throw std::invalid_argument(
            "Microsoft SEAL parameters do not support batching; plain_modulus must be a prime congruent to 1 modulo 2*poly_modulus_degree");

    p.item_bit_count_per_felt = (uint32_t)hm::bit_length(p.plain_modulus) - 1;
    p.item_bit_count = p.item_bit_count_per_felt * p.felts_per_item;
    if (p.item_bit_count < 80 || p.item_bit_count > 128) throw std::invalid_argument("parameters result in too large or too small item_bit_count");
    p.items_per_bundle = N / p.felts_per_item;
    if (!p.items_per_bundle) throw std::invalid_argument("poly_modulus_degree is too small");
    p.bins_per_bundle = p.items_per_bundle * p.felts_per_item;
    if (p.table_size % p.items_per_bundle) throw std::invalid_argument("table_size must be a multiple of floor(poly_modulus_degree / felts_per_item)");
    p.bundle_idx_count = p.table_size / p.items_per_bundle;
}

std::set<uint32_t> create_powers_set(uint32_t ps_low_degree, uint32_t target_degree)
{
    if (ps_low_degree > target_degree) throw std::invalid_argument("ps_low_degree cannot be bigger than target_degree");
    if (!target_degree) throw std::invalid_argument("target_degree cannot be zero");
    std::set<uint32_t> s;
    uint32_t dense_upto = ps_low_degree ? ps_low_degree : target_degree;
    for (uint32_t e = 1; e <= dense_upto; e++) s.insert(e);
    if (ps_low_degree) // Paterson-Stockmeyer: the multiples of ps_low_degree+1 that fit
        for (uint32_t step = ps_low_degree + 1, e = step; e <= target_degree; e += step) s.insert(e);
    return s;
}

bool PowersDag::configure(const std::set<uint32_t> &sources, const std::set<uint32_t> &targets)
{
    nodes_.clear();
    targets_.clear();
    configured_ = false;
    depth_ = source_count_ = 0;
    auto bad = [](const std::set<uint32_t> &s) { return s.count(0) || !s.count(1); };
    if (bad(sources) || bad(targets)) return false;
    if (!std::includes(targets.begin(), targets.end(), sources.begin(), sources.end())) return false;

    for (uint32_t s : sources) nodes_[s] = PowersNode{ s, 0, 0, 0 };
    for (uint32_t e : targets) {
        if (nodes_.count(e)) continue;
        // best split e = a + b over target powers: minimal depth, first (smallest a) wins ties;
        // the fallback (e-1, 1) with depth e-1 is what the reference starts from.
        PowersNode best{ e, e - 1, e - 1, 1 };
        for (uint32_t a : targets) {
            if (a >= e) break;
            uint32_t b = e - a;
            if (!targets.count(b)) continue;
            uint32_t d = std::max(nodes_.at(a).depth, nodes_.at(b).depth) + 1;
            if (d < best.depth) best = PowersNode{ e, d, a, b };
        }
        nodes_[e] = best;
        depth_ = std::max(depth_, best.depth);
    }
    targets_ = targets;
    source_count_ = (uint32_t)sources.size();
    configured_ = true;
    return true;
}

std::vector<std::vector<PowersNode>> PowersDag::levels() const
{
    need();
    std::vector<std::vector<PowersNode>> lv(depth_ + 1);
    for (auto &kv : nodes_) lv[kv.second.depth].push_back(kv.second);
    return lv;
}

} // namespace apsu_b200
