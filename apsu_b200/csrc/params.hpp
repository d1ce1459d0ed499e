// PSUParams (JSON loader + invariants), create_powers_set and PowersDag for the B200 receiver.
// Reference: common/apsu/psu_params.cpp:95-180,290-374; common/apsu/util/utils.cpp:146-177;
// common/apsu/powers.cpp:22-107.
#pragma once
#include "../../include/apsu_b200.h"
#include <cstdint>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace apsu_b200 {

// status-carrying exceptions are the std ones; the C ABI maps them to apsu_b200_status.

void params_load_json(const std::string &text, apsu_b200_params &out); // throws
void params_validate(apsu_b200_params &p);                             // throws
std::vector<uint64_t> coeff_modulus_create(uint32_t N, const std::vector<int> &bit_sizes);
uint64_t plain_modulus_batching(uint32_t N, int bits);

std::set<uint32_t> create_powers_set(uint32_t ps_low_degree, uint32_t target_degree);

struct PowersNode {
    uint32_t power = 0, depth = 0, parent1 = 0, parent2 = 0;
    bool is_source() const { return !parent1 && !parent2; }
};

class PowersDag {
public:
    bool configure(const std::set<uint32_t> &source_powers, const std::set<uint32_t> &target_powers);
    bool is_configured() const { return configured_; }
    uint32_t depth() const { need(); return depth_; }
    uint32_t source_count() const { need(); return source_count_; }
    const std::set<uint32_t> &target_powers() const { need(); return targets_; }
    const PowersNode &node(uint32_t power) const { return nodes_.at(power); }
    // nodes grouped by depth (depth 0 = sources); every node's parents live in strictly lower groups,
    // so a group is one batch of independent ciphertext products.
    std::vector<std::vector<PowersNode>> levels() const;

private:
    void need() const
    {
        if (!configured_) throw std::logic_error("PowersDag has not been configured");
    }
    std::map<uint32_t, PowersNode> nodes_;
    std::set<uint32_t> targets_;
    uint32_t depth_ = 0, source_count_ = 0;
    bool configured_ = false;
};

} // namespace apsu_b200
