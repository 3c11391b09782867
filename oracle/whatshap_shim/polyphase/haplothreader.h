// TEST INFRASTRUCTURE.  Stand-in for WhatsHap HaploThreader (reference
// src/alignmentstoreadset.cpp:320-329,408).  Algorithm: oracle/core/phase_core.hpp rule R3.
// PARITY UNPINNED.
#pragma once
#include <unordered_map>
#include "../../core/phase_core.hpp"
typedef uint32_t Position;
typedef uint32_t GlobalClusterId;
typedef uint32_t LocalClusterId;
class HaploThreader {
public:
    HaploThreader(uint32_t ploidy, double switchCost = 32.0, double affineSwitchCost = 8.0,
                  bool symmetryOptimization = false, uint32_t rowLimit = 0)
        : ploidy_(ploidy), sc_(switchCost), asc_(affineSwitchCost) { (void)symmetryOptimization; (void)rowLimit; }
    std::vector<std::vector<GlobalClusterId>> computePaths(
        Position start, Position end, const std::vector<std::vector<GlobalClusterId>>& covMap,
        const std::vector<std::vector<double>>& coverage, const std::vector<std::vector<uint32_t>>& consensus,
        const std::vector<std::unordered_map<uint32_t, uint32_t>>& genotypes, Position displayedEnd = 0) const {
        (void)displayedEnd;
        ahs_oracle::ThreadResult r = ahs_oracle::thread_paths(ploidy_, sc_, asc_, start, end, covMap, coverage, consensus, genotypes);
        last_cost = r.cost;
        return r.path;
    }
    mutable double last_cost = 0.0;
private:
    uint32_t ploidy_; double sc_, asc_;
};
