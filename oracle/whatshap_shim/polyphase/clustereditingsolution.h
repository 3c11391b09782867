// TEST INFRASTRUCTURE.  Stand-in for WhatsHap ClusterEditingSolution (reference
// src/alignmentstoreadset.cpp:313-315,333,605-607,662,674).
#pragma once
#include <vector>
#include "staticsparsegraph.h"
class ClusterEditingSolution {
public:
    ClusterEditingSolution() {}
    explicit ClusterEditingSolution(const std::vector<std::vector<StaticSparseGraph::NodeId>>& c) : c_(c) {}
    unsigned int getNumClusters() const { return (unsigned)c_.size(); }
    const std::vector<StaticSparseGraph::NodeId>& getCluster(unsigned int i) const { return c_.at(i); }
private:
    std::vector<std::vector<StaticSparseGraph::NodeId>> c_;
};
