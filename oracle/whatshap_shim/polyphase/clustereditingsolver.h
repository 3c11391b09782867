// TEST INFRASTRUCTURE.  Stand-in for WhatsHap ClusterEditingSolver (reference
// src/alignmentstoreadset.cpp:312-314).  Algorithm: oracle/core/phase_core.hpp rule R2.
// Nodes = all reads of the scored ReadSet (count handed over by the shim's ReadScoring, see
// readscoring.h), so reads without any partner form singleton clusters.
// PARITY UNPINNED.
#pragma once
#include "../../core/phase_core.hpp"
#include "clustereditingsolution.h"
#include "readscoring.h"
#include "trianglesparsematrix.h"
class ClusterEditingSolver {
public:
    ClusterEditingSolver(TriangleSparseMatrix& m, bool bundleEdges = false) : m_(m) { (void)bundleEdges; }
    ClusterEditingSolution run() {
        uint32_t n = std::max(m_.getMaxDim(), ahs_shim_num_reads());
        std::vector<ahs_oracle::PairScore> ps;
        for (auto& e : m_.getEntries()) {
            ahs_oracle::PairScore s; s.i = (int32_t)e.first; s.j = (int32_t)e.second; s.n = 0; s.k = 0;
            s.w = (int32_t)lrintf(m_.get(e.first, e.second) * 1024.0f);
            ps.push_back(s);
        }
        auto cl = ahs_oracle::cluster_edit((int)n, ps);
        std::vector<std::vector<StaticSparseGraph::NodeId>> out;
        for (auto& c : cl) out.emplace_back(c.begin(), c.end());
        return ClusterEditingSolution(out);
    }
private:
    TriangleSparseMatrix& m_;
};
