// TEST INFRASTRUCTURE.  Stand-in for WhatsHap ReadScoring (reference
// src/alignmentstoreadset.cpp:308-311).  Algorithm: oracle/core/phase_core.hpp rule R1.
// Scores are stored as float = w / 1024 (exact: |w| <= 2^17).  PARITY UNPINNED.
#pragma once
#include "../../core/phase_core.hpp"
#include "../readset.h"
#include "trianglesparsematrix.h"
// Side channel scoring -> solver: a sparse matrix cannot say how many reads exist; the reference
// always builds the solver right after scoring (src/alignmentstoreadset.cpp:311-312), and every
// read (also one that overlaps nobody) must end up in a cluster, as in rule R2.
inline uint32_t& ahs_shim_num_reads() { static thread_local uint32_t v = 0; return v; }
class ReadScoring {
public:
    ReadScoring() {}
    void scoreReadsetLocal(TriangleSparseMatrix* result, ReadSet* readset, uint32_t minOverlap = 1, uint32_t ploidy = 2) const {
        std::vector<ahs_oracle::Row> rows(readset->size());
        for (int i = 0; i < readset->size(); i++) {
            Read* r = readset->get(i);
            for (int v = 0; v < r->getVariantCount(); v++) { rows[i].pos.push_back(r->getPosition(v)); rows[i].allele.push_back(r->getAllele(v)); }
        }
        std::vector<ahs_oracle::PairScore> ps;
        ahs_oracle::score_reads_local(rows, minOverlap, ploidy, ps);
        // every scored pair gets an entry (also weight 0) so the dimension of the matrix = #reads seen
        for (auto& s : ps) result->set((uint32_t)s.i, (uint32_t)s.j, (float)s.w / 1024.0f);
        ahs_shim_num_reads() = (uint32_t)readset->size();
    }
};
