// TEST INFRASTRUCTURE.  Stand-in for WhatsHap TriangleSparseMatrix (reference
// src/alignmentstoreadset.cpp:309): sparse symmetric float matrix keyed by (i<j).
#pragma once
#include <cstdint>
#include <map>
#include <utility>
#include <vector>
class TriangleSparseMatrix {
public:
    TriangleSparseMatrix() {}
    void set(uint32_t i, uint32_t j, float v) { if (i == j) return; m_[key(i, j)] = v; }
    float get(uint32_t i, uint32_t j) const { auto it = m_.find(key(i, j)); return it == m_.end() ? 0.0f : it->second; }
    unsigned int size() const { return (unsigned)m_.size(); }
    std::vector<std::pair<uint32_t, uint32_t>> getEntries() const {
        std::vector<std::pair<uint32_t, uint32_t>> r; for (auto& kv : m_) r.push_back(kv.first); return r;
    }
    uint32_t getMaxDim() const { uint32_t d = 0; for (auto& kv : m_) d = std::max(d, kv.first.second + 1); return d; }
private:
    static std::pair<uint32_t, uint32_t> key(uint32_t i, uint32_t j) { return i < j ? std::make_pair(i, j) : std::make_pair(j, i); }
    std::map<std::pair<uint32_t, uint32_t>, float> m_;
};
