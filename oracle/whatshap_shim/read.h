// oracle/whatshap_shim/read.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// From-scratch stand-in for WhatsHap's `Read` (src/read.h @ 8f4c0c07, absent here) covering
// exactly the members the reference calls (src/alignmentstoreadset.cpp:94,117-129,155-180,
// 234-248,267-270,596-598,618,678).  Container semantics only; PARITY UNPINNED for toString().
#pragma once
#include <algorithm>
#include <cstdint>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

class Read {
public:
    Read(const std::string& name, int mapq, int source_id, int sample_id, int reference_start = -1,
         const std::string& BX_tag = "")
        : name_(name), mapqs_(1, mapq), source_id_(source_id), sample_id_(sample_id),
          reference_start_(reference_start), bx_(BX_tag) {}
    const std::string& getName() const { return name_; }
    const std::vector<int>& getMapqs() const { return mapqs_; }
    int getSourceID() const { return source_id_; }
    void addVariant(int position, int allele, int quality) { v_.push_back({position, allele, quality}); }
    void sortVariants() {
        std::sort(v_.begin(), v_.end(), [](const Var& a, const Var& b) { return a.position < b.position; });
        for (size_t i = 1; i < v_.size(); i++)
            if (v_[i].position == v_[i - 1].position) throw std::runtime_error("Duplicate variant in read " + name_);
    }
    int getVariantCount() const { return (int)v_.size(); }
    int firstPosition() const { if (v_.empty()) throw std::runtime_error("No variants in read"); return v_.front().position; }
    int lastPosition() const { if (v_.empty()) throw std::runtime_error("No variants in read"); return v_.back().position; }
    int getPosition(size_t i) const { return v_.at(i).position; }
    int getAllele(size_t i) const { return v_.at(i).allele; }
    int getVariantQuality(size_t i) const { return v_.at(i).quality; }
    void addPositionsToSet(std::unordered_set<unsigned int>* s) const { for (auto& x : v_) s->insert((unsigned)x.position); }
    std::string toString() const {
        std::ostringstream oss;
        oss << name_ << " (";
        for (size_t i = 0; i < v_.size(); i++) { if (i) oss << ";"; oss << "[" << v_[i].position << "," << v_[i].allele << "," << v_[i].quality << "]"; }
        oss << ")";
        return oss.str();
    }
private:
    struct Var { int position, allele, quality; };
    std::string name_;
    std::vector<int> mapqs_;
    int source_id_, sample_id_, reference_start_;
    std::string bx_;
    std::vector<Var> v_;
};
