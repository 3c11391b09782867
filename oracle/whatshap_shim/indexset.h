// oracle/whatshap_shim/indexset.h — TEST INFRASTRUCTURE.  Stand-in for WhatsHap IndexSet
// (call sites reference src/alignmentstoreadset.cpp:150,159,263,271).
#pragma once
#include <set>
class IndexSet {
public:
    void add(int i) { s_.insert(i); }
    bool contains(int i) const { return s_.count(i) != 0; }
    size_t size() const { return s_.size(); }
    std::set<int>::const_iterator begin() const { return s_.begin(); }
    std::set<int>::const_iterator end() const { return s_.end(); }
private:
    std::set<int> s_;
};
