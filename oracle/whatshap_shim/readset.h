// oracle/whatshap_shim/readset.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Stand-in for WhatsHap ReadSet (call sites reference src/alignmentstoreadset.cpp:79,89,116-120,
// 144,151-152,163,193-195,233-237,274,288-303,317,346).  sort() is std::sort with a comparator
// that looks at firstPosition() only — the libstdc++ introsort is therefore part of the
// observable behaviour.  PARITY UNPINNED for toString().
#pragma once
#include <map>
#include <unordered_map>
#include "indexset.h"
#include "read.h"

class ReadSet {
public:
    ReadSet() {}
    void add(Read* r) {
        auto key = std::make_pair(r->getName(), r->getSourceID());
        if (index_.count(key)) throw std::runtime_error("ReadSet::add: duplicate read name " + r->getName());
        index_[key] = reads_.size();
        reads_.push_back(r);
    }
    Read* getByName(const std::string& name, int source_id) const {
        auto it = index_.find(std::make_pair(name, source_id));
        return it == index_.end() ? 0 : reads_[it->second];
    }
    int size() const { return (int)reads_.size(); }
    Read* get(int i) const { return reads_.at(i); }
    ReadSet* subset(const IndexSet* idx) const {
        ReadSet* r = new ReadSet();
        for (int i : *idx) r->add(new Read(*reads_.at(i)));
        return r;
    }
    std::vector<unsigned int>* get_positions() const {
        std::set<unsigned int> s;
        for (auto* r : reads_) for (int i = 0; i < r->getVariantCount(); i++) s.insert((unsigned)r->getPosition(i));
        return new std::vector<unsigned int>(s.begin(), s.end());
    }
    void sort() {
        std::sort(reads_.begin(), reads_.end(), [](const Read* a, const Read* b) { return a->firstPosition() < b->firstPosition(); });
        index_.clear();
        for (size_t i = 0; i < reads_.size(); i++) index_[std::make_pair(reads_[i]->getName(), reads_[i]->getSourceID())] = i;
    }
    std::string toString() const {
        std::ostringstream oss;
        oss << "ReadSet:" << std::endl;
        for (auto* r : reads_) oss << "  " << r->toString() << std::endl;
        return oss.str();
    }
private:
    std::vector<Read*> reads_;
    std::map<std::pair<std::string, int>, size_t> index_;
};
