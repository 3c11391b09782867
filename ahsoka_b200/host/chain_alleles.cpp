// chain_alleles.cpp — SURVEY §8 rows a1 / f4 (allele-path enumeration at scale).
// Result-identical replacement of ChainsToReadsetDetailed + findPathsSimple / findPathsComplex /
// addSequence (reference src/chainstoreadset.cpp:161-203, 17-30, 84-116, 44-82).
//
// The reference takes the Graph BY VALUE (every node sequence copied), copies each Bubble (three or
// more Nodes with sequences) per call, and re-assigns the whole per-chain map after every bubble
// (:183, :198) — O(B²) per chain, 4.4 s for one 10,000-bubble chain.  Here the bubbles are read in
// place and the per-chain map is assigned once; because the map is built by the same insertions in
// the same order, its iteration order (stage A's bubble order, alignmentstoreadset.cpp:76,90) is the
// reference's.
#include <algorithm>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "graph.hpp"   // reference header

namespace ahs_host {

typedef std::unordered_map<int, std::unordered_map<int, std::vector<std::vector<int>>>> ChainAlleles;

namespace {

struct BubbleView {
    const Bubble& b;
    std::vector<int> ids;                      // Bubble::getNodeIds(): source, inner nodes, sink (graph.cpp:106-113)
    explicit BubbleView(const Bubble& bb) : b(bb) {
        ids.push_back(b.source.node_id);
        for (auto& n : b.innerNodes) ids.push_back(n.node_id);
        ids.push_back(b.sink.node_id);
    }
    bool has(int id) const { return std::find(ids.begin(), ids.end(), id) != ids.end(); }
    const Node& node(int id) const {           // Bubble::getNode (graph.cpp:115-126): source, then sink, then first inner match
        if (b.source.node_id == id) return b.source;
        if (b.sink.node_id == id) return b.sink;
        for (auto& n : b.innerNodes) if (n.node_id == id) return n;
        throw std::logic_error("bubble node not found");   // unreachable: callers check has() first
    }
};

// addSequence (chainstoreadset.cpp:44-82)
void add_sequence(const Node& node, int direction, const BubbleView& bv, std::vector<int>& seq,
                  std::vector<std::vector<int>>& paths, int depth) {
    if (depth > 100000) throw std::runtime_error("allele-path enumeration does not terminate (cyclic bubble)");
    if (std::find(seq.begin(), seq.end(), node.node_id) == seq.end()) seq.push_back(node.node_id);
    const std::vector<std::pair<int, int>>& next = direction == 0 ? node.childrenright : node.childrenleft;
    bool within = true;
    for (auto& child : next) if (!bv.has(child.first)) within = false;
    if (!next.empty() && within) {
        for (auto& child : next) {
            const size_t index = (size_t)(std::find(seq.begin(), seq.end(), node.node_id) - seq.begin());
            std::vector<int> prefix(seq.begin(), seq.begin() + index + 1);
            add_sequence(bv.node(child.first), child.second, bv, prefix, paths, depth + 1);
        }
    } else paths.push_back(seq);
}

}  // namespace

ChainAlleles chain_alleles(const Graph& graph) {
    ChainAlleles out;
    for (auto& chain : graph.chains) {
        std::unordered_map<int, std::vector<std::vector<int>>> per_bubble;
        for (auto& bubble : chain.bubbles) {
            std::vector<std::vector<int>>& alleles = per_bubble[bubble.id];
            if (bubble.innerNodes.size() == 2) {
                // findPathsSimple (:17-30): [source, inner_i, sink] in innerNodes order
                for (auto& inner : bubble.innerNodes) alleles.push_back({bubble.source.node_id, inner.node_id, bubble.sink.node_id});
            } else {
                // findPathsComplex (:84-116): depth-first from the SINK, to the right unless a right child leaves the bubble
                BubbleView bv(bubble);
                bool right_within = true;
                for (auto& r : bubble.sink.childrenright) if (!bv.has(r.first)) right_within = false;
                std::vector<int> seq;
                add_sequence(bubble.sink, right_within ? 0 : 1, bv, seq, alleles, 0);
            }
        }
        // the last of the reference's per-bubble copy assignments; a copy-assigned libstdc++ unordered_map takes the bucket
        // count and the element order of its source, which is exactly what moving the source hands over
        if (!chain.bubbles.empty()) out[chain.id] = std::move(per_bubble);
    }
    return out;
}

}  // namespace ahs_host
