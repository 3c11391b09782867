// graph_native.cpp — SURVEY §8 row f4: GFA parsing and bubble/chain detection at scale.
// Result-identical replacements of
//     Graph::ReadGraph     (reference src/graph.cpp:186-249)
//     Graph::findBubbles   (reference src/graph.cpp:342-379)
//     Graph::findBubble    (reference src/graph.cpp:381-500)
// that fill the reference's own Graph object (graph.hpp), so everything downstream — the
// reference's or this repo's — sees the same nodes, edges, offsets, chains, bubbles and ids.
//
// Why: the reference passes Node objects BY VALUE everywhere (`for (auto node: nodes)`,
// findBubble(Node node, …), set<pair<Node,bool>>, `children = v.first.childrenleft`), and a Node owns
// its sequence string: every step of the detection copies unitig sequences.  Here the detection walks
// node ids; Node copies are made only where the result type holds them (Bubble::source/sink/innerNodes),
// with the field values the reference's copies have at that moment.
//
// Identity of the containers: chain ids follow the iteration order of Graph::nodes
// (std::unordered_map).  read_gfa performs the same container operations in the
// same order as ReadGraph (operator[] / assignment / push_back per line), so the tables — and their
// iteration order — are the reference's.
#include <cerrno>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "graph.hpp"   // reference header

namespace ahs_host {
namespace {

inline bool is_ws(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
struct Tok { const char* p; size_t n; bool is(const char* s) const { return n == strlen(s) && memcmp(p, s, n) == 0; } };

// raw_id (graph.cpp:160-164): the digits of the name, then stoi
bool raw_id(Tok t, int& out) {
    int64_t v = 0; bool any = false;
    for (size_t i = 0; i < t.n; i++) {
        const unsigned char c = (unsigned char)t.p[i];
        if (c >= '0' && c <= '9') { any = true; v = v * 10 + (c - '0'); if (v > INT_MAX) return false; }
    }
    if (!any) return false;
    out = (int)v;
    return true;
}

}  // namespace

// Graph::ReadGraph.  Returns 0, or 66 with `err` = "file:line: reason" where the reference dies on an assert or an
// uncaught std::stoi exception.
int read_gfa(const std::string& filename, Graph& graph, std::string& err) {
    const char* base = nullptr; size_t size = 0; bool mapped = false; std::string owned;
    int fd = open(filename.c_str(), O_RDONLY);
    if (fd >= 0) {
        struct stat sb;
        if (fstat(fd, &sb) == 0 && sb.st_size > 0) {
            void* a = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (a != MAP_FAILED) { base = (const char*)a; size = (size_t)sb.st_size; mapped = true; madvise(a, size, MADV_SEQUENTIAL); }
            else {
                owned.resize((size_t)sb.st_size);
                size_t got = 0; ssize_t r;
                while (got < owned.size() && (r = read(fd, &owned[got], owned.size() - got)) > 0) got += (size_t)r;
                owned.resize(got); base = owned.data(); size = got;
            }
        }
        close(fd);
    }
    size_t valid = size;                                      // `if (!file.good()) break;` (:192): unterminated last line is dropped
    while (valid > 0 && base[valid - 1] != '\n') valid--;
    int rc = 0;
    int64_t lineno = 0;
    auto fail = [&](const char* what) { err = filename + ":" + std::to_string(lineno) + ": " + what; rc = 66; };
    for (const char* ls = base; ls < base + valid && !rc;) {
        const char* le = (const char*)memchr(ls, '\n', (size_t)(base + valid - ls));
        lineno++;
        const char* cur = ls;
        ls = le + 1;
        if (le == cur) continue;                              // :193
        if (*cur != 'S' && *cur != 'L') continue;             // :194
        Tok tok[7]; int nt = 0;
        const char* p = cur;
        const int want = *cur == 'S' ? 3 : 5;
        while (nt < want) {
            while (p < le && is_ws((unsigned char)*p)) p++;
            if (p == le) break;
            const char* q = p;
            while (q < le && !is_ws((unsigned char)*q)) q++;
            tok[nt++] = Tok{p, (size_t)(q - p)};
            p = q;
        }
        for (int i = nt; i < 7; i++) tok[i] = Tok{le, 0};
        if (*cur == 'S') {
            if (!tok[0].is("S")) { fail("record type is not exactly S (assert, graph.cpp:201)"); break; }
            int id;
            if (!raw_id(tok[1], id)) { fail("segment name without a usable integer id (stoi, graph.cpp:163)"); break; }
            if (tok[2].n < 1) { fail("segment without a sequence (assert, graph.cpp:205)"); break; }
            graph.nodes[id] = Node(id, std::string(tok[2].p, tok[2].n));                       // :208
        } else {
            if (!tok[0].is("L")) { fail("record type is not exactly L (assert, graph.cpp:219)"); break; }
            int start_id, end_id;
            if (!raw_id(tok[1], start_id)) { fail("link start without a usable integer id (stoi, graph.cpp:163)"); break; }
            if (!raw_id(tok[3], end_id)) { fail("link end without a usable integer id (stoi, graph.cpp:163)"); break; }
            const bool sp = tok[2].is("+"), ep = tok[4].is("+");
            if (!(sp || tok[2].is("-")) || !(ep || tok[4].is("-"))) { fail("link orientation is not + or - (assert, graph.cpp:225-226)"); break; }
            // `sstr >> offset` (int) then `sstr >> dummyc` (char), :227-230
            while (p < le && is_ws((unsigned char)*p)) p++;
            const char* q = p;
            if (q < le && (*q == '+' || *q == '-')) q++;
            const char* dig = q;
            while (q < le && *q >= '0' && *q <= '9') q++;
            if (q == dig) { fail("overlap does not start with an integer (assert on an unread char, graph.cpp:230)"); break; }
            char buf[32];
            if ((size_t)(q - p) >= sizeof buf) { fail("overlap out of int range"); break; }
            memcpy(buf, p, (size_t)(q - p)); buf[q - p] = 0;
            errno = 0;
            const long off = strtol(buf, nullptr, 10);
            if (errno == ERANGE || off > INT_MAX || off < INT_MIN) { fail("overlap out of int range"); break; }
            while (q < le && is_ws((unsigned char)*q)) q++;
            const char dummyc = q < le ? *q : 0;
            if (!(dummyc == 'M' || (dummyc == 'S' && off == 0))) { fail("overlap is not <int>M (assert, graph.cpp:230)"); break; }
            if (off < 0) { fail("negative overlap (assert, graph.cpp:232)"); break; }
            DirectedNode from(start_id, sp);
            DirectedNode to(end_id, ep);
            graph.edges[from].push_back(to);                                                    // :236
            if (sp) graph.nodes[start_id].childrenleft.push_back(std::make_pair(graph.nodes[end_id].node_id, ep));    // :237-238
            else graph.nodes[start_id].childrenright.push_back(std::make_pair(graph.nodes[end_id].node_id, ep));      // :239-241
            graph.offsets[std::make_pair(from, to)] = (size_t)off;                               // :242
        }
    }
    if (mapped) munmap((void*)base, size);
    return rc;
}

namespace {

struct Ref { int id; bool dir; bool visited_snap; };     // one pair<Node,bool> of the reference: which node, and the `visited` its copy carries

inline Node snapshot(const Graph& g, const Ref& r) {
    Node n = g.nodes.find(r.id)->second;
    n.visited = r.visited_snap;
    return n;
}

// Graph::findBubble (graph.cpp:381-500).  The recursion at :497 is the last thing its caller does (S is empty when it
// returns), so it is a loop here: chains of any length use constant stack.
void find_bubble(Graph& g, Ref node, Chain* bchain) {
    std::unordered_set<DirectedNode> seen;
    std::unordered_set<int> visited;
    std::vector<Ref> inside, S;                              // S: the reference's std::set ordered by (node_id, bool), keys unique
    for (;;) {
        seen.clear(); visited.clear(); inside.clear(); S.clear();
        seen.insert(DirectedNode(node.id, node.dir));
        S.push_back(node);
        bool chained = false;
        Ref t{0, false, false};
        while (!S.empty()) {
            const Ref v = S.front();
            S.erase(S.begin());
            visited.insert(v.id);
            Node& vn = g.nodes.find(v.id)->second;
            vn.visited = true;
            inside.push_back(v);
            seen.erase(DirectedNode(v.id, v.dir));
            const std::vector<std::pair<int, int>>& children = v.dir == false ? vn.childrenleft : vn.childrenright;
            if (children.empty()) break;                                                        // tip (:408-409)
            for (auto& u : children) {
                const Node& un = g.nodes.find(u.first)->second;
                const std::vector<std::pair<int, int>>& u_parents = u.second == 0 ? un.childrenleft : un.childrenright;
                const bool u_child_direction = u.second == 0;
                if (u.first == node.id) { S.clear(); break; }                                   // loop found (:431-437)
                seen.insert(DirectedNode(u.first, u.second == 0));
                bool all_visited = true;
                for (auto& p : u_parents) if (visited.find(p.first) == visited.end()) all_visited = false;
                if (all_visited) {                                                              // S.insert(make_pair(nodes[u.first], …)) (:478)
                    size_t at = 0;
                    while (at < S.size() && (S[at].id < u.first || (S[at].id == u.first && S[at].dir < u_child_direction))) at++;
                    if (at == S.size() || S[at].id != u.first || S[at].dir != u_child_direction)
                        S.insert(S.begin() + at, Ref{u.first, u_child_direction, un.visited});
                }
            }
            if (S.size() == 1 && seen.size() == 1) {                                            // :482
                t = S.front();
                S.clear();
                inside.push_back(t);
                if (inside.size() == 2) break;
                for (size_t i = 0; i < inside.size(); i++) if (inside[i].id == node.id) { inside.erase(inside.begin() + i); break; }
                for (size_t i = 0; i < inside.size(); i++) if (inside[i].id == t.id) { inside.erase(inside.begin() + i); break; }
                bchain->bubbles.emplace_back();                                                  // Bubble(node, t.first, nodesInside), id 0; addBubble —
                Bubble& bubble = bchain->bubbles.back();                                         // one copy per node instead of three
                bubble.source = snapshot(g, node);
                bubble.sink = snapshot(g, t);
                bubble.innerNodes.reserve(inside.size());
                for (auto& r : inside) bubble.innerNodes.push_back(snapshot(g, r));
                chained = true;
                break;                                                                          // findBubble(t.first, t.second, bchain), then S is empty
            }
        }
        if (!chained) return;
        node = t;
    }
}

}  // namespace

// Graph::findBubbles.  Returns 0, or 67 when a link names a segment that has no S line (the reference then inserts
// default nodes into Graph::nodes while iterating over it: undefined behaviour, not reproduced).
int find_bubbles(Graph& g, std::string& err) {
    for (auto& kv : g.nodes) {
        if (kv.second.node_id != kv.first) {                  // default node made by an L line (graph.cpp:238, 241)
            err = "link from a segment without an S line (id " + std::to_string(kv.first) + ")";
            return 67;
        }
        for (auto* list : {&kv.second.childrenleft, &kv.second.childrenright})
            for (auto& ch : *list)
                if (g.nodes.find(ch.first) == g.nodes.end()) {
                    err = "link from segment " + std::to_string(kv.first) + " to a segment without an S line (id " + std::to_string(ch.first) + ")";
                    return 67;
                }
    }
    for (auto& kv : g.nodes) {                                // `for (auto node: nodes)` (:344): the copy's flag is the flag now
        if (kv.second.visited) continue;
        Chain bchain;
        for (int i : {0, 1}) find_bubble(g, Ref{kv.first, i == 1, false}, &bchain);     // both calls get the same unvisited copy
        if (bchain.bubbles.size() != 0) g.chains.push_back(std::move(bchain));        // :354; the loop at :355-356 marks copies only
    }
    int chain_id = 0;
    for (auto& chain : g.chains) {                                                      // :361-374
        chain.id = chain_id;
        int bubble_id = 0;
        for (auto& bubble : chain.bubbles) {
            bubble.id = bubble_id;
            auto tag = [&](int id) { Node& n = g.nodes[id]; n.chain_id = chain_id; n.bubble_id = bubble_id; };
            tag(bubble.source.node_id);
            for (auto& n : bubble.innerNodes) tag(n.node_id);
            tag(bubble.sink.node_id);
            bubble_id += 1;
        }
        chain_id += 1;
    }
    return 0;
}

}  // namespace ahs_host
