// oracle/core/phase_core.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the three algorithms the reference calls but does not contain:
//   * ReadScoring::scoreReadsetLocal      (call site reference src/alignmentstoreadset.cpp:308-311)
//   * ClusterEditingSolver::run           (call site :312-315)
//   * HaploThreader::computePaths         (call site :320, :408)
// They live in WhatsHap @ 8f4c0c070d0b5d8e6d2c03e363965d0efb50a960 (reference
// container/ahsoka.def:16-18, src/CMakeLists.txt:8-19), which is NOT vendored and NOT
// available offline.  The reference has no tests, fixtures or golden vectors.
//
//   >>> PARITY UNPINNED for these three algorithms. <<<
//
// What follows restates the *published* algorithms (Schrinner et al., "Haplotype
// threading: accurate polyploid phasing from long reads", Genome Biology 2020; Boecker
// et al., "Exact algorithms for cluster editing: evaluation and experiments", 2011) with
// every free choice (tie-breaks, number formats, ploidy generalisation) fixed and written
// down here.  The CUDA path must agree with THIS file bit for bit.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may execute anything under oracle/.
//
// ---------------------------------------------------------------------------------------
// R1. Read-pair scoring (scoreReadsetLocal, minOverlap, ploidy p)
//   * For reads i<j (order = ReadSet order after sort()), n_ij = #positions both cover,
//     k_ij = #of those where the alleles differ.  Pairs with n_ij < minOverlap have no entry.
//   * Local error-rate estimate per read i: take all partners j (j<i and j>i) with
//     n_ij >= minOverlap, order them by Hamming rate k/n ascending (exact rational
//     compare), ties by n ascending; the first cut = max(1, m/p) (m = #partners, integer
//     division) are taken to be same-haplotype pairs (1/p of the pairs are expected to be),
//     the rest different-haplotype pairs.  Pooled rates in 1/1024 units:
//         es_i = rq(sum k_same, sum n_same),  ed_i = rq(sum k_diff, sum n_diff) (= es_i if
//         the diff set is empty),  rq(K,N) = floor((1024 K + floor(N/2)) / N).
//   * Pair rates: es = floor((es_i+es_j)/2) clamped to [10,460]; ed = floor((ed_i+ed_j)/2)
//     clamped to [es+51, 972].
//   * Score = log( Binom(k;n,es) / Binom(k;n,ed) ) in fixed point:
//         s20 = k*(LN[es]-LN[ed]) + (n-k)*(LN1[es]-LN1[ed]),
//         LN[x] = llrint(ln(x/1024) * 2^20),  LN1[x] = llrint(ln(1 - x/1024) * 2^20),
//         w = floor(s20 / 1024) clamped to [-2^17, 2^17]      (Q10: w/1024 is the float score)
//     All arithmetic after the two tables is integer, hence order independent.
//
// R2. Cluster editing (induced-cost greedy heuristic with node merging)
//   * Nodes = reads.  W[a][b] = w of R1, 0 where no entry.  FORB = forbidden (-inf).
//   * Candidates = pairs of active nodes with W != 0 and W != FORB.
//         icf(a,b) = max(0, W_ab) + sum_c tf(W_ac, W_bc)   tf(x,y) = (x>0 && y>0) ? min(x,y) : 0
//         icp(a,b) = max(0,-W_ab) + sum_c tp(W_ac, W_bc)   tp(x,y) = x>0&&y<0 ? min(x,|y|)
//                                                                  : x<0&&y>0 ? min(|x|,y) : 0
//     with |FORB| = +inf, c over active nodes other than a,b.
//   * Repeat until no candidate is left: eF = argmax icf, eP = argmax icp (ties: smallest
//     (a,b), a<b, lexicographic).  If icf(eF) >= icp(eP): merge eF=(a,b) into a:
//     W_ac <- FORB if W_ac or W_bc is FORB else W_ac + W_bc; b becomes inactive.
//     Otherwise W(eP) <- FORB.
//   * Clusters = member sets of the active nodes, numbered by smallest member, members
//     ascending (ClusterEditingSolution::getCluster).
//
// R3. Haplotype threading (computePaths with symmetryOptimization=false, rowLimit=0)
//   * Column q: tuples t in [0,k_q)^p over LOCAL indices of covMap[q]; code(t) reads t as a
//     base-k_q number with haplotype 0 the most significant digit.
//   * Genotype conformity: multiset{consensus[q][t_h]} == genotypes[q] (allele -> count).
//     An EMPTY genotype map means "heterozygous, dosage unknown": conform iff the
//     consensus alleles of the tuple are not all equal (used for ploidy > 2, which has no
//     reference behaviour: alignmentstoreadset.cpp:306,341-344 hard-code p=2, {0:1,1:1}).
//     If no tuple of a column conforms, all tuples are allowed.
//   * Coverage cost of t = #haplotypes h whose cluster l=t_h (multiplicity m in t) has
//     coverage[q][l] < (2m-1)/(2p) or > (2m+1)/(2p).
//   * Transition s->t = switchCost * #{h : covMap[q-1][s_h] != covMap[q][t_h]}
//                       + affineSwitchCost * [any h switched].
//   * D_q[t] = covcost(t) + min_s (D_{q-1}[s] + trans); predecessor ties and the final
//     column minimum take the lowest code.  Output path[q][h] = covMap[q][t_h].
//
// R3c. Haplotype threading for ploidy > 4 (no reference behaviour at all: the reference fixes p = 2; k^p ordered
//     tuples with k <= 2p is 2.99 M states per column at p = 6).  Same costs on CANONICAL tuples:
//   * Column q uses the first k_q = min(|covMap[q]|, p + 2) entries of covMap[q] (descending coverage).
//   * States = multisets of p local indices, written as non-decreasing tuples t_0 <= ... <= t_{p-1}; the state's
//     index is the tuple's rank in lexicographic order.  Conformity and coverage cost are R3's (they depend on the
//     multiset only).
//   * Transition s->t = switchCost * (p - |G(s) n G(t)|) + affineSwitchCost * [G(s) != G(t)], G(.) = the multiset of
//     GLOBAL cluster ids: the minimum of R3's transition cost over all ways of labelling the haplotypes.
//   * D_q[t], predecessor ties and the final minimum: lowest index.
//   * Haplotype labels are threaded along the chain: path[0][h] = covMap[0][t_h]; at q > 0, for h = 0..p-1 in turn,
//     haplotype h keeps its cluster while G(t) still holds an unclaimed copy of it, then the haplotypes left over
//     take the unclaimed elements of t in ascending local index.  (Exactly p - |G(s) n G(t)| haplotypes switch.)
//   Two implementations below: thread_paths_canonical_direct (the definition, all pairs of states) and
//   thread_paths_canonical (sub-multiset DP, what the parity tests use at size); tests/test_oracle_cpu.py checks
//   that they agree.
// ---------------------------------------------------------------------------------------
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <set>
#include <unordered_map>
#include <vector>

namespace ahs_oracle {

// ------------------------------------------------------------------ R1 scoring
struct Row {                      // one read: ascending positions + alleles
    std::vector<int32_t> pos;
    std::vector<int32_t> allele;
};
struct PairScore { int32_t i, j; int32_t n, k; int32_t w; };   // i<j, w in Q10

static const int32_t W_CLAMP = 1 << 17;
static const int32_t ES_MIN = 10, ES_MAX = 460, ED_GAP = 51, ED_MAX = 972;

struct LogTables {
    int64_t ln[1025], ln1[1025];
    LogTables() {
        for (int x = 0; x <= 1024; x++) {
            ln[x]  = (x >= 1)    ? llrint(std::log((double)x / 1024.0) * 1048576.0) : 0;
            ln1[x] = (x <= 1023) ? llrint(std::log(1.0 - (double)x / 1024.0) * 1048576.0) : 0;
        }
    }
};
inline const LogTables& log_tables() { static LogTables t; return t; }

inline int64_t floordiv(int64_t a, int64_t b) {   // b > 0
    int64_t q = a / b, r = a % b;
    return (r != 0 && r < 0) ? q - 1 : q;
}
inline int32_t rate_q(int64_t K, int64_t N) { return (int32_t)((K * 1024 + N / 2) / N); }

inline void overlap_diff(const Row& a, const Row& b, int32_t& n, int32_t& k) {
    n = 0; k = 0;
    size_t x = 0, y = 0;
    while (x < a.pos.size() && y < b.pos.size()) {
        if (a.pos[x] < b.pos[y]) x++;
        else if (a.pos[x] > b.pos[y]) y++;
        else { n++; if (a.allele[x] != b.allele[y]) k++; x++; y++; }
    }
}

inline int32_t pair_weight(int32_t n, int32_t k, int32_t es_i, int32_t ed_i, int32_t es_j, int32_t ed_j) {
    const LogTables& T = log_tables();
    int32_t es = (es_i + es_j) / 2, ed = (ed_i + ed_j) / 2;
    es = std::min(std::max(es, ES_MIN), ES_MAX);
    ed = std::min(std::max(ed, es + ED_GAP), ED_MAX);
    int64_t s20 = (int64_t)k * (T.ln[es] - T.ln[ed]) + (int64_t)(n - k) * (T.ln1[es] - T.ln1[ed]);
    int64_t w = floordiv(s20, 1024);
    if (w > W_CLAMP) w = W_CLAMP;
    if (w < -W_CLAMP) w = -W_CLAMP;
    return (int32_t)w;
}

// reads must be ordered so that first positions are non-decreasing (ReadSet::sort()); the
// function itself does not rely on it (all pairs are tested through an interval sweep).
inline void score_reads_local(const std::vector<Row>& reads, uint32_t minOverlap, uint32_t ploidy,
                              std::vector<PairScore>& out, std::vector<int32_t>* es_out = nullptr,
                              std::vector<int32_t>* ed_out = nullptr) {
    const int R = (int)reads.size();
    out.clear();
    // candidate pairs: span intersection
    std::vector<std::vector<std::pair<int32_t, int32_t>>> partners(R);   // (k, n) per partner
    std::vector<int> order(R);
    for (int i = 0; i < R; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        int fa = reads[a].pos.empty() ? 0 : reads[a].pos.front();
        int fb = reads[b].pos.empty() ? 0 : reads[b].pos.front();
        return fa < fb;
    });
    for (int x = 0; x < R; x++) {
        int i = order[x];
        if (reads[i].pos.empty()) continue;
        int last_i = reads[i].pos.back();
        for (int y = x + 1; y < R; y++) {
            int j = order[y];
            if (reads[j].pos.empty()) continue;
            if (reads[j].pos.front() > last_i) break;
            int32_t n, k;
            overlap_diff(reads[i], reads[j], n, k);
            if (n >= (int32_t)minOverlap && n > 0) {
                PairScore ps; ps.i = std::min(i, j); ps.j = std::max(i, j); ps.n = n; ps.k = k; ps.w = 0;
                out.push_back(ps);
                partners[i].push_back({k, n});
                partners[j].push_back({k, n});
            }
        }
    }
    std::vector<int32_t> es(R, 0), ed(R, 0);
    for (int i = 0; i < R; i++) {
        auto& P = partners[i];
        const int m = (int)P.size();
        if (m == 0) continue;
        std::sort(P.begin(), P.end(), [](const std::pair<int32_t, int32_t>& a, const std::pair<int32_t, int32_t>& b) {
            int64_t l = (int64_t)a.first * b.second, r = (int64_t)b.first * a.second;   // k_a/n_a < k_b/n_b
            if (l != r) return l < r;
            if (a.second != b.second) return a.second < b.second;
            return a.first < b.first;
        });
        int cut = std::max(1, m / (int)ploidy);
        int64_t Ks = 0, Ns = 0, Kd = 0, Nd = 0;
        for (int x = 0; x < m; x++) {
            if (x < cut) { Ks += P[x].first; Ns += P[x].second; }
            else         { Kd += P[x].first; Nd += P[x].second; }
        }
        es[i] = rate_q(Ks, Ns);
        ed[i] = Nd > 0 ? rate_q(Kd, Nd) : es[i];
    }
    for (auto& ps : out) ps.w = pair_weight(ps.n, ps.k, es[ps.i], ed[ps.i], es[ps.j], ed[ps.j]);
    std::sort(out.begin(), out.end(), [](const PairScore& a, const PairScore& b) {
        return a.i != b.i ? a.i < b.i : a.j < b.j;
    });
    if (es_out) *es_out = es;
    if (ed_out) *ed_out = ed;
}

// ------------------------------------------------------------------ R2 cluster editing
static const int32_t FORB = std::numeric_limits<int32_t>::min();
static const int64_t INF64 = (int64_t)1 << 60;

inline int64_t tf(int32_t x, int32_t y) { return (x > 0 && y > 0) ? (int64_t)std::min(x, y) : 0; }
inline int64_t absw(int32_t x) { return x == FORB ? INF64 : (x < 0 ? -(int64_t)x : (int64_t)x); }
inline int64_t tp(int32_t x, int32_t y) {
    if (x > 0 && y < 0) return std::min((int64_t)x, absw(y));
    if (x < 0 && y > 0) return std::min(absw(x), (int64_t)y);
    return 0;
}

struct ClusterEditStats { int64_t steps = 0, merges = 0, forbids = 0; };

// Dense implementation with incremental (exact, integer) maintenance of icf/icp.
// `paranoid` re-derives every candidate's icf/icp from the definition after each step.
inline std::vector<std::vector<int32_t>> cluster_edit(int n, const std::vector<PairScore>& scores,
                                                     bool paranoid = false, ClusterEditStats* stats = nullptr) {
    std::vector<std::vector<int32_t>> clusters;
    if (n == 0) return clusters;
    const size_t N = (size_t)n;
    std::vector<int32_t> W(N * N, 0);
    std::vector<int64_t> F(N * N, 0), P(N * N, 0);       // icf / icp, valid for candidates, index a*N+b with a<b
    std::vector<char> active(N, 1);
    std::vector<std::vector<int32_t>> members(N);
    for (int i = 0; i < n; i++) members[i].push_back(i);
    for (auto& s : scores) { W[(size_t)s.i * N + s.j] = s.w; W[(size_t)s.j * N + s.i] = s.w; }
    auto w = [&](int a, int b) -> int32_t& { return W[(size_t)a * N + b]; };
    auto is_cand = [&](int a, int b) { int32_t x = w(a, b); return active[a] && active[b] && x != 0 && x != FORB; };
    auto full_icf = [&](int a, int b) {
        int64_t s = std::max<int64_t>(0, w(a, b));
        for (int c = 0; c < n; c++) if (active[c] && c != a && c != b) s += tf(w(a, c), w(b, c));
        return s;
    };
    auto full_icp = [&](int a, int b) {
        int64_t s = std::max<int64_t>(0, -(int64_t)w(a, b));
        for (int c = 0; c < n; c++) if (active[c] && c != a && c != b) s += tp(w(a, c), w(b, c));
        return s;
    };
    // ordered candidate sets: (-value, a, b) so begin() = max value, smallest (a,b)
    typedef std::tuple<int64_t, int32_t, int32_t> Key;
    std::set<Key> SF, SP;
    auto insert_c = [&](int a, int b) { SF.insert(Key(-F[(size_t)a * N + b], a, b)); SP.insert(Key(-P[(size_t)a * N + b], a, b)); };
    auto erase_c  = [&](int a, int b) { SF.erase(Key(-F[(size_t)a * N + b], a, b)); SP.erase(Key(-P[(size_t)a * N + b], a, b)); };
    // adjacency lists (supersets of the non-zero entries of each row) keep the work sparse
    std::vector<std::vector<int32_t>> nbr(N);
    for (auto& s : scores) if (s.w != 0) { nbr[s.i].push_back(s.j); nbr[s.j].push_back(s.i); }
    for (int a = 0; a < n; a++) { std::sort(nbr[a].begin(), nbr[a].end()); nbr[a].erase(std::unique(nbr[a].begin(), nbr[a].end()), nbr[a].end()); }
    auto sparse_icf_icp = [&](int a, int b, int64_t& f, int64_t& p) {
        f = std::max<int64_t>(0, w(a, b)); p = std::max<int64_t>(0, -(int64_t)w(a, b));
        // third nodes with both weights non-zero are in nbr[a] (terms need both non-zero)
        for (int c : nbr[a]) if (active[c] && c != a && c != b) { f += tf(w(a, c), w(b, c)); p += tp(w(a, c), w(b, c)); }
    };
    for (auto& s : scores) if (s.w != 0) {
        int64_t f, p; sparse_icf_icp(s.i, s.j, f, p);
        F[(size_t)s.i * N + s.j] = f; P[(size_t)s.i * N + s.j] = p; insert_c(s.i, s.j);
    }
    ClusterEditStats st;
    std::vector<int32_t> S; std::vector<int32_t> newrow; std::vector<int32_t> pos_in_S;
    while (!SF.empty()) {
        Key kf = *SF.begin(), kp = *SP.begin();
        int64_t mF = -std::get<0>(kf), mP = -std::get<0>(kp);
        st.steps++;
        if (mF >= mP) {
            // ---- merge (a,b), a<b, into a
            const int a = std::get<1>(kf), b = std::get<2>(kf);
            st.merges++;
            S.clear();
            {
                std::vector<int32_t> tmp(nbr[a]); tmp.insert(tmp.end(), nbr[b].begin(), nbr[b].end());
                std::sort(tmp.begin(), tmp.end()); tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
                for (int x : tmp) if (active[x] && x != a && x != b && (w(a, x) != 0 || w(b, x) != 0)) S.push_back(x);
            }
            newrow.assign(S.size(), 0);
            for (size_t u = 0; u < S.size(); u++) {
                int x = S[u]; int32_t wa = w(a, x), wb = w(b, x);
                newrow[u] = (wa == FORB || wb == FORB) ? FORB : wa + wb;
            }
            // pairs inside S: swap the terms through a and b for the term through the merged node.  A candidate pair (x,y) has
            // y in the adjacency list of x, so the lists are walked instead of all |S|^2 pairs (same pairs, same arithmetic).
            if (pos_in_S.size() != N) pos_in_S.assign(N, -1);
            for (size_t u = 0; u < S.size(); u++) pos_in_S[S[u]] = (int32_t)u;
            for (size_t u = 0; u < S.size(); u++) {
                const int x = S[u];
                for (int y : nbr[x]) {
                    if (y <= x || pos_in_S[y] < 0 || !is_cand(x, y)) continue;
                    const size_t v = (size_t)pos_in_S[y];
                    int64_t df = tf(newrow[u], newrow[v]) - tf(w(x, a), w(y, a)) - tf(w(x, b), w(y, b));
                    int64_t dp = tp(newrow[u], newrow[v]) - tp(w(x, a), w(y, a)) - tp(w(x, b), w(y, b));
                    if (df != 0 || dp != 0) { erase_c(x, y); F[(size_t)x * N + y] += df; P[(size_t)x * N + y] += dp; insert_c(x, y); }
                }
            }
            for (size_t u = 0; u < S.size(); u++) pos_in_S[S[u]] = -1;
            // drop all candidates touching a or b
            erase_c(a, b);
            for (size_t u = 0; u < S.size(); u++) {
                int x = S[u];
                if (is_cand(a, x)) erase_c(std::min(a, x), std::max(a, x));
                if (is_cand(b, x)) erase_c(std::min(b, x), std::max(b, x));
            }
            active[b] = 0;
            for (size_t u = 0; u < S.size(); u++) { int x = S[u]; w(a, x) = newrow[u]; w(x, a) = newrow[u]; w(b, x) = 0; w(x, b) = 0; }
            w(a, b) = 0; w(b, a) = 0;
            members[a].insert(members[a].end(), members[b].begin(), members[b].end()); members[b].clear();
            nbr[a] = S;
            for (int x : S) { if (std::find(nbr[x].begin(), nbr[x].end(), a) == nbr[x].end()) nbr[x].push_back(a); }
            // fresh icf/icp for the pairs (a,x)
            for (size_t u = 0; u < S.size(); u++) {
                int x = S[u];
                if (!is_cand(a, x)) continue;
                int lo = std::min(a, x), hi = std::max(a, x);
                int64_t f, p; sparse_icf_icp(a, x, f, p);
                F[(size_t)lo * N + hi] = f; P[(size_t)lo * N + hi] = p; insert_c(lo, hi);
            }
        } else {
            // ---- forbid (a,b)
            const int a = std::get<1>(kp), b = std::get<2>(kp);
            st.forbids++;
            const int32_t old = w(a, b);
            erase_c(a, b);
            for (int side = 0; side < 2; side++) {
                int u = side ? b : a, v = side ? a : b;       // pairs (u,c), third node v
                for (int c : nbr[v]) {
                    if (!active[c] || c == a || c == b) continue;
                    int lo = std::min(u, c), hi = std::max(u, c);
                    if (!is_cand(lo, hi)) continue;
                    int64_t df = tf(FORB, w(c, v)) - tf(old, w(c, v));
                    int64_t dp = tp(FORB, w(c, v)) - tp(old, w(c, v));
                    if (df != 0 || dp != 0) { erase_c(lo, hi); F[(size_t)lo * N + hi] += df; P[(size_t)lo * N + hi] += dp; insert_c(lo, hi); }
                }
            }
            w(a, b) = FORB; w(b, a) = FORB;
        }
        if (paranoid) {
            size_t cnt = 0;
            for (int a = 0; a < n; a++) for (int b = a + 1; b < n; b++) if (is_cand(a, b)) {
                cnt++;
                if (F[(size_t)a * N + b] != full_icf(a, b) || P[(size_t)a * N + b] != full_icp(a, b)) {
                    fprintf(stderr, "cluster_edit paranoid check failed at (%d,%d)\n", a, b); abort();
                }
            }
            if (cnt != SF.size() || cnt != SP.size()) { fprintf(stderr, "cluster_edit candidate set mismatch\n"); abort(); }
        }
    }
    for (int a = 0; a < n; a++) if (active[a]) {
        std::sort(members[a].begin(), members[a].end());
        clusters.push_back(members[a]);
    }
    std::sort(clusters.begin(), clusters.end(), [](const std::vector<int32_t>& x, const std::vector<int32_t>& y) { return x[0] < y[0]; });
    if (stats) *stats = st;
    return clusters;
}

// ------------------------------------------------------------------ R3 threading
struct ThreadResult {
    std::vector<std::vector<uint32_t>> path;   // [pos][hap] global cluster ids
    double cost = 0.0;
};

inline ThreadResult thread_paths_ordered(uint32_t ploidy, double switchCost, double affineSwitchCost,
                                 uint32_t start, uint32_t end,
                                 const std::vector<std::vector<uint32_t>>& covMap,
                                 const std::vector<std::vector<double>>& coverage,
                                 const std::vector<std::vector<uint32_t>>& consensus,
                                 const std::vector<std::unordered_map<uint32_t, uint32_t>>& genotypes) {
    ThreadResult res;
    if (end <= start) return res;
    const uint32_t p = ploidy;
    const uint32_t L = end - start;
    std::vector<std::vector<double>> D(L);
    std::vector<std::vector<int64_t>> back(L);
    std::vector<uint32_t> t(p), s(p);
    auto n_states = [&](uint32_t k) { uint64_t S = 1; for (uint32_t h = 0; h < p; h++) S *= k; return S; };
    auto decode = [&](uint64_t code, uint32_t k, std::vector<uint32_t>& out) {
        for (int h = (int)p - 1; h >= 0; h--) { out[h] = (uint32_t)(code % k); code /= k; }
    };
    const double INF = std::numeric_limits<double>::infinity();
    for (uint32_t q = 0; q < L; q++) {
        const uint32_t pos = start + q;
        const uint32_t k = (uint32_t)covMap[pos].size();
        const uint64_t S = n_states(k);
        D[q].assign(S, INF); back[q].assign(S, -1);
        // conformity
        std::vector<char> allowed(S, 0); bool any = false;
        const bool have_gt = pos < genotypes.size();
        for (uint64_t c = 0; c < S; c++) {
            decode(c, k, t);
            bool ok;
            if (have_gt && !genotypes[pos].empty()) {
                std::unordered_map<uint32_t, uint32_t> cnt;
                for (uint32_t h = 0; h < p; h++) cnt[consensus[pos][t[h]]]++;
                ok = cnt.size() == genotypes[pos].size();
                if (ok) for (auto& kv : cnt) { auto it = genotypes[pos].find(kv.first); if (it == genotypes[pos].end() || it->second != kv.second) { ok = false; break; } }
            } else {
                ok = false;
                for (uint32_t h = 1; h < p; h++) if (consensus[pos][t[h]] != consensus[pos][t[0]]) ok = true;
            }
            allowed[c] = ok; any |= ok;
        }
        if (!any) std::fill(allowed.begin(), allowed.end(), 1);
        const uint32_t kprev = q ? (uint32_t)covMap[pos - 1].size() : 0;
        const uint64_t Sprev = q ? n_states(kprev) : 0;
        for (uint64_t c = 0; c < S; c++) {
            if (!allowed[c]) continue;
            decode(c, k, t);
            double cc = 0.0;
            for (uint32_t h = 0; h < p; h++) {
                uint32_t m = 0; for (uint32_t g = 0; g < p; g++) if (t[g] == t[h]) m++;
                double cov = coverage[pos][t[h]];
                double lo = (2.0 * m - 1.0) / (2.0 * p), hi = (2.0 * m + 1.0) / (2.0 * p);
                if (cov < lo || cov > hi) cc += 1.0;
            }
            if (q == 0) { D[q][c] = cc; continue; }
            double best = INF; int64_t arg = -1;
            for (uint64_t d = 0; d < Sprev; d++) {
                if (D[q - 1][d] == INF) continue;
                decode(d, kprev, s);
                uint32_t sw = 0;
                for (uint32_t h = 0; h < p; h++) if (covMap[pos - 1][s[h]] != covMap[pos][t[h]]) sw++;
                double v = D[q - 1][d] + switchCost * sw + (sw ? affineSwitchCost : 0.0);
                if (v < best) { best = v; arg = (int64_t)d; }
            }
            D[q][c] = best + cc; back[q][c] = arg;
        }
    }
    // backtrace
    uint64_t cur = 0; double best = INF;
    for (uint64_t c = 0; c < D[L - 1].size(); c++) if (D[L - 1][c] < best) { best = D[L - 1][c]; cur = c; }
    res.cost = best;
    res.path.assign(L, std::vector<uint32_t>(p));
    for (int q = (int)L - 1; q >= 0; q--) {
        const uint32_t pos = start + q;
        decode(cur, (uint32_t)covMap[pos].size(), t);
        for (uint32_t h = 0; h < p; h++) res.path[q][h] = covMap[pos][t[h]];
        if (q > 0) cur = (uint64_t)back[q][cur];
    }
    return res;
}

// ------------------------------------------------------------------ R3c threading on canonical tuples (ploidy > 4)
struct Canon {
    // N(j, k) = number of non-decreasing j-tuples over k symbols = C(k + j - 1, j)
    static int64_t N(int j, int k) {
        if (j == 0) return 1;
        if (k <= 0) return 0;
        int64_t r = 1;
        for (int i = 1; i <= j; i++) r = r * (k - 1 + i) / i;
        return r;
    }
    static int rank(const uint8_t* x, int j, int k) {          // lexicographic rank of a non-decreasing tuple
        int64_t r = 0; int prev = 0;
        for (int i = 0; i < j; i++) { for (int v = prev; v < x[i]; v++) r += N(j - 1 - i, k - v); prev = x[i]; }
        return (int)r;
    }
    static std::vector<std::array<uint8_t, 8>> tuples(int j, int k) {      // in rank order
        std::vector<std::array<uint8_t, 8>> out;
        std::array<uint8_t, 8> t{};
        if (j == 0) { out.push_back(t); return out; }
        if (k <= 0) return out;
        std::vector<int> x(j, 0);
        while (true) {
            for (int i = 0; i < j; i++) t[i] = (uint8_t)x[i];
            out.push_back(t);
            int i = j - 1;
            while (i >= 0 && x[i] == k - 1) i--;
            if (i < 0) break;
            const int v = x[i] + 1;
            for (int u = i; u < j; u++) x[u] = v;
        }
        return out;
    }
};

static const int CANON_INF = 1 << 29;

// per-state column cost: coverage cost if the tuple is allowed, -1 - cost if it does not conform (R3's rules)
inline void canon_column(uint32_t p, uint32_t pos, int k, const std::vector<std::array<uint8_t, 8>>& T,
                         const std::vector<std::vector<double>>& coverage, const std::vector<std::vector<uint32_t>>& consensus,
                         const std::vector<std::unordered_map<uint32_t, uint32_t>>& genotypes, std::vector<int>& cc, bool& any) {
    (void)k;
    cc.assign(T.size(), 0); any = false;
    const bool have_gt = pos < genotypes.size() && !genotypes[pos].empty();
    for (size_t c = 0; c < T.size(); c++) {
        const uint8_t* t = T[c].data();
        bool ok;
        if (have_gt) {
            std::unordered_map<uint32_t, uint32_t> cnt;
            for (uint32_t h = 0; h < p; h++) cnt[consensus[pos][t[h]]]++;
            ok = cnt.size() == genotypes[pos].size();
            if (ok) for (auto& kv : cnt) { auto it = genotypes[pos].find(kv.first); if (it == genotypes[pos].end() || it->second != kv.second) { ok = false; break; } }
        } else {
            ok = false;
            for (uint32_t h = 1; h < p; h++) if (consensus[pos][t[h]] != consensus[pos][t[0]]) ok = true;
        }
        int cost = 0;
        for (uint32_t h = 0; h < p; h++) {
            uint32_t m = 0; for (uint32_t g = 0; g < p; g++) if (t[g] == t[h]) m++;
            const double cov = coverage[pos][t[h]];
            const double lo = (2.0 * m - 1.0) / (2.0 * p), hi = (2.0 * m + 1.0) / (2.0 * p);
            if (cov < lo || cov > hi) cost++;
        }
        cc[c] = ok ? cost : -1 - cost;
        any |= ok;
    }
}

inline ThreadResult canon_backtrace(uint32_t p, uint32_t start, uint32_t L, const std::vector<std::vector<uint32_t>>& covMap,
                                    const std::vector<int>& kq, const std::vector<std::vector<std::array<uint8_t, 8>>>& tup,
                                    const std::vector<int>& Dlast, const std::vector<std::vector<int32_t>>& back) {
    ThreadResult res;
    int best = INT32_MAX; int cur = 0;
    for (size_t c = 0; c < Dlast.size(); c++) if (Dlast[c] < best) { best = Dlast[c]; cur = (int)c; }
    res.cost = (double)best;
    std::vector<int> state(L);
    for (int q = (int)L - 1; q >= 0; q--) { state[q] = cur; if (q > 0) cur = back[q][cur]; }
    res.path.assign(L, std::vector<uint32_t>(p));
    for (uint32_t q = 0; q < L; q++) {
        const uint32_t pos = start + q;
        const uint8_t* t = tup[q][state[q]].data();
        (void)kq;
        if (q == 0) { for (uint32_t h = 0; h < p; h++) res.path[0][h] = covMap[pos][t[h]]; continue; }
        std::vector<char> claimed(p, 0), placed(p, 0);
        for (uint32_t h = 0; h < p; h++)                       // keep the cluster while an unclaimed copy is left
            for (uint32_t e = 0; e < p; e++) if (!claimed[e] && covMap[pos][t[e]] == res.path[q - 1][h]) { claimed[e] = 1; placed[h] = 1; res.path[q][h] = res.path[q - 1][h]; break; }
        uint32_t e = 0;
        for (uint32_t h = 0; h < p; h++) if (!placed[h]) { while (claimed[e]) e++; res.path[q][h] = covMap[pos][t[e]]; claimed[e] = 1; }
    }
    return res;
}

// the definition: all pairs of states
inline ThreadResult thread_paths_canonical_direct(uint32_t p, double switchCost, double affineSwitchCost, uint32_t start, uint32_t end,
                                                  const std::vector<std::vector<uint32_t>>& covMap, const std::vector<std::vector<double>>& coverage,
                                                  const std::vector<std::vector<uint32_t>>& consensus,
                                                  const std::vector<std::unordered_map<uint32_t, uint32_t>>& genotypes) {
    if (end <= start) return ThreadResult();
    const uint32_t L = end - start; const int sw = (int)switchCost, aff = (int)affineSwitchCost;
    std::vector<int> kq(L); std::vector<std::vector<std::array<uint8_t, 8>>> tup(L);
    std::vector<std::vector<int32_t>> back(L);
    std::vector<int> Dprev, Dcur, cc;
    for (uint32_t q = 0; q < L; q++) {
        const uint32_t pos = start + q;
        kq[q] = (int)std::min<size_t>(covMap[pos].size(), p + 2);
        tup[q] = Canon::tuples((int)p, kq[q]);
        const size_t S = tup[q].size();
        bool any; canon_column(p, pos, kq[q], tup[q], coverage, consensus, genotypes, cc, any);
        Dcur.assign(S, CANON_INF); back[q].assign(S, -1);
        std::vector<std::vector<uint32_t>> G(S), Gp;
        for (size_t c = 0; c < S; c++) { for (uint32_t h = 0; h < p; h++) G[c].push_back(covMap[pos][tup[q][c][h]]); std::sort(G[c].begin(), G[c].end()); }
        if (q > 0) { Gp.resize(tup[q - 1].size()); for (size_t c = 0; c < Gp.size(); c++) { for (uint32_t h = 0; h < p; h++) Gp[c].push_back(covMap[pos - 1][tup[q - 1][c][h]]); std::sort(Gp[c].begin(), Gp[c].end()); } }
        for (size_t c = 0; c < S; c++) {
            const bool allowed = cc[c] >= 0 || !any;
            const int cost = cc[c] >= 0 ? cc[c] : -1 - cc[c];
            if (!allowed) continue;
            if (q == 0) { Dcur[c] = cost; continue; }
            int best = INT32_MAX; int arg = -1;
            for (size_t d = 0; d < Gp.size(); d++) {
                size_t x = 0, y = 0; int common = 0;
                while (x < p && y < p) { if (Gp[d][x] < G[c][y]) x++; else if (Gp[d][x] > G[c][y]) y++; else { common++; x++; y++; } }
                const int dist = (int)p - common;
                const int v = Dprev[d] + sw * dist + (dist ? aff : 0);
                if (v < best) { best = v; arg = (int)d; }
            }
            Dcur[c] = std::min(best + cost, CANON_INF); back[q][c] = arg;
        }
        Dprev = Dcur;
    }
    return canon_backtrace(p, start, L, covMap, kq, tup, Dprev, back);
}

// sub-multiset DP: E_j[c] = min D[s] over the states s that contain the j-multiset c (down pass over the previous
// column's indices), G_j[c] = min over c' inside c of E[c'] + switchCost * (j - |c'|) (up pass over this column's
// indices; a multiset takes part as c' only if all its clusters exist in the previous column).  Values are
// (cost, predecessor index) pairs under lexicographic minimum, so ties resolve to the lowest index as in the definition.
inline ThreadResult thread_paths_canonical(uint32_t p, double switchCost, double affineSwitchCost, uint32_t start, uint32_t end,
                                           const std::vector<std::vector<uint32_t>>& covMap, const std::vector<std::vector<double>>& coverage,
                                           const std::vector<std::vector<uint32_t>>& consensus,
                                           const std::vector<std::unordered_map<uint32_t, uint32_t>>& genotypes) {
    if (end <= start) return ThreadResult();
    const uint32_t L = end - start; const int sw = (int)switchCost, aff = (int)affineSwitchCost;
    typedef std::pair<int, int> VA;                              // (value, predecessor)
    const VA NONE(INT32_MAX, INT32_MAX);
    std::vector<int> kq(L); std::vector<std::vector<std::array<uint8_t, 8>>> tup(L);
    std::vector<std::vector<int32_t>> back(L);
    std::vector<int> Dprev, Dcur, cc;
    for (uint32_t q = 0; q < L; q++) {
        const uint32_t pos = start + q;
        const int kc = kq[q] = (int)std::min<size_t>(covMap[pos].size(), p + 2);
        tup[q] = Canon::tuples((int)p, kc);
        const size_t S = tup[q].size();
        bool any; canon_column(p, pos, kc, tup[q], coverage, consensus, genotypes, cc, any);
        Dcur.assign(S, CANON_INF); back[q].assign(S, -1);
        if (q == 0) { for (size_t c = 0; c < S; c++) if (cc[c] >= 0 || !any) Dcur[c] = cc[c] >= 0 ? cc[c] : -1 - cc[c]; Dprev = Dcur; continue; }
        const int kp = kq[q - 1];
        // down pass
        std::vector<std::vector<VA>> E(p + 1);
        E[p].resize(Dprev.size());
        for (size_t s = 0; s < Dprev.size(); s++) E[p][s] = VA(Dprev[s], (int)s);
        for (int j = (int)p - 1; j >= 0; j--) {
            auto Tj = Canon::tuples(j, kp);
            E[j].assign(Tj.size(), NONE);
            for (size_t c = 0; c < Tj.size(); c++)
                for (int g = 0; g < kp; g++) {
                    uint8_t x[8]; int w = 0; bool done = false;
                    for (int i = 0; i < j; i++) { if (!done && g < Tj[c][i]) { x[w++] = (uint8_t)g; done = true; } x[w++] = Tj[c][i]; }
                    if (!done) x[w++] = (uint8_t)g;
                    E[j][c] = std::min(E[j][c], E[j + 1][Canon::rank(x, j + 1, kp)]);
                }
        }
        // this column's local index -> previous column's local index of the same global cluster
        std::vector<int> mp(kc, -1);
        for (int l = 0; l < kc; l++) for (int m = 0; m < kp; m++) if (covMap[pos - 1][m] == covMap[pos][l]) { mp[l] = m; break; }
        auto mapped = [&](const uint8_t* c, int j, VA& out) {
            uint8_t x[8];
            for (int i = 0; i < j; i++) { if (mp[c[i]] < 0) return false; x[i] = (uint8_t)mp[c[i]]; }
            std::sort(x, x + j);
            out = E[j][Canon::rank(x, j, kp)];
            return true;
        };
        // up pass
        std::vector<std::vector<VA>> G(p + 1);
        G[0].assign(1, E[0][0]);
        for (int j = 1; j <= (int)p; j++) {
            auto Tj = Canon::tuples(j, kc);
            G[j].assign(Tj.size(), NONE);
            for (size_t c = 0; c < Tj.size(); c++) {
                VA best = NONE, m;
                if (mapped(Tj[c].data(), j, m)) best = m;
                for (int i = 0; i < j; i++) {
                    if (i > 0 && Tj[c][i] == Tj[c][i - 1]) continue;
                    uint8_t x[8]; int w = 0;
                    for (int u = 0; u < j; u++) if (u != i) x[w++] = Tj[c][u];
                    VA g = G[j - 1][Canon::rank(x, j - 1, kc)];
                    if (g.first < CANON_INF) g.first += sw; else g = NONE;
                    best = std::min(best, g);
                }
                G[j][c] = best;
            }
        }
        for (size_t c = 0; c < S; c++) {
            const bool allowed = cc[c] >= 0 || !any;
            const int cost = cc[c] >= 0 ? cc[c] : -1 - cc[c];
            if (!allowed) continue;
            VA a = G[p][c];
            if (a.first < CANON_INF) a.first += aff; else a = NONE;
            VA b;
            if (mapped(tup[q][c].data(), (int)p, b)) a = std::min(a, b);
            Dcur[c] = std::min(a.first + cost, CANON_INF); back[q][c] = a.second;
        }
        Dprev = Dcur;
    }
    return canon_backtrace(p, start, L, covMap, kq, tup, Dprev, back);
}

// computePaths: rule R3 on ordered tuples up to ploidy 4 (the reference runs ploidy 2), rule R3c above
inline ThreadResult thread_paths(uint32_t ploidy, double switchCost, double affineSwitchCost, uint32_t start, uint32_t end,
                                 const std::vector<std::vector<uint32_t>>& covMap, const std::vector<std::vector<double>>& coverage,
                                 const std::vector<std::vector<uint32_t>>& consensus,
                                 const std::vector<std::unordered_map<uint32_t, uint32_t>>& genotypes) {
    if (ploidy > 4) return thread_paths_canonical(ploidy, switchCost, affineSwitchCost, start, end, covMap, coverage, consensus, genotypes);
    return thread_paths_ordered(ploidy, switchCost, affineSwitchCost, start, end, covMap, coverage, consensus, genotypes);
}

}  // namespace ahs_oracle
