// oracle/phase_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's per-chain phasing path on the flattened CSR batch of
// include/ahsoka_b200.h.  Ahsoka-owned stages follow reference src/alignmentstoreadset.cpp
// literally (line numbers cited at each step); the three WhatsHap-held algorithms come from
// oracle/core/phase_core.hpp (PARITY UNPINNED, see its header).
//
// Pinning: this file is checked against the reference sources compiled VERBATIM with the
// WhatsHap API shim (oracle/_ref/Ahsoka_ref) by tests/test_cli_parity.py — byte-identical
// -result.txt — which pins every Ahsoka-owned stage (projection, filter, ordering,
// coverage, consensus, re-packing, emission) to the reference text itself.
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline, --impl reference) may
// load this library.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/ahsoka_b200.h"
#include "core/phase_core.hpp"

#include <atomic>
#include <thread>

using namespace ahs_oracle;

namespace {

struct ORead {                       // a WhatsHap `Read` reduced to what the path uses
    int32_t id;                      // chain-local read index (= name)
    int32_t mapq;
    std::vector<std::pair<int32_t, int32_t>> var;   // (position, allele), kept sorted
    bool has(int32_t pos) const { for (auto& v : var) if (v.first == pos) return true; return false; }
    void add(int32_t pos, int32_t allele) { var.push_back({pos, allele}); std::sort(var.begin(), var.end()); }
};
struct OReadSet {
    std::vector<ORead> reads;                       // insertion order
    std::unordered_map<int32_t, int32_t> index;     // name -> slot
    ORead* by_name(int32_t id) { auto it = index.find(id); return it == index.end() ? nullptr : &reads[it->second]; }
    void add(const ORead& r) { index[r.id] = (int32_t)reads.size(); reads.push_back(r); }
};

struct ChainOut {
    int32_t status = AHS_CHAIN_OK, n_clusters = 0, maxpos = -1;
    std::vector<int32_t> read_id, read_mapq, read_cluster;
    std::vector<int64_t> cell_cnt;
    std::vector<int32_t> cell_pos; std::vector<uint8_t> cell_allele;
    std::vector<int32_t> pos, path; std::vector<uint8_t> hap_allele;
    double dp_cost = 0.0;
    int64_t n_pairs = 0;
};

// is_subset, reference src/alignmentstoreadset.cpp:495-548
bool is_subset(std::vector<int32_t> allelepath, const std::vector<int32_t>& alignment_sorted, bool take_partial) {
    if (take_partial) {
        if (allelepath.empty()) return false;          // reference: pop_back on empty = UB; never generated
        allelepath.pop_back();
        if (!allelepath.empty()) allelepath.erase(allelepath.begin(), allelepath.begin() + 1);   // :510-511 (UB when empty)
    }
    std::sort(allelepath.begin(), allelepath.end());
    return std::includes(alignment_sorted.begin(), alignment_sorted.end(), allelepath.begin(), allelepath.end());
}

bool cmp_desc(const std::pair<double, double>& a, const std::pair<double, double>& b) { return a.second > b.second; }   // :699-701
bool cmp_asc(const std::pair<double, double>& a, const std::pair<double, double>& b) { return a.second < b.second; }    // :703-705

void phase_chain(const ahs_batch_in* in, int c, ChainOut& o) {
    const int ploidy = in->ploidy;
    const int64_t gb0 = in->bubble_off[c];
    const int B = (int)(in->bubble_off[c + 1] - gb0);
    if (B <= 1) { o.status = AHS_CHAIN_TRIVIAL; return; }                        // :86
    const int64_t e0 = in->entry_off[c], e1 = in->entry_off[c + 1];
    const int E = (int)(e1 - e0);
    // sorted node lists per entry (is_subset sorts a copy for every test, :496-497)
    std::vector<std::vector<int32_t>> enodes(E);
    for (int e = 0; e < E; e++) {
        enodes[e].assign(in->enode + in->enode_off[e0 + e], in->enode + in->enode_off[e0 + e + 1]);
        std::sort(enodes[e].begin(), enodes[e].end());
    }
    auto allele_count = [&](int b) { return (b >= 0 && b < B) ? (int)(in->allele_off[gb0 + b + 1] - in->allele_off[gb0 + b]) : 0; };   // :216 default-inserts an empty list
    auto allele_path = [&](int b, int a) {
        int64_t ga = in->allele_off[gb0 + b] + a;
        return std::vector<int32_t>(in->anode + in->anode_off[ga], in->anode + in->anode_off[ga + 1]);
    };
    auto mapq_of = [&](int e) { return (int32_t)(in->entry_identity[e0 + e] * 100); };     // float*int -> float -> int, :117,:234

    // ---- stage A, :90-135
    OReadSet readset;
    for (int ob = 0; ob < B; ob++) {
        const int bubbleid = in->stage_a_order ? in->stage_a_order[gb0 + ob] : (B - 1 - ob);
        for (int it = 0; it < allele_count(bubbleid); it++) {
            std::vector<int32_t> ap = allele_path(bubbleid, it);
            for (int e = 0; e < E; e++) {
                if (!is_subset(ap, enodes[e], false)) continue;
                const int32_t readid = in->entry_read[e0 + e];
                ORead* r = readset.by_name(readid);
                if (!r) { ORead nr; nr.id = readid; nr.mapq = mapq_of(e); nr.add(bubbleid, it); readset.add(nr); }
                else if (!r->has(bubbleid)) r->add(bubbleid, it);
            }
        }
    }
    if (readset.reads.empty()) { o.status = AHS_CHAIN_EMPTY; return; }           // reference: .back() on empty vector (UB), :193
    // filter, :146-163
    std::vector<const ORead*> testset;
    for (auto& r : readset.reads) if ((int)r.var.size() > 1 && r.mapq >= 93) testset.push_back(&r);
    // boundaries, :173-189
    std::unordered_set<int> firstp, lastp;
    for (auto* r : testset) { firstp.insert(r->var.front().first); lastp.insert(r->var.back().first); }
    std::set<int> to_be_added;
    for (int el : lastp) if (!firstp.count(el)) { to_be_added.insert(el); to_be_added.insert(el + 1); }
    int maxpos = -1;
    for (auto& r : readset.reads) maxpos = std::max(maxpos, r.var.back().first);    // get_positions().back(), :193,:206
    o.maxpos = maxpos;
    for (int i = 0; i < maxpos; i++) to_be_added.insert(i);                       // gaps are a subset, :192-208

    // ---- stage B, :210-254
    OReadSet partial;
    for (int boundary : to_be_added) {
        for (int it = 0; it < allele_count(boundary); it++) {
            std::vector<int32_t> ap = allele_path(boundary, it);
            for (int e = 0; e < E; e++) {
                if (!is_subset(ap, enodes[e], true)) continue;
                const int32_t readid = in->entry_read[e0 + e];
                ORead* r = partial.by_name(readid);
                if (!r) { ORead nr; nr.id = readid; nr.mapq = mapq_of(e); nr.add(boundary, it); partial.add(nr); }
                else if (!r->has(boundary) && (in->entry_identity[e0 + e] * 100) > 90) r->add(boundary, it);   // :245
            }
        }
    }
    // filter, :262-274
    std::vector<ORead> fin;
    for (auto& r : partial.reads) if ((int)r.var.size() > 1 && r.mapq >= 93) fin.push_back(r);
    if (fin.empty()) { o.status = AHS_CHAIN_EMPTY; return; }                      // :279-282
    // ReadSet::sort(), :297 — std::sort, comparator on firstPosition() only
    std::sort(fin.begin(), fin.end(), [](const ORead& a, const ORead& b) { return a.var.front().first < b.var.front().first; });
    const int R = (int)fin.size();
    // get_positions(), :317,:346
    std::set<int32_t> posset;
    for (auto& r : fin) for (auto& v : r.var) posset.insert(v.first);
    std::vector<int32_t> pos(posset.begin(), posset.end());
    const int n_pos = (int)pos.size();
    const int lastpos = pos.back();
    std::vector<int32_t> compact(lastpos + 1, -1);
    for (int i = 0; i < n_pos; i++) compact[pos[i]] = i;

    // ---- scoring + cluster editing, :308-315
    std::vector<Row> rows(R);
    for (int i = 0; i < R; i++) for (auto& v : fin[i].var) { rows[i].pos.push_back(v.first); rows[i].allele.push_back(v.second); }
    std::vector<PairScore> ps;
    score_reads_local(rows, 1, (uint32_t)ploidy, ps);
    o.n_pairs = (int64_t)ps.size();
    std::vector<std::vector<int32_t>> clusters = cluster_edit(R, ps);
    const int n_clusters = (int)clusters.size();
    std::vector<int32_t> read_cluster(R, -1);
    for (int ci = 0; ci < n_clusters; ci++) for (int r : clusters[ci]) read_cluster[r] = ci;

    // ---- get_coverage, :660-697
    std::vector<std::map<double, double>> coverage(lastpos + 1);
    std::vector<double> coverage_sum(lastpos + 1, 0.0);
    for (int ci = 0; ci < n_clusters; ci++) for (int r : clusters[ci]) for (auto& v : fin[r].var) {
        coverage[v.first][(double)ci] += 1; coverage_sum[v.first] += 1;
    }
    for (int i = 0; i <= lastpos; i++) for (auto& kv : coverage[i]) kv.second = kv.second / coverage_sum[i];     // :694
    // ---- get_pos_to_clusters_map, :751-779 (std::sort with cmp on .second, :707-727)
    std::vector<std::vector<uint32_t>> covMap;
    for (int p = 0; p <= lastpos; p++) {
        if (compact[p] == -1) continue;
        std::vector<std::pair<double, double>> A(coverage[p].begin(), coverage[p].end());
        std::sort(A.begin(), A.end(), cmp_desc);
        std::vector<uint32_t> sorted_cids; for (auto& a : A) sorted_cids.push_back((uint32_t)(int)a.first);
        size_t cut_off = std::min<uint32_t>((uint32_t)sorted_cids.size(), 2u * ploidy);
        for (uint32_t i = (uint32_t)ploidy; i < std::min<uint32_t>((uint32_t)sorted_cids.size(), 2u * ploidy); i++)
            if (coverage[p][(double)sorted_cids[i]] < (1.0 / (8.0 * ploidy))) { cut_off = i; break; }      // :768
        covMap.emplace_back(sorted_cids.begin(), sorted_cids.begin() + cut_off);
    }
    // ---- get_local_cluster_consensus / get_single_cluster_consensus_frac, :550-655
    std::vector<std::vector<uint32_t>> rel_pos(n_clusters);
    for (int q = 0; q < n_pos; q++) for (uint32_t cl : covMap[q]) rel_pos[cl].push_back((uint32_t)q);
    std::vector<std::map<uint32_t, uint32_t>> cons_by_cluster(n_clusters);     // cluster -> compact pos -> allele
    for (int ci = 0; ci < n_clusters; ci++) {
        std::vector<std::map<double, double>> poswise(n_pos);
        for (int r : clusters[ci]) for (auto& v : fin[r].var) poswise[compact[v.first]][(double)v.second] += 1;
        for (uint32_t q : rel_pos[ci]) {
            if (!poswise[q].empty()) {
                std::vector<std::pair<double, double>> A(poswise[q].begin(), poswise[q].end());
                std::sort(A.begin(), A.end(), cmp_asc);                      // sort_asc, :729-749
                int max_allele = 0; double max_count = 0;
                for (auto& a : A) if (a.second > max_count) { max_allele = (int)a.first; max_count = a.second; }   // :637-644
                cons_by_cluster[ci][q] = (uint32_t)max_allele;
            } else cons_by_cluster[ci][q] = 0;                               // :647-649
        }
    }
    std::vector<std::map<uint32_t, uint32_t>> new_consensus(n_pos);            // compact pos -> cluster -> allele
    for (int q = 0; q < n_pos; q++) for (uint32_t cl : covMap[q]) new_consensus[q][cl] = cons_by_cluster[cl][(uint32_t)q];
    // ---- re-packing, :378-402 (ascending cluster id, NOT covMap order: SURVEY A#12)
    std::vector<std::vector<double>> cov_vec(n_pos);
    for (int i = 0; i <= lastpos; i++) if (!coverage[i].empty()) for (auto& kv : coverage[i]) cov_vec[compact[i]].push_back(kv.second);
    std::vector<std::vector<uint32_t>> cons_vec(n_pos);
    for (int q = 0; q < n_pos; q++) for (auto& kv : new_consensus[q]) cons_vec[q].push_back(kv.second);
    // ---- genotypes, :340-344 (ploidy 2: {0:1,1:1}); ploidy > 2: empty map = "not all equal" (rule R3)
    std::vector<std::unordered_map<uint32_t, uint32_t>> genotypes(n_pos + 1);
    if (ploidy == 2) for (auto& g : genotypes) g = {{0, 1}, {1, 1}};
    // ---- threading, :320,:408
    ThreadResult tr = thread_paths((uint32_t)ploidy, 32.0, 8.0, 0, (uint32_t)n_pos, covMap, cov_vec, cons_vec, genotypes);

    // ---- outputs
    o.n_clusters = n_clusters; o.dp_cost = tr.cost; o.pos = pos;
    for (int i = 0; i < R; i++) {
        o.read_id.push_back(fin[i].id); o.read_mapq.push_back(fin[i].mapq); o.read_cluster.push_back(read_cluster[i]);
        o.cell_cnt.push_back((int64_t)fin[i].var.size());
        for (auto& v : fin[i].var) { o.cell_pos.push_back(v.first); o.cell_allele.push_back((uint8_t)v.second); }
    }
    o.path.resize((size_t)n_pos * ploidy); o.hap_allele.resize((size_t)n_pos * ploidy);
    for (int q = 0; q < n_pos; q++) for (int h = 0; h < ploidy; h++) {
        uint32_t cid = tr.path[q][h];
        o.path[(size_t)q * ploidy + h] = (int32_t)cid;
        o.hap_allele[(size_t)q * ploidy + h] = (uint8_t)new_consensus[q][cid];     // :420-423
    }
}

template <class T> T* dup(const std::vector<T>& v) {
    T* p = (T*)malloc(sizeof(T) * (v.size() ? v.size() : 1));
    if (!v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

}  // namespace

extern "C" {

int ahs_oracle_phase_batch(const ahs_batch_in* in, ahs_batch_out* out, int n_threads) {
    if (!in || !out || in->ploidy < 1) return AHS_ERR_ARG;
    memset(out, 0, sizeof(*out));
    const int C = in->n_chains;
    std::vector<ChainOut> co(C);
    {   // chains are independent (alignmentstoreadset.cpp:75): dynamic work queue over host threads
        int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
        if (nt < 1) nt = 1;
        if (nt > C) nt = C > 0 ? C : 1;
        std::atomic<int> next(0);
        auto worker = [&]() { for (int c = next.fetch_add(1); c < C; c = next.fetch_add(1)) phase_chain(in, c, co[c]); };
        if (nt == 1) worker();
        else { std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back(worker); for (auto& t : th) t.join(); }
    }
    std::vector<int32_t> status, read_id, read_mapq, read_cluster, cell_pos, n_clusters, pos, path, maxpos;
    std::vector<int64_t> read_off{0}, cell_off{0}, pos_off{0};
    std::vector<uint8_t> cell_allele, hap_allele; std::vector<double> dp_cost;
    int64_t n_pairs = 0, n_ok = 0;
    for (int c = 0; c < C; c++) {
        ChainOut& o = co[c];
        status.push_back(o.status); n_clusters.push_back(o.n_clusters); dp_cost.push_back(o.dp_cost); maxpos.push_back(o.maxpos);
        read_id.insert(read_id.end(), o.read_id.begin(), o.read_id.end());
        read_mapq.insert(read_mapq.end(), o.read_mapq.begin(), o.read_mapq.end());
        read_cluster.insert(read_cluster.end(), o.read_cluster.begin(), o.read_cluster.end());
        for (int64_t n : o.cell_cnt) cell_off.push_back(cell_off.back() + n);
        cell_pos.insert(cell_pos.end(), o.cell_pos.begin(), o.cell_pos.end());
        cell_allele.insert(cell_allele.end(), o.cell_allele.begin(), o.cell_allele.end());
        pos.insert(pos.end(), o.pos.begin(), o.pos.end());
        path.insert(path.end(), o.path.begin(), o.path.end());
        hap_allele.insert(hap_allele.end(), o.hap_allele.begin(), o.hap_allele.end());
        read_off.push_back((int64_t)read_id.size()); pos_off.push_back((int64_t)pos.size());
        n_pairs += o.n_pairs; if (o.status == AHS_CHAIN_OK) n_ok++;
    }
    out->n_chains = C; out->ploidy = in->ploidy;
    out->status = dup(status); out->read_off = dup(read_off); out->read_id = dup(read_id); out->read_mapq = dup(read_mapq);
    out->read_cluster = dup(read_cluster); out->cell_off = dup(cell_off); out->cell_pos = dup(cell_pos); out->cell_allele = dup(cell_allele);
    out->n_clusters = dup(n_clusters); out->pos_off = dup(pos_off); out->pos = dup(pos); out->path = dup(path);
    out->hap_allele = dup(hap_allele); out->dp_cost = dup(dp_cost); out->maxpos = dup(maxpos);
    out->n_cells = (int64_t)cell_pos.size(); out->n_pairs = n_pairs; out->n_chains_ok = n_ok;
    return AHS_OK;
}

void ahs_oracle_free_out(ahs_batch_out* o) {
    if (!o) return;
    free(o->status); free(o->read_off); free(o->read_id); free(o->read_mapq); free(o->read_cluster); free(o->cell_off);
    free(o->cell_pos); free(o->cell_allele); free(o->n_clusters); free(o->pos_off); free(o->pos); free(o->path);
    free(o->hap_allele); free(o->dp_cost); free(o->maxpos);
    memset(o, 0, sizeof(*o));
}

// Stand-alone access to the three restated algorithms, for unit tests and fixtures.
// rows: CSR (row_off[n+1], pos[], allele[]).  Output pairs: caller buffers of capacity cap.
int64_t ahs_oracle_score(int n, const int64_t* row_off, const int32_t* pos, const int32_t* allele, int ploidy,
                         int64_t cap, int32_t* pi, int32_t* pj, int32_t* pn, int32_t* pk, int32_t* pw, int32_t* es, int32_t* ed) {
    std::vector<Row> rows(n);
    for (int i = 0; i < n; i++) for (int64_t x = row_off[i]; x < row_off[i + 1]; x++) { rows[i].pos.push_back(pos[x]); rows[i].allele.push_back(allele[x]); }
    std::vector<PairScore> ps; std::vector<int32_t> ves, ved;
    score_reads_local(rows, 1, (uint32_t)ploidy, ps, &ves, &ved);
    for (int64_t x = 0; x < (int64_t)ps.size() && x < cap; x++) { pi[x] = ps[x].i; pj[x] = ps[x].j; pn[x] = ps[x].n; pk[x] = ps[x].k; pw[x] = ps[x].w; }
    if (es) for (int i = 0; i < n; i++) { es[i] = ves[i]; ed[i] = ved[i]; }
    return (int64_t)ps.size();
}

static thread_local ClusterEditStats g_last_stats;
void ahs_oracle_cluster_stats(int64_t* out3) { out3[0] = g_last_stats.steps; out3[1] = g_last_stats.merges; out3[2] = g_last_stats.forbids; }

int ahs_oracle_cluster(int n, int64_t n_pairs, const int32_t* pi, const int32_t* pj, const int32_t* pw, int paranoid, int32_t* label) {
    std::vector<PairScore> ps(n_pairs);
    for (int64_t x = 0; x < n_pairs; x++) { ps[x].i = pi[x]; ps[x].j = pj[x]; ps[x].w = pw[x]; ps[x].n = ps[x].k = 0; }
    auto cl = cluster_edit(n, ps, paranoid != 0, &g_last_stats);
    for (size_t c = 0; c < cl.size(); c++) for (int r : cl[c]) label[r] = (int32_t)c;
    return (int)cl.size();
}

// libstdc++'s own std::sort on (key, value) pairs with a comparator that looks at the key only: the behaviour
// ReadSet::sort() (reference src/alignmentstoreadset.cpp:297) and the cluster sort (:720) have in the reference binary.
void ahs_oracle_std_sort(int32_t* keys, int32_t* values, int32_t n, int descending) {
    std::vector<std::pair<int32_t, int32_t>> a(n);
    for (int i = 0; i < n; i++) a[i] = {keys[i], values[i]};
    if (descending) std::sort(a.begin(), a.end(), [](const std::pair<int32_t, int32_t>& x, const std::pair<int32_t, int32_t>& y) { return x.first > y.first; });
    else std::sort(a.begin(), a.end(), [](const std::pair<int32_t, int32_t>& x, const std::pair<int32_t, int32_t>& y) { return x.first < y.first; });
    for (int i = 0; i < n; i++) { keys[i] = a[i].first; values[i] = a[i].second; }
}

// McIlroy's adversary ("A killer adversary for quicksort", 1999) run against std::sort itself: returns keys that drive
// libstdc++'s introsort to its depth limit, i.e. into the heap-sort fall-back.
void ahs_oracle_antiqsort(int32_t n, int32_t* keys_out) {
    std::vector<int32_t> val(n, n - 1), ptr(n);
    const int32_t gas = n - 1; int32_t nsolid = 0, candidate = 0;
    for (int i = 0; i < n; i++) ptr[i] = i;
    std::sort(ptr.begin(), ptr.end(), [&](int32_t x, int32_t y) {
        if (val[x] == gas && val[y] == gas) { if (x == candidate) val[x] = nsolid++; else val[y] = nsolid++; }
        if (val[x] == gas) candidate = x; else if (val[y] == gas) candidate = y;
        return val[x] < val[y];
    });
    for (int i = 0; i < n; i++) keys_out[i] = val[i];
}

// Rule R3c two ways (the definition over all pairs of states / the sub-multiset DP) on one random instance: returns 0 if
// cost and paths agree.  n_pos columns, k[q] clusters per column drawn from `n_ids` global ids.
int ahs_oracle_canonical_selfcheck(int ploidy, int n_pos, int n_ids, uint64_t seed, double* cost_out) {
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    std::vector<std::vector<uint32_t>> covMap(n_pos), consensus(n_pos); std::vector<std::vector<double>> coverage(n_pos);
    std::vector<std::unordered_map<uint32_t, uint32_t>> genotypes(n_pos + 1);
    for (int q = 0; q < n_pos; q++) {
        const int k = 1 + (int)(rnd() % (2 * ploidy));
        std::vector<uint32_t> ids; for (int i = 0; i < n_ids; i++) ids.push_back(i);
        for (int i = 0; i < k && !ids.empty(); i++) { const size_t j = rnd() % ids.size(); covMap[q].push_back(ids[j]); ids.erase(ids.begin() + j); }
        const int tot = 1 + (int)(rnd() % 60); int left = tot;
        const size_t kk = covMap[q].size();
        for (size_t i = 0; i < std::max<size_t>(kk, 2 * ploidy); i++) {           // coverage lists every cluster at the position (A#12): at least k entries
            const int cnt = i + 1 == kk ? left : (int)(rnd() % (left + 1)); left -= std::min(left, cnt);
            coverage[q].push_back((double)cnt / tot);
        }
        for (size_t i = 0; i < kk; i++) consensus[q].push_back((uint32_t)(rnd() % 3));
    }
    ThreadResult a = thread_paths_canonical_direct(ploidy, 32.0, 8.0, 0, n_pos, covMap, coverage, consensus, genotypes);
    ThreadResult b = thread_paths_canonical(ploidy, 32.0, 8.0, 0, n_pos, covMap, coverage, consensus, genotypes);
    if (cost_out) *cost_out = a.cost;
    if (a.cost != b.cost || a.path != b.path) return 1;
    return 0;
}

int ahs_oracle_log_tables(int64_t* ln, int64_t* ln1) {
    const LogTables& T = log_tables();
    memcpy(ln, T.ln, sizeof(T.ln)); memcpy(ln1, T.ln1, sizeof(T.ln1));
    return 1025;
}

}  // extern "C"
