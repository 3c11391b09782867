// synth.cpp — seeded synthetic bubble-chain graphs and long reads (host only, no CUDA).
//
// Workload definition for BASELINE.json's configs: per chain
//   flank - a_0 - {K_b inner unitigs} - a_1 - ... - a_B - flank
// p truth haplotypes, reads of log-normal span sampled from one haplotype with substitution
// errors and missing cells, identity = 0.90 + 0.10 u.  One SplitMix64 stream per batch.
//
// Emits (a) the CSR batch the C ABI consumes (include/ahsoka_b200.h) and (b) the same
// instance as GFA + GAF text obeying the reference's parsers (reference src/graph.cpp:188-249,
// src/alignmentreader.cpp:84-135: S lines before L lines, both orientations of every link,
// node names utg%07dl, 16 whitespace-separated GAF tokens with id:f:<x> as the 16th).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ahsoka_b200.h"

extern "C" {

typedef struct ahs_synth_params {
    int32_t  ploidy;
    int32_t  n_chains;
    int32_t  len_mode;       /* 0 fixed mean_len, 1 uniform in [mean/2, 3*mean/2], 2 Zipf(alpha) clipped to [min_len,max_len] */
    int32_t  mean_len;
    int32_t  min_len, max_len;
    double   zipf_alpha;
    int32_t  n_forced_max;   /* len_mode 2: this many chains get max_len */
    double   depth;          /* reads per bubble */
    double   mean_span;      /* s-bar = 16 */
    double   span_sigma;     /* 0.5 */
    double   err, miss;      /* 0.05, 0.02 */
    int32_t  max_alleles;    /* K_b uniform in [2, max_alleles]; 2 for diploid */
    int32_t  dup_lines;      /* per mille of reads that get a second GAF line (SURVEY A#3) */
    uint64_t seed;
} ahs_synth_params;

struct ahs_synth_batch;
int   ahs_synth_generate(const ahs_synth_params *p, ahs_synth_batch **out);
const ahs_batch_in *ahs_synth_batch_in(const ahs_synth_batch *b);
int64_t ahs_synth_counts(const ahs_synth_batch *b, int what);
const int32_t *ahs_synth_truth_read_hap(const ahs_synth_batch *b);
const uint8_t *ahs_synth_truth_hap_allele(const ahs_synth_batch *b);
int   ahs_synth_write_gfa_gaf(const ahs_synth_batch *b, const char *gfa_path, const char *gaf_path);
void  ahs_synth_free(ahs_synth_batch *b);
}

namespace {
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double u() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    double normal() { double a = u(), b = u(); if (a < 1e-300) a = 1e-300; return std::sqrt(-2.0 * std::log(a)) * std::cos(6.283185307179586 * b); }
};
}  // namespace

struct ahs_synth_batch {
    ahs_synth_params prm;
    std::vector<int32_t> chain_id;
    std::vector<int64_t> bubble_off, allele_off, anode_off, read_off, entry_off, enode_off;
    std::vector<int32_t> anode, stage_a_order, enode, entry_read;
    std::vector<float> entry_identity;
    std::vector<std::string> entry_identity_txt;
    std::vector<uint8_t> entry_reverse;
    // truth
    std::vector<int32_t> read_hap;        // per entry
    std::vector<uint8_t> hap_allele;      // [n_bubbles * ploidy]
    // graph description for the GFA writer
    std::vector<int32_t> chain_flank_l, chain_flank_r;
    std::vector<int32_t> anchor_off;      // [C+1] anchors per chain = B+1
    std::vector<int32_t> anchors;         // node ids
    ahs_batch_in view;
    int64_t n_nodes = 0;
};

extern "C" int ahs_synth_generate(const ahs_synth_params *pp, ahs_synth_batch **out) {
    if (!pp || !out || pp->ploidy < 1 || pp->n_chains < 0) return AHS_ERR_ARG;
    ahs_synth_batch *b = new ahs_synth_batch();
    b->prm = *pp;
    const ahs_synth_params &P = b->prm;
    Rng rng(P.seed);
    const int C = P.n_chains, p = P.ploidy;
    // chain lengths
    std::vector<int32_t> len(C);
    for (int c = 0; c < C; c++) {
        int L;
        if (P.len_mode == 0) L = P.mean_len;
        else if (P.len_mode == 1) { int lo = P.mean_len / 2, hi = P.mean_len + P.mean_len / 2; L = lo + (int)rng.below((uint32_t)(hi - lo + 1)); }
        else {
            if (c < P.n_forced_max) L = P.max_len;
            else {  // inverse-CDF Zipf over [min_len, max_len]
                double a = P.zipf_alpha, u = rng.u();
                double lo = std::pow((double)P.min_len, 1.0 - a), hi = std::pow((double)P.max_len + 1.0, 1.0 - a);
                L = (int)std::floor(std::pow(lo + u * (hi - lo), 1.0 / (1.0 - a)));
            }
        }
        if (L < P.min_len) L = P.min_len;
        if (P.max_len > 0 && L > P.max_len) L = P.max_len;
        len[c] = L;
    }
    // largest first, ties by chain id descending: the reference's size_sorting (polyassembly.cpp:136-140)
    std::vector<int32_t> order(C);
    for (int c = 0; c < C; c++) order[c] = c;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return len[x] != len[y] ? len[x] > len[y] : x > y; });

    int32_t next_node = 1;
    b->bubble_off.push_back(0); b->allele_off.push_back(0); b->anode_off.push_back(0);
    b->read_off.push_back(0); b->entry_off.push_back(0); b->enode_off.push_back(0); b->anchor_off.push_back(0);
    std::vector<int32_t> inner_first;   // per bubble: first inner node id
    std::vector<uint8_t> kb;            // per bubble: allele count
    for (int oc = 0; oc < C; oc++) {
        const int c = order[oc];
        const int B = len[c];
        b->chain_id.push_back(c);
        const int64_t b0 = (int64_t)kb.size();
        b->chain_flank_l.push_back(next_node++);
        for (int x = 0; x < B; x++) {
            b->anchors.push_back(next_node++);
            int K = 2;
            if (P.max_alleles > 2) K = 2 + (int)rng.below((uint32_t)(P.max_alleles - 1));
            kb.push_back((uint8_t)K);
            inner_first.push_back(next_node);
            next_node += K;
        }
        b->anchors.push_back(next_node++);
        b->chain_flank_r.push_back(next_node++);
        b->anchor_off.push_back((int32_t)b->anchors.size());
        const int32_t *anc = b->anchors.data() + b->anchor_off[oc];
        // allele paths: K==2 -> [source, inner, sink] (findPathsSimple, chainstoreadset.cpp:17-30);
        // otherwise [sink, inner, source] (DFS from the sink, chainstoreadset.cpp:84-116).
        for (int x = 0; x < B; x++) {
            int K = kb[b0 + x];
            for (int a = 0; a < K; a++) {
                int32_t src = anc[x], snk = anc[x + 1], in = inner_first[b0 + x] + a;
                if (K == 2) { b->anode.push_back(src); b->anode.push_back(in); b->anode.push_back(snk); }
                else        { b->anode.push_back(snk); b->anode.push_back(in); b->anode.push_back(src); }
                b->anode_off.push_back((int64_t)b->anode.size());
            }
            b->allele_off.push_back((int64_t)b->anode_off.size() - 1);
        }
        for (int x = B - 1; x >= 0; x--) b->stage_a_order.push_back(x);
        b->bubble_off.push_back((int64_t)kb.size());
        // truth haplotypes
        const size_t h0 = b->hap_allele.size();
        b->hap_allele.resize(h0 + (size_t)B * p);
        std::vector<int> perm(p);
        for (int x = 0; x < B; x++) {
            int K = kb[b0 + x];
            for (int h = 0; h < p; h++) perm[h] = h;
            for (int h = p - 1; h > 0; h--) { int r = (int)rng.below((uint32_t)(h + 1)); std::swap(perm[h], perm[r]); }
            for (int h = 0; h < p; h++) {
                int a = (h < K && h < p) ? h : (int)rng.below((uint32_t)K);
                b->hap_allele[h0 + (size_t)x * p + perm[h]] = (uint8_t)a;
            }
        }
        // reads
        const int R = B >= 1 ? (int)std::ceil(P.depth * B / P.mean_span) : 0;
        int n_reads_chain = 0;
        for (int r = 0; r < R; r++) {
            int span = (int)std::lround(std::exp(std::log(P.mean_span) + P.span_sigma * rng.normal()));
            if (span < 2) span = 2;
            if (span > B) span = B;
            int start = (int)rng.below((uint32_t)(B - span + 1));
            int hap = (int)rng.below((uint32_t)p);
            bool rev = rng.u() < 0.5;
            double ident = 0.90 + 0.10 * rng.u();
            int n_lines = 1;
            if (P.dup_lines > 0 && (int)rng.below(1000) < P.dup_lines && span >= 4) n_lines = 2;
            for (int ln = 0; ln < n_lines; ln++) {
                int s0 = start, s1 = start + span;
                if (n_lines == 2) { int mid = start + span / 2; if (ln == 0) s1 = mid + 1; else { s0 = mid - 1; ident = 0.88 + 0.12 * rng.u(); } }
                std::vector<int32_t> path;
                path.push_back(anc[s0]);
                for (int x = s0; x < s1; x++) {
                    int K = kb[b0 + x];
                    int a = b->hap_allele[h0 + (size_t)x * p + hap];
                    double e = rng.u();
                    if (e < P.miss) { /* skip the inner node: the cell is missing */ }
                    else {
                        if (e < P.miss + P.err) a = (a + 1 + (int)rng.below((uint32_t)(K - 1))) % K;
                        path.push_back(inner_first[b0 + x] + a);
                    }
                    path.push_back(anc[x + 1]);
                }
                if (rev) std::reverse(path.begin(), path.end());
                b->enode.insert(b->enode.end(), path.begin(), path.end());
                b->enode_off.push_back((int64_t)b->enode.size());
                b->entry_read.push_back(n_reads_chain);
                char buf[32]; snprintf(buf, sizeof buf, "%.6f", ident);
                b->entry_identity_txt.push_back(buf);
                b->entry_identity.push_back(strtof(buf, nullptr));
                b->entry_reverse.push_back(rev ? 1 : 0);
                b->read_hap.push_back(hap);
            }
            n_reads_chain++;
        }
        b->read_off.push_back(b->read_off.back() + n_reads_chain);
        b->entry_off.push_back((int64_t)b->entry_read.size());
    }
    b->n_nodes = next_node - 1;
    ahs_batch_in &v = b->view;
    v.n_chains = C; v.ploidy = p; v.chain_id = b->chain_id.data();
    v.bubble_off = b->bubble_off.data(); v.allele_off = b->allele_off.data(); v.anode_off = b->anode_off.data();
    v.anode = b->anode.data(); v.stage_a_order = b->stage_a_order.data();
    v.read_off = b->read_off.data(); v.entry_off = b->entry_off.data(); v.enode_off = b->enode_off.data();
    v.enode = b->enode.data(); v.entry_read = b->entry_read.data(); v.entry_identity = b->entry_identity.data();
    *out = b;
    return AHS_OK;
}

extern "C" const ahs_batch_in *ahs_synth_batch_in(const ahs_synth_batch *b) { return &b->view; }

extern "C" int64_t ahs_synth_counts(const ahs_synth_batch *b, int what) {
    switch (what) {
        case 0: return (int64_t)b->chain_id.size();
        case 1: return b->bubble_off.back();
        case 2: return (int64_t)b->anode_off.size() - 1;
        case 3: return (int64_t)b->anode.size();
        case 4: return b->read_off.back();
        case 5: return (int64_t)b->entry_read.size();
        case 6: return (int64_t)b->enode.size();
        case 7: return b->n_nodes;
    }
    return -1;
}
extern "C" const int32_t *ahs_synth_truth_read_hap(const ahs_synth_batch *b) { return b->read_hap.data(); }
extern "C" const uint8_t *ahs_synth_truth_hap_allele(const ahs_synth_batch *b) { return b->hap_allele.data(); }

extern "C" int ahs_synth_write_gfa_gaf(const ahs_synth_batch *b, const char *gfa_path, const char *gaf_path) {
    FILE *g = fopen(gfa_path, "w");
    if (!g) return AHS_ERR_ARG;
    const int C = (int)b->chain_id.size();
    auto name = [](int32_t id, char *buf) { snprintf(buf, 32, "utg%07dl", id); };
    char n1[32], n2[32];
    for (int32_t id = 1; id <= b->n_nodes; id++) { name(id, n1); fprintf(g, "S\t%s\tA\n", n1); }
    auto link = [&](int32_t from, int32_t to) {   // from+ -> to+ and its reverse complement to- -> from-
        name(from, n1); name(to, n2);
        fprintf(g, "L\t%s\t+\t%s\t+\t0M\n", n1, n2);
        fprintf(g, "L\t%s\t-\t%s\t-\t0M\n", n2, n1);
    };
    for (int c = 0; c < C; c++) {
        const int32_t *anc = b->anchors.data() + b->anchor_off[c];
        const int B = (int)(b->bubble_off[c + 1] - b->bubble_off[c]);
        link(b->chain_flank_l[c], anc[0]);
        for (int x = 0; x < B; x++) {
            int64_t gb = b->bubble_off[c] + x;
            int K = (int)(b->allele_off[gb + 1] - b->allele_off[gb]);
            for (int a = 0; a < K; a++) {
                int32_t in = b->anode[b->anode_off[b->allele_off[gb] + a] + 1];
                link(anc[x], in); link(in, anc[x + 1]);
            }
        }
        link(anc[B], b->chain_flank_r[c]);
    }
    fclose(g);
    FILE *f = fopen(gaf_path, "w");
    if (!f) return AHS_ERR_ARG;
    for (int c = 0; c < C; c++) {
        for (int64_t e = b->entry_off[c]; e < b->entry_off[c + 1]; e++) {
            std::string path;
            const bool rev = b->entry_reverse[e] != 0;
            for (int64_t x = b->enode_off[e]; x < b->enode_off[e + 1]; x++) { name(b->enode[x], n1); path += rev ? '<' : '>'; path += n1; }
            int64_t nn = b->enode_off[e + 1] - b->enode_off[e];
            long plen = (long)nn * 1000;
            fprintf(f, "c%dr%d\t%ld\t0\t%ld\t+\t%s\t%ld\t0\t%ld\t%ld\t%ld\t60\tNM:i:0\tAS:f:0\tdv:f:0\tid:f:%s\n",
                    b->chain_id[c], b->entry_read[e], plen, plen, path.c_str(), plen, plen, plen, plen, b->entry_identity_txt[e].c_str());
        }
    }
    fclose(f);
    return AHS_OK;
}

extern "C" void ahs_synth_free(ahs_synth_batch *b) { delete b; }
