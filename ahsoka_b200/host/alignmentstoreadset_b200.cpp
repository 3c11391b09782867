// alignmentstoreadset_b200.cpp — drop-in body for the reference's
//     void alignmentsToReadset(AlignmentReader&, Graph&, unordered_map<int, unordered_map<int,
//          vector<vector<int>>>>& pathToAlleles, string readsetfile, bool shell_logging,
//          vector<pair<int,int>>& size_sorting, std::mutex&)
// (reference src/alignmentstoreadset.cpp:55, called at src/polyassembly.cpp:171).
//
// It is compiled against the reference's own headers (graph.hpp, alignmentreader.hpp), does
// three things and nothing else:
//   1. flatten  alignmentreader.alignments / pathToAlleles / size_sorting  into the CSR batch
//      of include/ahsoka_b200.h  (SURVEY §8b);
//   2. call ahs_phase_batch() — the sm_100a CUDA path; there is no CPU path behind it;
//   3. write <prefix>-result.txt, <prefix>-chain<id>-result.txt and the stdout "hap:" lines with
//      the semantics of src/alignmentstoreadset.cpp:70-83 and :411-486.
//      and <prefix>-chain<id>-readset_final.txt (:298-303; the text of ReadSet::toString() is third-party and
//      unverifiable here: the format is the one of oracle/whatshap_shim, to which the tests pin it).
// Not reproduced: ./logfile.log and <prefix>-chain<id>-readset.txt (:284-293: dumps of the stage-A read sets, which the
// device reduces to two scalars per chain, SURVEY A#23), see INTEGRATION.md.
// A chain that exceeds a build limit (status >= AHS_CHAIN_TOO_LARGE) is reported on stderr, gets header-only output like
// an empty chain, and makes the process exit with code 3 (ahs_host::unphased_chains()).
#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "alignmentreader.hpp"   // reference header, found through -I<reference>/src
#include "graph.hpp"             // reference header
#include "ahsoka_b200.h"
#include "gaf_reader.hpp"

namespace ahs_host {

typedef std::unordered_map<int, std::unordered_map<int, std::vector<std::vector<int>>>> ChainAlleles;

// AHSOKA_TIMING=1: one "timing: <stage> <ms>" line per host stage on stderr
struct StageTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    const char* name;
    explicit StageTimer(const char* n) : name(n) {}
    ~StageTimer() {
        if (!getenv("AHSOKA_TIMING")) return;
        std::cerr << "timing: " << name << " " << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() << std::endl;
    }
};

struct FlatBatch {
    std::vector<int32_t> chain_id, anode, stage_a_order, enode, entry_read;
    std::vector<int64_t> bubble_off{0}, allele_off{0}, anode_off{0}, read_off{0}, entry_off{0}, enode_off{0};
    std::vector<float> entry_identity;
    std::vector<std::vector<std::string>> read_names;   // per chain: chain-local read index -> name
    ahs_batch_in view;
};

static bool same_entry(AlignmentPath& a, AlignmentPath& b) {
    return a.name == b.name && a.id == b.id && a.startpos == b.startpos && a.endpos == b.endpos && a.nodes == b.nodes;
}

// Alleles of one chain (reference src/polyassembly.cpp:126-140 containers), appended CSR; returns the
// number of bubbles B the phasing sees (0 for chains of ≤1 bubble, alignmentstoreadset.cpp:86).
static int flatten_chain_alleles(ChainAlleles& pathToAlleles, int chainid, FlatBatch& fb) {
    // the reference iterates a COPY of this map (alignmentstoreadset.cpp:76, :90) and that order is stage A's; a libstdc++
    // unordered_map copy keeps the element order of its source (_M_assign walks the source list), so the source is read in place
    const auto& bubbles = pathToAlleles[chainid];
    fb.chain_id.push_back(chainid);
    int B = 0;
    if (bubbles.size() > 1) for (auto& kv : bubbles) B = std::max(B, kv.first + 1);
    std::vector<char> seen(B, 0);
    for (auto& kv : bubbles) if (kv.first >= 0 && kv.first < B) { fb.stage_a_order.push_back(kv.first); seen[kv.first] = 1; }
    for (int b = 0; b < B; b++) if (!seen[b]) fb.stage_a_order.push_back(b);
    for (int b = 0; b < B; b++) {
        auto it = bubbles.find(b);
        if (it != bubbles.end()) for (auto& path : it->second) {
            fb.anode.insert(fb.anode.end(), path.begin(), path.end());
            fb.anode_off.push_back((int64_t)fb.anode.size());
        }
        fb.allele_off.push_back((int64_t)fb.anode_off.size() - 1);
    }
    fb.bubble_off.push_back((int64_t)fb.allele_off.size() - 1);
    return B;
}

static void finish_view(FlatBatch& fb, int ploidy) {
    ahs_batch_in& v = fb.view;
    v.n_chains = (int32_t)fb.chain_id.size(); v.ploidy = ploidy; v.chain_id = fb.chain_id.data();
    v.bubble_off = fb.bubble_off.data(); v.allele_off = fb.allele_off.data(); v.anode_off = fb.anode_off.data();
    v.anode = fb.anode.data(); v.stage_a_order = fb.stage_a_order.data(); v.read_off = fb.read_off.data();
    v.entry_off = fb.entry_off.data(); v.enode_off = fb.enode_off.data(); v.enode = fb.enode.data();
    v.entry_read = fb.entry_read.data(); v.entry_identity = fb.entry_identity.data();
}

// the containers of reference src/alignmentreader.hpp:38 and
// src/polyassembly.cpp:126-140, flattened CSR-by-chain in size_sorting order.
void flatten(AlignmentReader& reader, ChainAlleles& pathToAlleles,
             std::vector<std::pair<int, int>>& size_sorting, int ploidy, FlatBatch& fb) {
    for (auto& size : size_sorting) {
        const int chainid = size.second;
        const int B = flatten_chain_alleles(pathToAlleles, chainid, fb);
        fb.read_names.emplace_back();
        std::vector<std::string>& names = fb.read_names.back();
        if (B > 1) {
            std::unordered_map<std::string, int> intern;
            auto it = reader.alignments.find(chainid);
            if (it != reader.alignments.end()) {
                std::vector<AlignmentPath>& v = it->second;
                for (size_t e = 0; e < v.size(); e++) {
                    if (e > 0 && same_entry(v[e], v[e - 1])) continue;       // per-node duplicates, alignmentreader.cpp:176-183
                    auto ins = intern.emplace(v[e].name, (int)names.size());
                    if (ins.second) names.push_back(v[e].name);
                    std::vector<int> raw = v[e].getRawIds();
                    fb.enode.insert(fb.enode.end(), raw.begin(), raw.end());
                    fb.enode_off.push_back((int64_t)fb.enode.size());
                    fb.entry_read.push_back(ins.first->second);
                    fb.entry_identity.push_back(v[e].id);
                }
            }
        }
        fb.read_off.push_back(fb.read_off.back() + (int64_t)names.size());
        fb.entry_off.push_back((int64_t)fb.entry_read.size());
    }
    finish_view(fb, ploidy);
}

// the same batch from the native reader's store (gaf_reader.hpp) — no strings are
// parsed or copied here, the per-chain read index is the first-appearance order of the name ids.
void flatten_store(const GafStore& st, ChainAlleles& pathToAlleles,
                   std::vector<std::pair<int, int>>& size_sorting, int ploidy, FlatBatch& fb) {
    std::vector<int32_t> stamp(st.names.size(), -1), local(st.names.size(), 0);
    {
        size_t n_entries = 0, n_nodes = 0;
        for (auto& size : size_sorting) {
            auto it = st.by_chain.find(size.second);
            if (it == st.by_chain.end()) continue;
            n_entries += it->second.size();
            for (int32_t line : it->second) n_nodes += (size_t)(st.node_off[line + 1] - st.node_off[line]);
        }
        fb.enode.reserve(n_nodes); fb.enode_off.reserve(n_entries + 1); fb.entry_read.reserve(n_entries); fb.entry_identity.reserve(n_entries);
    }
    int32_t c = 0;
    for (auto& size : size_sorting) {
        const int chainid = size.second;
        const int B = flatten_chain_alleles(pathToAlleles, chainid, fb);
        fb.read_names.emplace_back();
        std::vector<std::string>& names = fb.read_names.back();
        if (B > 1) {
            auto it = st.by_chain.find(chainid);
            if (it != st.by_chain.end()) {
                for (int32_t line : it->second) {
                    const int32_t g = st.name_id[line];
                    if (stamp[g] != c) { stamp[g] = c; local[g] = (int32_t)names.size(); names.push_back(st.names[g]); }
                    fb.enode.insert(fb.enode.end(), st.node_raw.begin() + st.node_off[line], st.node_raw.begin() + st.node_off[line + 1]);
                    fb.enode_off.push_back((int64_t)fb.enode.size());
                    fb.entry_read.push_back(local[g]);
                    fb.entry_identity.push_back(st.identity[line]);
                }
            }
        }
        fb.read_off.push_back(fb.read_off.back() + (int64_t)names.size());
        fb.entry_off.push_back((int64_t)fb.entry_read.size());
        c++;
    }
    finish_view(fb, ploidy);
}

// Emission, semantics of reference src/alignmentstoreadset.cpp:70-83 and :411-486: the same bytes in the same files;
// a haplotype line is formatted once into a buffer and written to both files (the reference streams every node twice
// and flushes with endl after every line).
static inline void put_int(std::string& s, long v) {
    char buf[24];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    s.append(buf, (size_t)(r.ptr - buf));
}

static std::atomic<int> g_unphased{0};
int unphased_chains() { return g_unphased.load(); }

void emit(const ahs_batch_out& out, Graph& graph,
          std::unordered_map<int, std::unordered_map<int, std::vector<std::vector<int>>>>& pathToAlleles,
          std::vector<std::pair<int, int>>& size_sorting, const std::string& prefix,
          const std::vector<std::vector<std::string>>& read_names) {
    const bool dumps = getenv("AHSOKA_NO_READSET_DUMPS") == nullptr;
    const int ploidy = out.ploidy;
    const size_t C = size_sorting.size();
    // chains are independent: worker threads format a chain's text and write its <prefix>-chain<id>-result.txt; the
    // shared -result.txt and the stdout lines are then written in size_sorting order by this thread
    std::vector<std::string> full_of(C), haps_of(C);
    for (size_t c = 0; c < C; c++) pathToAlleles[size_sorting[c].second];      // the reference's operator[] (:76), before the threads start
    std::atomic<size_t> next{0};
    auto work = [&]() {
        std::string line;
        std::unordered_set<int> usednodes;
        for (;;) {
            const size_t c0 = next.fetch_add(64);
            if (c0 >= C) break;
            for (size_t c = c0; c < std::min(C, c0 + 64); c++) {
                const int chainid = size_sorting[c].second;
                const auto& alleles_of = pathToAlleles.find(chainid)->second;
                std::string& full = full_of[c];
                full += "chain id: "; put_int(full, chainid); full += '\n';
                full += "size of chain: "; put_int(full, (long)alleles_of.size()); full += '\n';
                if (out.status[c] != AHS_CHAIN_OK) continue;
                std::ofstream resfile(prefix + "-chain" + std::to_string(chainid) + "-result.txt");
                const int64_t p0 = out.pos_off[c], n_pos = out.pos_off[c + 1] - p0;
                for (int i = 0; i < ploidy; i++) {
                    usednodes.clear();
                    full += "haplotype "; put_int(full, i); full += ":\n";
                    line.clear();
                    for (int64_t j = 0; j < n_pos; j++) {
                        const uint32_t cons = out.hap_allele[(p0 + j) * ploidy + i];
                        const std::vector<int>& ap = alleles_of.at(out.pos[p0 + j]).at(cons);
                        for (size_t ind = 0; ind + 1 < ap.size(); ind++) {
                            const int single = ap[ind], nxt = ap[ind + 1];
                            if (usednodes.count(single)) continue;
                            // Graph::getNode + Graph::getEdge (graph.cpp:502-512, 251-261) without the scan over all
                            // nodes and without copying node sequences (f2): first of (single,+), (single,-) with an edge to nxt
                            bool end = false;
                            for (bool val : {true, false}) {
                                auto eit = graph.edges.find(DirectedNode(single, val));
                                if (eit == graph.edges.end()) continue;
                                bool hit = false;
                                for (auto& to : eit->second) if (to.id == nxt) { hit = true; break; }
                                if (hit) { end = val; break; }
                            }
                            put_int(line, single);
                            line += end ? "(+)," : "(-),";
                            usednodes.insert(single);
                        }
                    }
                    line += '\n';
                    resfile.write(line.data(), (std::streamsize)line.size());
                    full += line;
                }
                resfile.close();
                if (dumps) {                                                                 // :298-303, the final (sorted) read set
                    std::string rs;
                    const int64_t r0 = out.read_off[c], r1 = out.read_off[c + 1];
                    rs += "readset size: "; put_int(rs, (long)(r1 - r0)); rs += "\nReadSet:\n";
                    for (int64_t r = r0; r < r1; r++) {
                        rs += "  "; rs += read_names[c][(size_t)out.read_id[r]]; rs += " (";
                        for (int64_t x = out.cell_off[r]; x < out.cell_off[r + 1]; x++) {
                            if (x > out.cell_off[r]) rs += ';';
                            rs += '['; put_int(rs, out.cell_pos[x]); rs += ','; put_int(rs, (long)out.cell_allele[x]); rs += ",30]";      // quality 30: :118, :235
                        }
                        rs += ")\n";
                    }
                    rs += '\n';
                    std::ofstream rsf(prefix + "-chain" + std::to_string(chainid) + "-readset_final.txt");
                    rsf.write(rs.data(), (std::streamsize)rs.size());
                }
                std::string& haps = haps_of[c];
                for (int i = 0; i < ploidy; i++) {                                           // :479-486
                    haps += "hap: \n";
                    for (int64_t j = 0; j < n_pos; j++) {
                        put_int(haps, (long)out.hap_allele[(p0 + j) * ploidy + i]); haps += '('; put_int(haps, out.pos[p0 + j]); haps += "),";
                    }
                    haps += '\n';
                }
            }
        }
    };
    {
        unsigned T = std::thread::hardware_concurrency();
        if (T < 1) T = 1;
        if (C < 256) T = 1;
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < T; t++) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
    }
    std::ofstream full_output(prefix + "-result.txt", std::ios_base::app);        // append, :72
    for (size_t c = 0; c < C; c++) {
        full_output.write(full_of[c].data(), (std::streamsize)full_of[c].size());
        if (out.status[c] >= AHS_CHAIN_TOO_LARGE) {
            g_unphased++;
            std::cerr << "ahsoka_b200: chain " << size_sorting[c].second << " exceeds a build limit and was NOT phased (status " << out.status[c]
                      << "): header-only output; the process will exit with code 3" << std::endl;
        }
        std::cout.write(haps_of[c].data(), (std::streamsize)haps_of[c].size());
    }
    std::cout.flush();
    full_output.close();
}

}  // namespace ahs_host

namespace ahs_host {

static void dump_batch(const FlatBatch& fb, int ploidy, const char* dump) {
    // raw little-endian dump of the flattened batch, for the Python tests
    FILE* f = fopen(dump, "wb");
    if (!f) return;
    auto w64 = [&](int64_t v) { fwrite(&v, 8, 1, f); };
    auto wv = [&](const void* p, size_t bytes) { fwrite(p, 1, bytes, f); };
    w64(fb.view.n_chains); w64(ploidy); w64(fb.bubble_off.back()); w64((int64_t)fb.anode_off.size() - 1); w64((int64_t)fb.anode.size());
    w64((int64_t)fb.entry_read.size()); w64((int64_t)fb.enode.size());
    wv(fb.chain_id.data(), 4 * fb.chain_id.size()); wv(fb.bubble_off.data(), 8 * fb.bubble_off.size());
    wv(fb.allele_off.data(), 8 * fb.allele_off.size()); wv(fb.anode_off.data(), 8 * fb.anode_off.size());
    wv(fb.anode.data(), 4 * fb.anode.size()); wv(fb.stage_a_order.data(), 4 * fb.stage_a_order.size());
    wv(fb.read_off.data(), 8 * fb.read_off.size()); wv(fb.entry_off.data(), 8 * fb.entry_off.size());
    wv(fb.enode_off.data(), 8 * fb.enode_off.size()); wv(fb.enode.data(), 4 * fb.enode.size());
    wv(fb.entry_read.data(), 4 * fb.entry_read.size()); wv(fb.entry_identity.data(), 4 * fb.entry_identity.size());
    fclose(f);
}

static int ploidy_from_env() {
    const char* pl = getenv("AHSOKA_PLOIDY");                 // reference: hard-coded 2 (:306)
    return pl ? atoi(pl) : 2;
}

static void phase_and_emit(FlatBatch& fb, int ploidy, Graph& graph, ChainAlleles& pathToAlleles, const std::string& prefix,
                           std::vector<std::pair<int, int>>& size_sorting) {
    if (const char* dump = getenv("AHSOKA_DUMP_BATCH")) dump_batch(fb, ploidy, dump);
    ahs_batch_out out;
    const char* dv = getenv("AHSOKA_DEVICE");
    std::vector<int> devices;                                 // AHSOKA_DEVICES=0,1,...: the chains are dealt over these GPUs
    if (const char* dl = getenv("AHSOKA_DEVICES")) for (const char* q = dl; *q;) { devices.push_back(atoi(q)); while (*q && *q != ',') q++; if (*q == ',') q++; }
    int rc;
    {
        StageTimer t("phase_batch");
        if (devices.size() > 1) rc = ahs_phase_batch_multi(&fb.view, &out, devices.data(), (int)devices.size());
        else rc = ahs_phase_batch(&fb.view, &out, !devices.empty() ? devices[0] : dv ? atoi(dv) : 0);
    }
    if (rc != AHS_OK) {
        std::cerr << "ahsoka_b200: phasing failed (" << rc << "): " << ahs_last_error() << std::endl;
        exit(70);                                             // fail loudly: no CPU path behind the ABI
    }
    {
        StageTimer t("emit");
        emit(out, graph, pathToAlleles, size_sorting, prefix, fb.read_names);
    }
    ahs_free_out(&out);
}

}  // namespace ahs_host

void alignmentsToReadset(AlignmentReader& alignmentreader, Graph& graph,
                         std::unordered_map<int, std::unordered_map<int, std::vector<std::vector<int>>>>& pathToAlleles,
                         std::string readsetfile, bool shell_logging, std::vector<std::pair<int, int>>& size_sorting,
                         std::mutex& g_display_mutex) {
    std::lock_guard<std::mutex> guard(g_display_mutex);
    (void)shell_logging;
    const int ploidy = ahs_host::ploidy_from_env();
    ahs_host::FlatBatch fb;
    {
        ahs_host::StageTimer t("flatten");
        ahs_host::flatten(alignmentreader, pathToAlleles, size_sorting, ploidy, fb);
    }
    ahs_host::phase_and_emit(fb, ploidy, graph, pathToAlleles, readsetfile, size_sorting);
}

// Same call for a host that reads its alignments with ahs_host::read_gaf.
void alignmentsToReadset(const ahs_host::GafStore& store, Graph& graph,
                         std::unordered_map<int, std::unordered_map<int, std::vector<std::vector<int>>>>& pathToAlleles,
                         std::string readsetfile, bool shell_logging, std::vector<std::pair<int, int>>& size_sorting,
                         std::mutex& g_display_mutex) {
    std::lock_guard<std::mutex> guard(g_display_mutex);
    (void)shell_logging;
    const int ploidy = ahs_host::ploidy_from_env();
    ahs_host::FlatBatch fb;
    {
        ahs_host::StageTimer t("flatten");
        ahs_host::flatten_store(store, pathToAlleles, size_sorting, ploidy, fb);
    }
    ahs_host::phase_and_emit(fb, ploidy, graph, pathToAlleles, readsetfile, size_sorting);
}
