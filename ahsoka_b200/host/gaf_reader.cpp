// gaf_reader.cpp — see gaf_reader.hpp.  Behaviour follows reference src/alignmentreader.cpp:69-189
// line by line; the comments name the statement each rule comes from.
#include "gaf_reader.hpp"

#include <cerrno>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string_view>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace ahs_host {
namespace {

// whitespace of `stringstream >> std::string` in the C locale
inline bool is_ws(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline bool is_dig(unsigned char c) { return c >= '0' && c <= '9'; }

struct Tok { const char* p; size_t n; };

// node id → chain id as `graph.nodes[id].chain_id` reads it (alignmentreader.cpp:180): a node that
// is in no chain, or not in the graph at all, answers 0 (Node::Node sets chain_id(0), graph.cpp:28-49)
struct ChainLookup {
    std::vector<int32_t> dense;
    const std::unordered_map<int, Node>* sparse = nullptr;
    explicit ChainLookup(const Graph& g) {
        int64_t mx = -1;
        bool neg = false;
        for (auto& kv : g.nodes) { if (kv.first < 0) neg = true; if (kv.first > mx) mx = kv.first; }
        if (!neg && mx < (int64_t)4 * (int64_t)g.nodes.size() + (1 << 20)) {
            dense.assign((size_t)(mx + 1), 0);
            for (auto& kv : g.nodes) dense[(size_t)kv.first] = kv.second.chain_id;
        } else sparse = &g.nodes;
    }
    int operator()(int id) const {
        if (!sparse) return (size_t)id < dense.size() ? dense[(size_t)id] : 0;
        auto it = sparse->find(id);
        return it == sparse->end() ? 0 : it->second.chain_id;
    }
};

struct Slab {
    std::vector<Tok> name, path;
    std::vector<float> identity;
    std::vector<int32_t> start, end;
    std::vector<int64_t> node_off{0}, chain_off{0};
    std::vector<int32_t> node_raw, chains;
    std::string ident_text;
    const char* err_at = nullptr;
    std::string err;
};

// std::stoi (alignmentreader.cpp:171-172): strtol base 10, no conversion or out of int range throws
bool parse_stoi(Tok t, int32_t& out) {
    char buf[64];
    std::string big;
    const char* s;
    if (t.n < sizeof buf) { memcpy(buf, t.p, t.n); buf[t.n] = 0; s = buf; } else { big.assign(t.p, t.n); s = big.c_str(); }
    char* e;
    errno = 0;
    long v = strtol(s, &e, 10);
    if (e == s || errno == ERANGE || v < INT_MIN || v > INT_MAX) return false;
    out = (int32_t)v;
    return true;
}

// std::stof (alignmentreader.cpp:135): strtof, no conversion or ERANGE throws
bool parse_stof(const char* p, size_t n, float& out) {
    char buf[64];
    std::string big;
    const char* s;
    if (n < sizeof buf) { memcpy(buf, p, n); buf[n] = 0; s = buf; } else { big.assign(p, n); s = big.c_str(); }
    char* e;
    errno = 0;
    float v = strtof(s, &e);
    if (e == s || errno == ERANGE) return false;
    out = v;
    return true;
}

void parse_slab(const char* b, const char* e, const ChainLookup& chain_of, Slab& s) {
    std::vector<int32_t> line_chains;
    char num[48];
    const char* ls = b;
    while (ls < e) {
        const char* le = (const char*)memchr(ls, '\n', (size_t)(e - ls));   // slabs end on a '\n'
        if (le == ls) { ls = le + 1; continue; }                            // `if (line.size() == 0) continue;` (:84)
        auto fail = [&](const char* what) { s.err_at = ls; s.err = what; };
        // sixteen `sstr >>` extractions (:88-118); a missing one leaves its string empty
        Tok tok[16];
        int nt = 0;
        for (const char* p = ls; nt < 16;) {
            while (p < le && is_ws((unsigned char)*p)) p++;
            if (p == le) break;
            const char* q = p;
            while (q < le && !is_ws((unsigned char)*q)) q++;
            tok[nt++] = Tok{p, (size_t)(q - p)};
            p = q;
        }
        if (nt < 16) { fail("fewer than 16 fields: the id:f: field is missing (assert, alignmentreader.cpp:129)"); return; }
        const Tok name = tok[0], path = tok[5], length = tok[10], idt = tok[15];
        // split at ':' (:120-128); the first part must be "id" (:129); the value is what follows the last ':' (:134)
        const char* colon = (const char*)memchr(idt.p, ':', idt.n);
        if (!((colon && colon - idt.p == 2) || (!colon && idt.n == 2)) || idt.p[0] != 'i' || idt.p[1] != 'd') {
            fail("field 16 does not start with id: (assert, alignmentreader.cpp:129)"); return;
        }
        const char* val = idt.p;
        for (const char* p = idt.p; p < idt.p + idt.n; p++) if (*p == ':') val = p + 1;
        float id_val;
        if (!parse_stof(val, (size_t)(idt.p + idt.n - val), id_val)) { fail("identity is not a float (stof, alignmentreader.cpp:135)"); return; }
        // nodes and directions (:138-150)
        const size_t ident_mark = s.ident_text.size();
        s.ident_text.append(name.p, name.n);
        s.ident_text.push_back('\t');
        s.ident_text.append(num, (size_t)snprintf(num, sizeof num, "%g", (double)id_val));   // `myfile << id_val` (:151)
        s.ident_text.push_back('\t');
        const size_t node_mark = s.node_raw.size();
        line_chains.clear();
        const char* pe = path.p + path.n;
        for (const char* p = path.p; p < pe;) {
            while (p < pe && (*p == '<' || *p == '>')) p++;
            if (p == pe) break;
            if (p == path.p) {   // `path.substr(beg-1,1)` with beg == 0 throws std::out_of_range (:147)
                fail("path does not start with '<' or '>' (out_of_range, alignmentreader.cpp:147)"); break;
            }
            const char* q = p + 1;
            while (q < pe && *q != '<' && *q != '>') q++;
            // raw_node_id (:48-54): the digits of the name, then stoi
            int64_t v = 0;
            bool any = false, over = false;
            for (const char* c = p; c < q; c++) if (is_dig((unsigned char)*c)) { any = true; v = v * 10 + (*c - '0'); if (v > INT_MAX) { over = true; v = INT_MAX; } }
            if (!any || over) { fail("node name without a usable integer id (stoi, alignmentreader.cpp:53)"); break; }
            s.node_raw.push_back((int32_t)v);
            s.ident_text.append(p, (size_t)(q - p));
            s.ident_text.push_back(',');
            const int ch = chain_of((int)v);
            bool seen = false;
            for (int32_t c : line_chains) if (c == ch) { seen = true; break; }
            if (!seen) line_chains.push_back(ch);
            p = q;
        }
        if (s.err_at) { s.node_raw.resize(node_mark); s.ident_text.resize(ident_mark); return; }
        s.ident_text.push_back('\t');
        s.ident_text.append(length.p, length.n);
        s.ident_text.push_back('\n');
        int32_t sp, ep;
        if (!parse_stoi(tok[7], sp) || !parse_stoi(tok[8], ep)) {   // the identities line is already written when stoi throws
            s.node_raw.resize(node_mark);
            fail("start/end position is not an int (stoi, alignmentreader.cpp:171-172)"); return;
        }
        s.name.push_back(name); s.path.push_back(path);
        s.identity.push_back(id_val); s.start.push_back(sp); s.end.push_back(ep);
        s.node_off.push_back((int64_t)s.node_raw.size());
        s.chains.insert(s.chains.end(), line_chains.begin(), line_chains.end());
        s.chain_off.push_back((int64_t)s.chains.size());
        ls = le + 1;
    }
}

// `a.nodes == b.nodes` on the node NAMES of two path fields
bool same_node_names(Tok a, Tok b) {
    const char *p = a.p, *pe = a.p + a.n, *r = b.p, *re = b.p + b.n;
    for (;;) {
        while (p < pe && (*p == '<' || *p == '>')) p++;
        while (r < re && (*r == '<' || *r == '>')) r++;
        if (p == pe || r == re) return p == pe && r == re;
        const char* q = p + 1; while (q < pe && *q != '<' && *q != '>') q++;
        const char* t = r + 1; while (t < re && *t != '<' && *t != '>') t++;
        if (q - p != t - r || memcmp(p, r, (size_t)(q - p)) != 0) return false;
        p = q; r = t;
    }
}

struct Mapped {
    const char* p = nullptr; size_t n = 0; bool mapped = false; std::string owned;
    ~Mapped() { if (mapped) munmap((void*)p, n); }
};

}  // namespace

int read_gaf(const std::string& filename, const Graph& graph, GafStore& st, std::string& err, int threads) {
    // `filename.substr(0, filename.find(".gaf")) + "-alignment_identities.txt"` (:74-75)
    const std::string ident_name = filename.substr(0, filename.find(".gaf")) + "-alignment_identities.txt";
    FILE* ident = fopen(ident_name.c_str(), "wb");
    Mapped m;
    int fd = open(filename.c_str(), O_RDONLY);
    if (fd >= 0) {
        struct stat sb;
        if (fstat(fd, &sb) == 0 && sb.st_size > 0) {
            void* a = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (a != MAP_FAILED) { m.p = (const char*)a; m.n = (size_t)sb.st_size; m.mapped = true; madvise(a, m.n, MADV_SEQUENTIAL); }
            else {
                m.owned.resize((size_t)sb.st_size);
                size_t got = 0; ssize_t r;
                while (got < m.owned.size() && (r = read(fd, &m.owned[got], m.owned.size() - got)) > 0) got += (size_t)r;
                m.owned.resize(got); m.p = m.owned.data(); m.n = got;
            }
        }
        close(fd);
    }   // an unreadable file is an empty one for the reference too (ifstream fails, the loop never runs)
    // `if (!file.good()) break;` (:82): a last line without '\n' sets eofbit and is dropped
    size_t valid = m.n;
    while (valid > 0 && m.p[valid - 1] != '\n') valid--;
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (valid < ((size_t)1 << 20)) T = 1;
    std::vector<size_t> cut(T + 1, valid);
    cut[0] = 0;
    for (int t = 1; t < T; t++) {
        size_t c = valid / T * t;
        if (c < cut[t - 1]) c = cut[t - 1];
        while (c < valid && m.p[c] != '\n') c++;
        cut[t] = c < valid ? c + 1 : valid;
    }
    const bool trace = getenv("AHSOKA_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto lap = [&](const char* what, std::chrono::steady_clock::time_point& t0) {
        if (trace) fprintf(stderr, "timing:   gaf %s %.1f\n", what, std::chrono::duration<double, std::milli>(now() - t0).count());
        t0 = now();
    };
    auto t0 = now();
    ChainLookup chain_of(graph);
    lap("chain_lookup", t0);
    std::vector<Slab> slab(T);
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < T; t++) pool.emplace_back([&, t] { parse_slab(m.p + cut[t], m.p + cut[t + 1], chain_of, slab[t]); });
        parse_slab(m.p + cut[0], m.p + cut[1], chain_of, slab[0]);
        for (auto& th : pool) th.join();
    }
    lap("parse", t0);
    int rc = 0;
    for (int t = 0; t < T; t++) {
        if (ident) fwrite(slab[t].ident_text.data(), 1, slab[t].ident_text.size(), ident);
        if (slab[t].err_at) {   // the reference dies here: everything before is on disk, nothing after
            int64_t line = 1;
            for (const char* p = m.p; p < slab[t].err_at; p++) line += *p == '\n';
            err = filename + ":" + std::to_string(line) + ": " + slab[t].err;
            rc = 65;
            break;
        }
    }
    if (ident) fclose(ident);
    lap("identities_file", t0);
    if (rc) return rc;

    // merge in file order: concatenate, intern names, list each line once per chain it touches.  Every phase runs on
    // the T threads again: the slabs are copied in place; names are interned by the thread that owns their hash class
    // (ids are unique, first-appearance order WITHIN an owner, which is all the per-chain read numbering needs); chains are
    // filled by the thread that owns the chain id, each walking the lines in file order.
    std::vector<size_t> line_base(T + 1, 0), node_base(T + 1, 0);
    for (int t = 0; t < T; t++) { line_base[t + 1] = line_base[t] + slab[t].name.size(); node_base[t + 1] = node_base[t] + slab[t].node_raw.size(); }
    const size_t n_lines = line_base[T], n_nodes = node_base[T];
    if (n_lines > (size_t)INT32_MAX) { err = filename + ": more than 2^31 alignment lines"; return 65; }
    st.name_id.resize(n_lines); st.identity.resize(n_lines); st.startpos.resize(n_lines); st.endpos.resize(n_lines);
    st.node_off.resize(n_lines + 1); st.node_raw.resize(n_nodes);
    st.node_off[0] = 0;
    std::vector<Tok> path(n_lines), name_tok(n_lines);
    std::vector<size_t> name_hash(n_lines);
    std::vector<int32_t> local_id(n_lines);
    auto run = [&](auto&& fn) {
        std::vector<std::thread> pool;
        for (int t = 1; t < T; t++) pool.emplace_back([&, t] { fn(t); });
        fn(0);
        for (auto& th : pool) th.join();
    };
    run([&](int t) {                                          // slab t -> its place in the store
        Slab& sl = slab[t];
        const size_t lb = line_base[t], nb = node_base[t];
        if (!sl.node_raw.empty()) memcpy(&st.node_raw[nb], sl.node_raw.data(), sl.node_raw.size() * sizeof(int32_t));
        for (size_t i = 0; i < sl.name.size(); i++) {
            st.identity[lb + i] = sl.identity[i]; st.startpos[lb + i] = sl.start[i]; st.endpos[lb + i] = sl.end[i];
            st.node_off[lb + i + 1] = (int64_t)nb + sl.node_off[i + 1];
            path[lb + i] = sl.path[i]; name_tok[lb + i] = sl.name[i];
            name_hash[lb + i] = std::hash<std::string_view>()(std::string_view(sl.name[i].p, sl.name[i].n));
        }
        std::vector<int32_t>().swap(sl.node_raw);
    });
    std::vector<std::vector<Tok>> own_names(T);
    run([&](int t) {                                          // names whose hash class is t
        std::unordered_map<std::string_view, int32_t> intern;
        intern.reserve(n_lines / T + 16);
        for (size_t line = 0; line < n_lines; line++) {
            if ((int)(name_hash[line] % (size_t)T) != t) continue;
            auto ins = intern.emplace(std::string_view(name_tok[line].p, name_tok[line].n), (int32_t)own_names[t].size());
            if (ins.second) own_names[t].push_back(name_tok[line]);
            local_id[line] = ins.first->second;
        }
    });
    std::vector<size_t> gbase(T + 1, 0);
    for (int t = 0; t < T; t++) gbase[t + 1] = gbase[t] + own_names[t].size();
    st.names.resize(gbase[T]);
    run([&](int t) {
        for (size_t i = 0; i < own_names[t].size(); i++) st.names[gbase[t] + i].assign(own_names[t][i].p, own_names[t][i].n);
        for (size_t line = line_base[t]; line < line_base[t + 1]; line++)
            st.name_id[line] = (int32_t)(gbase[name_hash[line] % (size_t)T] + (size_t)local_id[line]);
    });
    for (auto& sl : slab) for (int32_t ch : sl.chains) st.by_chain[ch];       // create the lists; the threads below only look them up
    run([&](int t) {                                          // chains whose id class is t
        for (int u = 0; u < T; u++) {
            const Slab& sl = slab[u];
            for (size_t i = 0; i + 1 < sl.chain_off.size(); i++) {
                const int32_t line = (int32_t)(line_base[u] + i);
                for (int64_t k = sl.chain_off[i]; k < sl.chain_off[i + 1]; k++) {
                    const int32_t ch = sl.chains[k];
                    if ((int)((uint32_t)ch % (uint32_t)T) != t) continue;
                    std::vector<int32_t>& v = st.by_chain.find(ch)->second;
                    if (!v.empty()) {
                        const int32_t prev = v.back();
                        if (st.name_id[prev] == st.name_id[line] && st.identity[prev] == st.identity[line] &&
                            st.startpos[prev] == st.startpos[line] && st.endpos[prev] == st.endpos[line] &&
                            same_node_names(path[prev], path[line])) continue;
                    }
                    v.push_back(line);
                }
            }
        }
    });
    lap("merge", t0);
    return 0;
}

}  // namespace ahs_host
