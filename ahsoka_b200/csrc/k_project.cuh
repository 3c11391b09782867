// k_project.cuh — projection of alignment entries onto bubbles -> read x bubble allele matrix.
//
// Replaces reference src/alignmentstoreadset.cpp:90-136 (stage A), :146-209 (filter, boundary
// set), :210-254 (stage B), :262-297 (filter, ReadSet::sort()) and is_subset (:495-548).
//
// Layout in HBM (per chain c, R_c distinct read names, B_c bubbles):
//   mask[mrow_off[c] + r*B_c + b]  u16: bit a (a<15) = some entry of read r with identity*100 > 90
//                                  contains the INNER nodes of allele a of bubble b (stage B test);
//                                  bit 15 = some entry contains a FULL allele path of b (stage A).
//   Set with atomicOr, so the result does not depend on scheduling.  The few order-dependent
//   facts of the reference (which entry creates a read -> its mapq; first matching allele) are
//   recovered from 64-bit atomicMin keys (position | allele | entry index), i.e. the minimum in
//   the reference's own loop order (position asc, allele asc, entry asc).
//
// Finding matches: every allele path has one TRIGGER node (its first inner node; path[0] if it
// has no inner node).  A path can only be contained in an entry that contains its trigger, so a
// hash table trigger node -> alleles turns the reference's bubbles x alleles x entries scan into
// one lookup per entry node.  Paths without inner nodes match EVERY entry in stage B
// (std::includes of an empty range, SURVEY A#9): they are kept per bubble as `bubble_univ`.
#pragma once
#include "common.cuh"
#include "device_batch.cuh"

namespace ahs {

// ---------------------------------------------------------------- generic helpers
__global__ void k_owner(const int64_t* __restrict__ off, int n_owner, int64_t n_elem, int32_t* __restrict__ out) {
    for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < n_elem; x += (int64_t)gridDim.x * blockDim.x)
        out[x] = owner_of(off, n_owner, x);
}

// ---------------------------------------------------------------- input validation (first kernel of every pass)
// err_flags bits: 1 empty allele path, 2 more than 15 alleles, 4 offsets not monotone, 8 entry_read out of
// range, 16 stage_a_order is not a permutation.  Every later kernel of phase 1 returns at once if a bit is set,
// so a malformed batch can never drive an out-of-bounds access; the host reports the error after sync #1.
#define AHS_BAIL_ON_ERR(d) do { if (*(volatile const int32_t*)(d).err_flags) return; } while (0)

__global__ void k_validate(DB d, int32_t* __restrict__ max_k) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, x0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int bad = 0, mk = 0;
    for (int64_t b = x0; b < d.NB; b += stride) {
        const int64_t k = d.allele_off[b + 1] - d.allele_off[b];
        if (k < 0) bad |= 4; else if (k > MAX_ALLELES) bad |= 2; else mk = max(mk, (int)k);
    }
    for (int64_t a = x0; a < d.NA; a += stride) { const int64_t l = d.anode_off[a + 1] - d.anode_off[a]; if (l < 0) bad |= 4; else if (l == 0) bad |= 1; }
    for (int64_t e = x0; e < d.NE; e += stride) if (d.enode_off[e + 1] < d.enode_off[e]) bad |= 4;
    mk = warp_max_i32(mk);
    if (lane_id() == 0 && mk) atomicMax(max_k, mk);
    if (bad) atomicOr(d.err_flags, bad);
}

// entry_read in range (needs entry_chain), stage_a_order values in range (needs bubble_chain)
__global__ void k_validate_owned(DB d) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, x0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int bad = 0;
    for (int64_t e = x0; e < d.NE; e += stride) {
        const int c = d.entry_chain[e];
        const int32_t r = d.entry_read[e];
        if (r < 0 || r >= d.read_off[c + 1] - d.read_off[c]) bad |= 8;
    }
    if (d.stage_a_order) for (int64_t gb = x0; gb < d.NB; gb += stride) {
        const int c = d.bubble_chain[gb];
        const int32_t v = d.stage_a_order[gb];
        if (v < 0 || v >= d.bubble_off[c + 1] - d.bubble_off[c]) bad |= 16;
    }
    if (bad) atomicOr(d.err_flags, bad);
}

// every slot of rankA written <=> stage_a_order is a permutation of each chain's bubble ids
__global__ void k_validate_perm(DB d) {
    for (int64_t gb = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gb < d.NB; gb += (int64_t)gridDim.x * blockDim.x)
        if (d.rankA[gb] < 0) atomicOr(d.err_flags, 16);
}

// stage-A visit rank of each bubble (inverse of stage_a_order, alignmentstoreadset.cpp:90)
__global__ void k_rank_a(DB d) {
    AHS_BAIL_ON_ERR(d);
    for (int64_t gb = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gb < d.NB; gb += (int64_t)gridDim.x * blockDim.x) {
        int c = d.bubble_chain[gb];
        int64_t b0 = d.bubble_off[c];
        if (d.stage_a_order) d.rankA[b0 + d.stage_a_order[gb]] = (int32_t)(gb - b0);
        else d.rankA[gb] = (int32_t)(d.bubble_off[c + 1] - 1 - gb);
    }
}

// ---------------------------------------------------------------- K0: trigger table
// One open-addressing table PER CHAIN (region hoff[c], size hmaskc[c] + 1 = power of two >= 2 x alleles of the
// chain), so that all probes of a chain's entries fall into a few KB that stay in L1/L2.  Slot = (head << 32 | node
// id); all ones = empty; head = first allele of the list of alleles triggered by the node (inc_next links).
constexpr unsigned long long SLOT_EMPTY = ~0ull;

__global__ void k_build_triggers(DB d) {
    AHS_BAIL_ON_ERR(d);
    for (int64_t ga = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ga < d.NA; ga += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gb = d.allele_bubble[ga];
        const int c = d.bubble_chain[gb];
        const int64_t o = d.anode_off[ga];
        const int len = (int)(d.anode_off[ga + 1] - o);
        d.inc_next[ga] = -1;
        if (len <= 0) { atomicOr(d.err_flags, 1); continue; }
        if (d.bubble_off[c + 1] - d.bubble_off[c] <= 1) continue;          // trivial chain: never phased (:86)
        const int a = (int)(ga - d.allele_off[gb]);
        if (a >= MAX_ALLELES) { atomicOr(d.err_flags, 2); continue; }
        // everything k_project needs about a 3-node path in one 16-byte record: bubble (chain-local) | allele << 20 |
        // simple-path flag << 31, end nodes, stage-A rank
        d.arec[ga] = make_int4((int)((uint32_t)(gb - d.bubble_off[c]) | ((uint32_t)a << 20) | (len == 3 ? 0x80000000u : 0u)),
                               len == 3 ? d.anode[o] : 0, len == 3 ? d.anode[o + 2] : 0, d.rankA[gb]);
        if (len <= 2) { atomicMin(&d.bubble_univ[gb], (uint32_t)a); atomicOr(&d.ch_flags[c], CH_HAS_UNIV); }   // no inner node: matches every entry (A#9)
        const uint32_t trig = (uint32_t)(len >= 3 ? d.anode[o + 1] : d.anode[o]);
        unsigned long long* tab = d.hslots + d.hoff[c];
        const uint32_t mask = d.hmaskc[c];
        uint32_t slot = hash_slot(trig, mask);
        while (true) {
            const unsigned long long claimed = (0xfffffffeull << 32) | trig;          // head = -2: no allele linked yet
            const unsigned long long prev = atomicCAS(&tab[slot], SLOT_EMPTY, claimed);
            if (prev == SLOT_EMPTY || (uint32_t)prev == trig) break;
            slot = (slot + 1) & mask;
        }
        const int32_t prev_head = atomicExch((int32_t*)&tab[slot] + 1, (int32_t)ga);  // little endian: high word = head
        d.inc_next[ga] = prev_head == -2 ? -1 : prev_head;
    }
}

// ---------------------------------------------------------------- sub-warp groups
// The per-read kernels below are latency bound (a chain of dependent metadata loads per read, then a row of ~40
// cells): G lanes per read and 32/G reads in flight per warp multiply the memory-level parallelism.  Every
// collective uses the group's own lane mask, so groups of one warp proceed independently.
template <int G> __device__ __forceinline__ unsigned grp_mask() { return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G)); }
template <int G> __device__ __forceinline__ uint64_t grp_min_u64(uint64_t v, unsigned m) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) { const uint64_t t = __shfl_xor_sync(m, v, o); v = t < v ? t : v; }
    return v;
}

// ---------------------------------------------------------------- K1: project entries
// is value q among the L nodes of the entry?  (whole lane group)
template <int G>
__device__ __forceinline__ bool grp_entry_has(const int32_t* __restrict__ nodes, int L, int32_t q, int gl, unsigned gm) {
    bool f = false;
    for (int x = gl; x < L; x += G) f |= __ldg(nodes + x) == q;
    return __any_sync(gm, f);
}

// one group of G lanes per alignment entry; lanes stride over the entry's nodes.  A node that is the trigger of an
// allele path (hash hit) decides the two containment tests of the reference (is_subset with and without the end
// nodes, :495-548).  Common case, decided by the lane alone: a 3-node path whose end nodes are the neighbours of
// the trigger in the alignment.  Everything else is decided by the whole group scanning the entry.
template <int G>
__global__ void __launch_bounds__(256, 6) k_project(DB d, int64_t e_begin, int64_t e_end) {
    AHS_BAIL_ON_ERR(d);
    const unsigned gm = grp_mask<G>();
    const int lane = lane_id(), gl = lane % G;
    const int64_t gpb = blockDim.x / G, g0 = blockIdx.x * gpb + threadIdx.x / G;
    for (int64_t ge = e_begin + g0; ge < e_end; ge += (int64_t)gridDim.x * gpb) {
        const int c = d.entry_chain[ge];
        const int64_t b0 = d.bubble_off[c];
        const int B = (int)(d.bubble_off[c + 1] - b0);
        if (B <= 1) continue;
        const int rl = d.entry_read[ge];
        const int64_t r = d.read_off[c] + rl;
        const uint32_t el = (uint32_t)(ge - d.entry_off[c]);
        const float ident = d.entry_identity[ge];
        const bool good = good_identity(ident);
        uint16_t* mrow = d.mask + d.mrow_off[c] + (int64_t)rl * B;
        const int32_t* nodes = d.enode + d.enode_off[ge];
        const int L = (int)(d.enode_off[ge + 1] - d.enode_off[ge]);
        const unsigned long long* tab = d.hslots + d.hoff[c];
        const uint32_t mask = d.hmaskc[c];
        uint64_t ck = KEY_NONE, ckA = KEY_NONE;                 // smallest creation keys seen by this lane
        for (int x0 = 0; x0 < L; x0 += G) {
            const int x = x0 + gl;
            int32_t ga = -1, prev = 0, next = 0; bool has_prev = false, has_next = false;
            if (x < L) {
                const int32_t v = __ldg(nodes + x);
                has_prev = x > 0; has_next = x + 1 < L;
                if (has_prev) prev = __ldg(nodes + x - 1);
                if (has_next) next = __ldg(nodes + x + 1);
                uint32_t slot = hash_slot((uint32_t)v, mask);
                while (true) {
                    const unsigned long long sl = tab[slot];
                    if (sl == SLOT_EMPTY) break;
                    if ((uint32_t)sl == (uint32_t)v) { ga = (int32_t)(sl >> 32); break; }
                    slot = (slot + 1) & mask;
                }
            }
            while (__any_sync(gm, ga >= 0)) {
                const bool have = ga >= 0;
                int64_t gb = 0, o = 0; int len = 0;
                bool inner_ok = false, full_ok = false, slow = false;
                int4 rec = make_int4(0, 0, 0, 0);
                if (have) {
                    rec = d.arec[ga];
                    if (rec.x < 0) {                                                      // 3-node path
                        inner_ok = true;                                                  // the only inner node is the trigger
                        full_ok = has_prev && has_next && ((prev == rec.y && next == rec.z) || (prev == rec.z && next == rec.y));
                        slow = !full_ok;                                                  // the end nodes may still be elsewhere in the entry
                    } else slow = true;
                    if (slow) { gb = d.allele_bubble[ga]; o = d.anode_off[ga]; len = (int)(d.anode_off[ga + 1] - o); }
                }
                for (unsigned sm = __ballot_sync(gm, slow); sm; sm &= sm - 1) {
                    const int src_lane = __ffs(sm) - 1;                                   // absolute lane, inside this group
                    const int64_t oo = __shfl_sync(gm, o, src_lane);
                    const int ll = __shfl_sync(gm, len, src_lane);
                    bool in_ok = ll >= 3;                          // len <= 2: universal, handled in k_read_rows
                    for (int y = 2; in_ok && y < ll - 1; y++) in_ok = grp_entry_has<G>(nodes, L, d.anode[oo + y], gl, gm);
                    bool f_ok;
                    if (ll >= 3) f_ok = in_ok && grp_entry_has<G>(nodes, L, d.anode[oo], gl, gm) && grp_entry_has<G>(nodes, L, d.anode[oo + ll - 1], gl, gm);
                    else f_ok = (ll == 1) || grp_entry_has<G>(nodes, L, d.anode[oo + 1], gl, gm);
                    if (lane == src_lane) { inner_ok = in_ok; full_ok = f_ok; }
                }
                if (have) {
                    const int b = (int)((uint32_t)rec.x & 0xfffffu);
                    const int a = (int)(((uint32_t)rec.x >> 20) & 0xffu);
                    const uint32_t bits = ((inner_ok && good) ? (1u << a) : 0u) | (full_ok ? 0x8000u : 0u);
                    if (bits) atomic_or_u16(&mrow[b], (uint16_t)bits);                   // one atomic for both tests
                    if (inner_ok) { const uint64_t k = make_key((uint32_t)b, (uint32_t)a, el); ck = k < ck ? k : ck; }
                    if (full_ok) { const uint64_t k = make_key((uint32_t)rec.w, (uint32_t)a, el); ckA = k < ckA ? k : ckA; }
                    ga = d.inc_next[ga];
                }
            }
        }
        ck = grp_min_u64<G>(ck, gm); ckA = grp_min_u64<G>(ckA, gm);
        if (gl == 0) {
            atomicMin(&d.first_entry[r], el); if (good) d.has_good[r] = 1;
            if (ck != KEY_NONE) atomicMin((unsigned long long*)&d.create_key[r], (unsigned long long)ck);
            if (ckA != KEY_NONE) atomicMin((unsigned long long*)&d.createA_key[r], (unsigned long long)ckA);
        }
    }
}

// ---------------------------------------------------------------- K1b: stage-A statistics per read
// count / first / last fully contained bubble and the stage-A mapq (:146-165, :173-182)
template <int G>
__global__ void __launch_bounds__(256) k_read_stage_a(DB d) {
    AHS_BAIL_ON_ERR(d);
    const unsigned gm = grp_mask<G>();
    const int gl = lane_id() % G;
    const int64_t gpb = blockDim.x / G, g0 = blockIdx.x * gpb + threadIdx.x / G;
    for (int64_t r = g0; r < d.NR; r += (int64_t)gridDim.x * gpb) {
        const int c = d.read_chain[r];
        const int B = (int)(d.bubble_off[c + 1] - d.bubble_off[c]);
        if (B <= 1) continue;
        const uint16_t* mrow = d.mask + d.mrow_off[c] + (r - d.read_off[c]) * B;
        int cnt = 0, first = INT32_MAX, last = -1;
        for (int b = gl; b < B; b += G) if (mrow[b] & 0x8000u) { cnt++; first = min(first, b); last = max(last, b); }
        cnt = __reduce_add_sync(gm, cnt); first = __reduce_min_sync(gm, first); last = __reduce_max_sync(gm, last);
        if (gl == 0) {
            int mapq = 0;
            if (cnt > 0) {
                const uint32_t el = (uint32_t)(d.createA_key[r] & 0xffffffffu);
                mapq = mapq_of(d.entry_identity[d.entry_off[c] + el]);
                atomicMax(&d.ch_maxpos[c], last);
            }
            d.rdA_cnt[r] = cnt; d.rdA_first[r] = first; d.rdA_last[r] = last; d.rdA_mapq[r] = mapq;
        }
    }
}

// boundary flags: which of maxpos-1 / maxpos are last / first positions of filtered stage-A reads (:173-189)
__global__ void k_chain_flags(DB d) {
    AHS_BAIL_ON_ERR(d);
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < d.NR; r += (int64_t)gridDim.x * blockDim.x) {
        const int c = d.read_chain[r];
        if (d.bubble_off[c + 1] - d.bubble_off[c] <= 1) continue;
        if (!(d.rdA_cnt[r] > 1 && d.rdA_mapq[r] >= 93)) continue;
        const int mp = d.ch_maxpos[c];
        int f = 0;
        if (d.rdA_last[r] == mp) f |= 1;
        if (d.rdA_last[r] == mp - 1) f |= 2;
        if (d.rdA_first[r] == mp) f |= 4;
        if (d.rdA_first[r] == mp - 1) f |= 8;
        if (f) atomicOr(&d.ch_flags[c], f);
    }
}

// to_be_added = [0,maxpos) U {e, e+1 : e in last \ first} = [0, T)  (:173-209, SURVEY A#10)
__global__ void k_chain_T(DB d) {
    AHS_BAIL_ON_ERR(d);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d.C; c += gridDim.x * blockDim.x) {
        const int B = (int)(d.bubble_off[c + 1] - d.bubble_off[c]);
        if (B <= 1) { d.ch_status[c] = AHS_CHAIN_TRIVIAL; d.ch_T[c] = 0; continue; }
        const int mp = d.ch_maxpos[c];
        if (mp < 0) { d.ch_status[c] = AHS_CHAIN_EMPTY; d.ch_T[c] = 0; continue; }   // reference: UB (:193)
        const int f = d.ch_flags[c];
        const bool e_max = (f & 1) && !(f & 4);          // maxpos   in last \ first  -> adds maxpos, maxpos+1
        const bool e_m1  = (f & 2) && !(f & 8);          // maxpos-1 in last \ first  -> adds maxpos-1, maxpos
        int T = mp;
        if (e_max) T = mp + 2; else if (e_m1) T = mp + 1;
        d.ch_T[c] = min(T, B);                             // bubble ids >= B have no alleles (:216 default-insert)
        d.ch_status[c] = AHS_CHAIN_OK;
    }
}

// ---------------------------------------------------------------- K1d: final rows per read (stage B + filter)
template <int G>
__global__ void __launch_bounds__(256) k_read_rows(DB d) {
    AHS_BAIL_ON_ERR(d);
    const unsigned gm = grp_mask<G>();
    const int gl = lane_id() % G;
    const int64_t gpb = blockDim.x / G, g0 = blockIdx.x * gpb + threadIdx.x / G;
    long long cells_local = 0;
    for (int64_t r = g0; r < d.NR; r += (int64_t)gridDim.x * gpb) {
        const int c = d.read_chain[r];
        if (gl == 0) d.rd_pass[r] = 0;
        if (d.ch_status[c] != AHS_CHAIN_OK) continue;
        const int64_t b0g = d.bubble_off[c];
        const int B = (int)(d.bubble_off[c + 1] - b0g);
        const int T = d.ch_T[c];
        uint16_t* mrow = d.mask + d.mrow_off[c] + (r - d.read_off[c]) * B;
        const uint32_t fe = d.first_entry[r];
        const bool has_entry = fe != 0xffffffffu;
        const bool hg = d.has_good[r] != 0;
        const bool univ = (d.ch_flags[c] & CH_HAS_UNIV) != 0;              // rare: a bubble with an allele path of <= 2 nodes
        // creation triple: minimum (position, allele, entry) over all matches, universal alleles included
        uint64_t ck = d.create_key[r];
        if (has_entry && univ) for (int b = gl; b < T; b += G) {
            const uint32_t u = d.bubble_univ[b0g + b];
            if (u != 0xffffffffu) { uint64_t k = make_key((uint32_t)b, u, fe); ck = k < ck ? k : ck; }
        }
        ck = grp_min_u64<G>(ck, gm);
        const int bc = (ck == KEY_NONE) ? INT32_MAX : (int)(ck >> 40);
        int nv = 0, last = -1;
        uint32_t cov = 0;                                                   // this lane's covered positions b = gl + G j, j < 32
        if (bc < T) {
            const int ac = (int)((ck >> 32) & 0xff);
            for (int b = gl; b < B; b += G) {
                uint32_t code = 0;
                if (b < T) {
                    uint32_t m = mrow[b] & 0x7fffu;
                    const uint32_t u = univ ? d.bubble_univ[b0g + b] : 0xffffffffu;
                    if (u != 0xffffffffu && hg) m |= 1u << u;
                    if (b == bc) code = (uint32_t)ac + 1u;
                    else if (m) code = (uint32_t)__ffs((int)m);             // lowest set bit = first matching allele
                }
                mrow[b] = (uint16_t)code;
                if (code) { nv++; last = max(last, b); if (b < 32 * G) cov |= 1u << (b / G); }
            }
            nv = __reduce_add_sync(gm, nv); last = __reduce_max_sync(gm, last);
        }
        int mapq = 0; bool pass = false;
        if (bc < T) {
            mapq = mapq_of(d.entry_identity[d.entry_off[c] + (uint32_t)(ck & 0xffffffffu)]);
            pass = nv > 1 && mapq >= 93;                                    // :270
        }
        __syncwarp(gm);                                                     // the group's codes are written
        if (pass) {
            for (uint32_t m = cov; m; m &= m - 1) d.poscov[b0g + gl + G * (__ffs((int)m) - 1)] = 1;   // codes are only set below T
            for (int b = gl + 32 * G; b < T; b += G) if (mrow[b]) d.poscov[b0g + b] = 1;
        }
        if (gl == 0) {
            d.rd_nv[r] = nv; d.rd_first[r] = bc; d.rd_last[r] = last; d.rd_mapq[r] = mapq; d.rd_pass[r] = pass ? 1 : 0;
            d.create_key[r] = ck;
            if (pass) { atomicAdd(&d.ch_nfinal[c], 1); atomicAdd(&d.ch_cells[c], (unsigned long long)nv); cells_local += nv; }
        }
    }
    // one atomic per warp on the batch-wide cell counter (a per-read atomic on one address serialises in L2)
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) cells_local += __shfl_xor_sync(0xffffffffu, cells_local, o);
    if (lane_id() == 0 && cells_local) atomicAdd((unsigned long long*)d.tot_cells, (unsigned long long)cells_local);
}

// ---------------------------------------------------------------- K1e: read order
// Insertion order of the stage-B read set = order of the creation triples; rank by counting.
__global__ void k_read_rank(DB d) {
    AHS_BAIL_ON_ERR(d);
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < d.NR; r += (int64_t)gridDim.x * blockDim.x) {
        if (!d.rd_pass[r]) continue;
        const int c = d.read_chain[r];
        const int64_t r0 = d.read_off[c], r1 = d.read_off[c + 1];
        const uint64_t key = d.create_key[r];
        int rank = 0;
        for (int64_t x = r0; x < r1; x++) rank += (d.rd_pass[x] && d.create_key[x] < key) ? 1 : 0;
        d.ord[r0 + rank] = (int32_t)(r - r0);
        d.okey[r0 + rank] = d.rd_first[r];
    }
}

// libstdc++ std::sort (introsort + final insertion sort) replayed move for move on (key, value)
// pairs with a comparator that looks at the key only: ReadSet::sort() at alignmentstoreadset.cpp:297
// (comparator on firstPosition()), and the cluster sort at :720.  SURVEY A#11, A#22.
struct KV { int32_t* k; int32_t* v; };
__device__ __forceinline__ void kv_swap(KV a, int x, int y) {
    int32_t t = a.k[x]; a.k[x] = a.k[y]; a.k[y] = t; t = a.v[x]; a.v[x] = a.v[y]; a.v[y] = t;
}
template <bool DESC> __device__ __forceinline__ bool kv_less(int32_t x, int32_t y) { return DESC ? x > y : x < y; }

template <bool DESC>
__device__ void kv_unguarded_linear_insert(KV a, int last) {
    int32_t vk = a.k[last], vv = a.v[last];
    int next = last - 1;
    while (kv_less<DESC>(vk, a.k[next])) { a.k[last] = a.k[next]; a.v[last] = a.v[next]; last = next; --next; }
    a.k[last] = vk; a.v[last] = vv;
}
template <bool DESC>
__device__ void kv_insertion_sort(KV a, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (kv_less<DESC>(a.k[i], a.k[first])) {
            int32_t vk = a.k[i], vv = a.v[i];
            for (int x = i; x > first; --x) { a.k[x] = a.k[x - 1]; a.v[x] = a.v[x - 1]; }
            a.k[first] = vk; a.v[first] = vv;
        } else kv_unguarded_linear_insert<DESC>(a, i);
    }
}
// std::__adjust_heap + std::__push_heap (bits/stl_heap.h) on the sub-array starting at `first`
template <bool DESC>
__device__ void kv_adjust_heap(KV a, int first, int hole, int len, int32_t vk, int32_t vv) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (kv_less<DESC>(a.k[first + child], a.k[first + child - 1])) child--;
        a.k[first + hole] = a.k[first + child]; a.v[first + hole] = a.v[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a.k[first + hole] = a.k[first + child - 1]; a.v[first + hole] = a.v[first + child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && kv_less<DESC>(a.k[first + parent], vk)) {
        a.k[first + hole] = a.k[first + parent]; a.v[first + hole] = a.v[first + parent];
        hole = parent; parent = (hole - 1) / 2;
    }
    a.k[first + hole] = vk; a.v[first + hole] = vv;
}
// std::__partial_sort(first, last, last): __heap_select (= __make_heap, the selection loop is empty) + __sort_heap —
// what __introsort_loop falls back to when its depth limit is used up
template <bool DESC>
__device__ void kv_heap_sort(KV a, int first, int last) {
    const int len = last - first;
    if (len < 2) return;
    for (int parent = (len - 2) / 2; ; parent--) {
        kv_adjust_heap<DESC>(a, first, parent, len, a.k[first + parent], a.v[first + parent]);
        if (parent == 0) break;
    }
    for (int end = last; end - first > 1; ) {
        --end;
        const int32_t vk = a.k[end], vv = a.v[end];
        a.k[end] = a.k[first]; a.v[end] = a.v[first];
        kv_adjust_heap<DESC>(a, first, 0, end - first, vk, vv);
    }
}
// std::sort(first, last, comp) of libstdc++ (bits/stl_algo.h: __introsort_loop with depth limit 2 lg n, then
// __final_insertion_sort), heap-sort fall-back included
template <bool DESC>
__device__ void kv_std_sort(KV a, int n) {
    if (n <= 1) return;
    int depth0 = 0; for (int m = n; m > 1; m >>= 1) depth0++;
    depth0 *= 2;
    int stk_first[64], stk_last[64], stk_depth[64]; int sp = 0;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = depth0; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > 16) {
            if (depth == 0) { kv_heap_sort<DESC>(a, first, last); break; }
            --depth;
            // __move_median_to_first(first, first+1, mid, last-1)
            int mid = first + (last - first) / 2, A = first + 1, Bm = mid, Cc = last - 1;
            if (kv_less<DESC>(a.k[A], a.k[Bm])) {
                if (kv_less<DESC>(a.k[Bm], a.k[Cc])) kv_swap(a, first, Bm);
                else if (kv_less<DESC>(a.k[A], a.k[Cc])) kv_swap(a, first, Cc);
                else kv_swap(a, first, A);
            } else if (kv_less<DESC>(a.k[A], a.k[Cc])) kv_swap(a, first, A);
            else if (kv_less<DESC>(a.k[Bm], a.k[Cc])) kv_swap(a, first, Cc);
            else kv_swap(a, first, Bm);
            // __unguarded_partition(first+1, last, pivot=first)
            int lo = first + 1, hi = last;
            const int32_t pk = a.k[first];
            while (true) {
                while (kv_less<DESC>(a.k[lo], pk)) ++lo;
                --hi;
                while (kv_less<DESC>(pk, a.k[hi])) --hi;
                if (!(lo < hi)) break;
                kv_swap(a, lo, hi);
                ++lo;
            }
            // recurse on [cut,last), loop on [first,cut): the right part is sorted first, but the
            // parts are disjoint, so the order of processing does not change the result.  The stack
            // cannot overflow: every push lowers `depth`, so at most depth0 <= 2 * 31 ranges are pending.
            stk_first[sp] = lo; stk_last[sp] = last; stk_depth[sp] = depth; ++sp;
            last = lo;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        kv_insertion_sort<DESC>(a, 0, 16);
        for (int i = 16; i < n; ++i) kv_unguarded_linear_insert<DESC>(a, i);
    } else kv_insertion_sort<DESC>(a, 0, n);
}

// diagnostic kernel behind ahs_debug_std_sort: one thread replays std::sort on (key, value) pairs
__global__ void k_debug_std_sort(int32_t* k, int32_t* v, int n, int desc) {
    KV a; a.k = k; a.v = v;
    if (desc) kv_std_sort<true>(a, n); else kv_std_sort<false>(a, n);
}

__global__ void k_chain_sort(DB d) {
    AHS_BAIL_ON_ERR(d);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d.C; c += gridDim.x * blockDim.x) {
        if (d.ch_status[c] != AHS_CHAIN_OK) continue;
        const int n = d.ch_nfinal[c];
        if (n == 0) { d.ch_status[c] = AHS_CHAIN_EMPTY; continue; }            // :279-282
        KV a; a.k = d.okey + d.read_off[c]; a.v = d.ord + d.read_off[c];
        kv_std_sort<false>(a, n);
        // covered positions (ReadSet::get_positions, :317)
    }
}

// covered positions per chain: count, then (after the host has the offsets) the compact list
__global__ void k_count_pos(DB d) {
    AHS_BAIL_ON_ERR(d);
    const int wpb = blockDim.x >> 5, lane = lane_id();
    for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < d.C; c += gridDim.x * wpb) {
        const int64_t b0 = d.bubble_off[c];
        const int B = (int)(d.bubble_off[c + 1] - b0);
        int n = 0;
        for (int b = lane; b < B; b += 32) n += d.poscov[b0 + b] ? 1 : 0;
        n = warp_sum_i32(n);
        if (lane == 0) d.ch_npos[c] = n;
    }
}

__global__ void k_compact_pos(DB d) {
    const int wpb = blockDim.x >> 5, lane = lane_id();
    for (int c = blockIdx.x * wpb + (threadIdx.x >> 5); c < d.C; c += gridDim.x * wpb) {
        const int64_t b0 = d.bubble_off[c];
        const int B = (int)(d.bubble_off[c + 1] - b0);
        const bool live = d.ch_status[c] == AHS_CHAIN_OK;
        int base = 0;
        for (int bb = 0; bb < B; bb += 32) {
            const int b = bb + lane;
            const bool cov = live && b < B && d.poscov[b0 + b];
            const unsigned m = __ballot_sync(0xffffffffu, cov);
            if (b < B) d.pos_compact[b0 + b] = cov ? base + __popc(m & ((1u << lane) - 1u)) : -1;
            if (cov) d.pos[d.pos_off[c] + base + __popc(m & ((1u << lane) - 1u))] = b;
            base += __popc(m);
        }
    }
}

// ---------------------------------------------------------------- K1f: pack final rows
// G lanes per final read: mask row (u16 codes) -> packed 2/4-bit codes, dense by bubble id.  Lanes read consecutive
// codes (coalesced), shift them into place and OR them together across the group, one word at a time.
template <int G>
__global__ void __launch_bounds__(256) k_pack_rows(DB d) {
    const unsigned gm = grp_mask<G>();
    const int gl = lane_id() % G;
    const int64_t gpb = blockDim.x / G, g0 = blockIdx.x * gpb + threadIdx.x / G;
    const int bits = d.bits, per_word = 32 / bits;
    for (int64_t f = g0; f < d.NF; f += (int64_t)gridDim.x * gpb) {
        const int c = d.fr_chain[f];
        const int B = (int)(d.bubble_off[c + 1] - d.bubble_off[c]);
        const int i = (int)(f - d.frow_off[c]);
        const int rl = d.ord[d.read_off[c] + i];
        const int64_t r = d.read_off[c] + rl;
        const uint16_t* mrow = d.mask + d.mrow_off[c] + (int64_t)rl * B;
        uint32_t* row = d.codes + d.code_off[c] + (int64_t)i * d.ch_words[c];
        const int first = d.rd_first[r], last = d.rd_last[r];
        for (int w = first / per_word; w <= last / per_word; w++) {
            uint32_t word = 0;
            for (int x = gl; x < per_word; x += G) { const int b = w * per_word + x; if (b < B) word |= (uint32_t)mrow[b] << (x * bits); }
            word = __reduce_or_sync(gm, word);
            if (gl == 0) row[w] = word;
        }
        if (gl == 0) {
            d.fr_first[f] = first; d.fr_last[f] = last; d.fr_mapq[f] = d.rd_mapq[r]; d.fr_id[f] = rl; d.fr_nv[f] = d.rd_nv[r];
            atomicMax(&d.ch_maxspan[c], last - first);
        }
    }
}

}  // namespace ahs
