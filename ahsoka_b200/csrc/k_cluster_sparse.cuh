// k_cluster_sparse.cuh — cluster editing (rule R2) of chains ABOVE CC_MAXN final reads: edge slots, per-node slot lists
// and a two-level maximum tree, so that one greedy step touches only the edges around the two nodes it works on.
//
// Replaces ClusterEditingSolver(sim,false).run() (call site reference src/alignmentstoreadset.cpp:312-315; algorithm:
// oracle/core/phase_core.hpp rule R2) where the shared-memory kernel of k_chain.cuh does not fit: BASELINE config 4
// (200-330 reads per chain), config 1 (one chain of 1,400 reads), config 5 (chains of up to 10 k bubbles: ~35,000 reads).
// A long chain's graph is banded — a read overlaps ~2 x depth others — and stays local while clusters grow, so the work
// of a step is bounded by the neighbourhoods of the two end nodes, not by the chain.
//
// State of one chain (HBM; one block runs its greedy loop: 1024 threads above SP_SMALL_N reads, 256 threads — five blocks to
// an SM — below, where a batch has many such chains):
//   W[n][n]      int32 dense weights (0 = no edge, CC_FORB = forbidden): O(1) third-side look-ups
//   slots        one per edge that ever had a non-zero weight: key (a << 16 | b, a < b, CURRENT node ids), flags, icf, icp
//                (int64).  A forbidden edge keeps its slot (flag SPF_FORB): it still counts in the icp of its neighbours.
//   lists        list(x) = slot indices of the edges at node x; dead slots are skipped and dropped in place by the walk of a merge.
//                When b merges into a, the slot of (b,x) is relabelled (a,x) or dies in favour of (a,x), and list(a) is
//                written afresh (bump pool).
//   tree         leaf = maxima of 64 consecutive slots, level 2 = maxima of 64 leaves; a step marks the leaves of the
//                slots it changed; they are queued, dealt round-robin over the warps and recomputed, then their level-2
//                entries, then level 2 is reduced (warp maxima by redux.sync).
// Same decisions as k_cluster_chain: merge the pair of largest icf if it is >= the largest icp, else forbid — runs of
// forbids on negative edges batched exactly (validated against the sequential definition, rolled back otherwise).
#pragma once
#include "common.cuh"
#include "device_batch.cuh"
#include "k_chain.cuh"

namespace ahs {

constexpr int SP_THREADS = 1024;                     // chains above SP_SMALL_N reads; the others run in blocks of SP_THREADS_SMALL
constexpr int SP_THREADS_SMALL = 256;
constexpr int SP_SMALL_N = 1024;
constexpr uint32_t SP_NONE = 0xffffffffu;
constexpr uint8_t SPF_POS = 1, SPF_DEAD = 2, SPF_FORB = 4, SPF_FLAG = 8;

struct SpBest {                                      // maxima over a set of slots; keys break ties (smallest pair first)
    long long M, maxP, maxPpos, maxPneg; uint32_t kF, kP;
    __device__ __forceinline__ void clear() { M = -1; maxP = -1; maxPpos = -1; maxPneg = -1; kF = SP_NONE; kP = SP_NONE; }
    __device__ __forceinline__ void consider(uint32_t key, uint8_t fl, long long f, long long p) {
        if (fl & (SPF_DEAD | SPF_FORB)) return;      // not a candidate
        if (f > M || (f == M && key < kF)) { M = f; kF = key; }
        if (p > maxP || (p == maxP && key < kP)) { maxP = p; kP = key; }
        if (fl & SPF_POS) { if (p > maxPpos) maxPpos = p; } else if (p > maxPneg) maxPneg = p;
    }
    __device__ __forceinline__ void merge(const SpBest& o) {
        if (o.M > M || (o.M == M && o.kF < kF)) { M = o.M; kF = o.kF; }
        if (o.maxP > maxP || (o.maxP == maxP && o.kP < kP)) { maxP = o.maxP; kP = o.kP; }
        if (o.maxPpos > maxPpos) maxPpos = o.maxPpos;
        if (o.maxPneg > maxPneg) maxPneg = o.maxPneg;
    }
};
// warp maxima with redux.sync: a 64-bit signed maximum is the maximum of the high words, then of the (unsigned) low words among the
// lanes that hold it; a tie key is the smallest key among the lanes that hold the maximum.  ~35 instructions instead of the ~150
// of a five-level shuffle tree over the six fields.
__device__ __forceinline__ long long sp_warp_max_i64(long long v) {
    const int hi = (int)(v >> 32); const unsigned lo = (unsigned)v;
    const int mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return (long long)(((unsigned long long)(unsigned)mh << 32) | ml);
}
__device__ __forceinline__ SpBest sp_warp_reduce(SpBest b) {
    SpBest r;
    r.M = sp_warp_max_i64(b.M); r.kF = __reduce_min_sync(0xffffffffu, b.M == r.M ? b.kF : SP_NONE);
    r.maxP = sp_warp_max_i64(b.maxP); r.kP = __reduce_min_sync(0xffffffffu, b.maxP == r.maxP ? b.kP : SP_NONE);
    r.maxPpos = sp_warp_max_i64(b.maxPpos); r.maxPneg = sp_warp_max_i64(b.maxPneg);
    return r;
}

// one big chain: where its state lives (filled by the host, see phase_batch.cu)
struct SpChain {
    int32_t chain, n;
    int64_t w_off;                                   // d.W: dense n x n
    int64_t slot_off; int32_t slot_cap, n_leaf, n_sup;      // slot arrays; leaves = ceil(slot_cap / 64), level 2 = ceil(leaves / 64)
    int64_t list_off, list_cap;                      // pool of slot indices (u32): the initial lists first, the bump area behind
    int64_t node_off;                                // node arrays (n entries each)
    int64_t leaf_off, sup_off;                       // SpBest entries
};
struct SpArrays {
    const SpChain* chains; int n_chains;
    int64_t n_nodes, n_slot_cap, n_leaves, n_sups;   // totals over the chains (the chains' ranges are consecutive)
    uint32_t* fl_slot; int32_t* fl_old;              // per slot capacity: the edges flagged in a forbid round (never truncated)
    uint32_t* supq;                                  // per level-2 entry: the entries to visit in a round
    uint32_t* leafq;                                 // per leaf: the dirty leaves of a refresh
    uint32_t* key; uint8_t* flag; long long *F, *P;  // slots
    uint32_t* pool;                                  // lists
    long long* lptr; uint32_t *llen, *sa, *sb, *up; int32_t *wa, *wb, *nw; long long *frF, *frP;      // per node
    SpBest *leaf, *sup;
    long long* bump;                                 // per chain: next free entry of the pool (relative to list_off)
    int32_t* n_slots;                                // per chain: slots in use
};

__device__ __forceinline__ long long sp_tf(int x, int y) { return (long long)max(min(x, y), 0); }
__device__ __forceinline__ long long sp_tp(int x, int y) { const int lo = min(x, y), hi = max(x, y); return (long long)max(min(hi, -lo), 0); }
__device__ __forceinline__ int sp_other(uint32_t key, int x) { const int p = (int)(key >> 16), q = (int)(key & 0xffffu); return p == x ? q : p; }

// chain that owns element x of a per-chain range array (ranges consecutive, `off` = member pointer to the range start)
template <class F> __device__ __forceinline__ int sp_owner(const SpArrays& sp, int64_t x, F off) {
    int lo = 0, hi = sp.n_chains - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (off(sp.chains[mid]) <= x) lo = mid; else hi = mid - 1; }
    return lo;
}

// ---------------------------------------------------------------- set-up, grid wide
// degrees: up[x] = edges (x,y), y > x; llen[x] = all edges at x.  One warp per read, lanes over its partner band.
__global__ void __launch_bounds__(256) k_sp_degrees(DB d, SpArrays sp) {
    const int wpb = blockDim.x >> 5, lane = lane_id();
    for (int64_t g = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); g < sp.n_nodes; g += (int64_t)gridDim.x * wpb) {
        const SpChain ch = sp.chains[sp_owner(sp, g, [](const SpChain& c) { return c.node_off; })];
        const int n = ch.n, x = (int)(g - ch.node_off); const int32_t* W = d.W + ch.w_off;
        const int64_t f0 = d.frow_off[ch.chain];
        const int32_t* first = d.fr_first + f0; const int32_t* lastp = d.fr_last + f0;
        int lo, hi; partner_band(d, ch.chain, x, n, first, first[x], lastp[x], lo, hi);
        int nu = 0, na = 0;
        for (int y = lo + lane; y <= hi; y += 32) if (y != x && W[(int64_t)x * n + y] != 0) { na++; nu += y > x ? 1 : 0; }
        nu = warp_sum_i32(nu); na = warp_sum_i32(na);
        if (lane == 0) { sp.up[g] = (uint32_t)nu; sp.llen[g] = (uint32_t)na; }
    }
}

// exclusive scans per chain (one block per chain): first slot of every row, first list entry of every node
__global__ void __launch_bounds__(1024) k_sp_scan(SpArrays sp) {
    __shared__ long long s_a[1024], s_b[1024];
    const SpChain ch = sp.chains[blockIdx.x];
    const int n = ch.n, tid = threadIdx.x;
    const int per = (n + 1023) / 1024, x0 = tid * per, x1 = min(n, x0 + per);
    long long a = 0, b = 0;
    for (int x = x0; x < x1; x++) { a += sp.up[ch.node_off + x]; b += sp.llen[ch.node_off + x]; }
    s_a[tid] = a; s_b[tid] = b;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const long long ta = tid >= o ? s_a[tid - o] : 0, tb = tid >= o ? s_b[tid - o] : 0;
        __syncthreads();
        s_a[tid] += ta; s_b[tid] += tb;
        __syncthreads();
    }
    long long ra = s_a[tid] - a, rb = s_b[tid] - b;
    for (int x = x0; x < x1; x++) {
        const uint32_t u = sp.up[ch.node_off + x], l = sp.llen[ch.node_off + x];
        sp.up[ch.node_off + x] = (uint32_t)ra;       // row start (slot index)
        sp.lptr[ch.node_off + x] = rb;               // list start (pool index, relative)
        sp.llen[ch.node_off + x] = 0;                // becomes the fill cursor
        sp.sa[ch.node_off + x] = SP_NONE; sp.sb[ch.node_off + x] = SP_NONE;
        ra += u; rb += l;
    }
    if (tid == 1023) { sp.n_slots[blockIdx.x] = (int32_t)s_a[1023]; sp.bump[blockIdx.x] = s_b[1023]; }
}

// slots in row-major order (x ascending, y ascending) and the lists
__global__ void __launch_bounds__(256) k_sp_fill(DB d, SpArrays sp) {
    const int wpb = blockDim.x >> 5, lane = lane_id();
    for (int64_t g = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); g < sp.n_nodes; g += (int64_t)gridDim.x * wpb) {
        const SpChain ch = sp.chains[sp_owner(sp, g, [](const SpChain& c) { return c.node_off; })];
        const int n = ch.n, x = (int)(g - ch.node_off); const int32_t* W = d.W + ch.w_off;
        const int64_t f0 = d.frow_off[ch.chain];
        const int32_t* first = d.fr_first + f0; const int32_t* lastp = d.fr_last + f0;
        uint32_t* pool = sp.pool + ch.list_off;
        int lo, hi; partner_band(d, ch.chain, x, n, first, first[x], lastp[x], lo, hi);
        uint32_t base = sp.up[g];
        for (int y0 = x + 1; y0 <= hi; y0 += 32) {
            const int y = y0 + lane;
            const int w = y <= hi ? W[(int64_t)x * n + y] : 0;
            const unsigned bal = __ballot_sync(0xffffffffu, w != 0);
            if (w != 0) {
                const uint32_t s = base + __popc(bal & ((1u << lane) - 1u));
                sp.key[ch.slot_off + s] = ((uint32_t)x << 16) | (uint32_t)y; sp.flag[ch.slot_off + s] = w > 0 ? SPF_POS : 0;
                pool[sp.lptr[ch.node_off + x] + atomicAdd(&sp.llen[ch.node_off + x], 1u)] = s;
                pool[sp.lptr[ch.node_off + y] + atomicAdd(&sp.llen[ch.node_off + y], 1u)] = s;
            }
            base += __popc(bal);
        }
    }
}

// initial induced costs: one warp per slot, lanes over the list of its first node (a common neighbour is in both lists)
__global__ void __launch_bounds__(256) k_sp_init_costs(DB d, SpArrays sp) {
    const int wpb = blockDim.x >> 5, lane = lane_id();
    for (int64_t g = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); g < sp.n_slot_cap; g += (int64_t)gridDim.x * wpb) {
        const int ci = sp_owner(sp, g, [](const SpChain& c) { return c.slot_off; });
        const SpChain ch = sp.chains[ci];
        const int s = (int)(g - ch.slot_off);
        if (s >= sp.n_slots[ci]) continue;
        const int n = ch.n; const int32_t* W = d.W + ch.w_off;
        const uint32_t* pool = sp.pool + ch.list_off;
        const uint32_t key = sp.key[g];
        const int x = (int)(key >> 16), y = (int)(key & 0xffffu);
        const int w = W[(int64_t)x * n + y];
        const long long lp = sp.lptr[ch.node_off + x]; const int ll = (int)sp.llen[ch.node_off + x];
        long long f = 0, p = 0;
        for (int i = lane; i < ll; i += 32) {
            const int t = sp_other(sp.key[ch.slot_off + pool[lp + i]], x);
            if (t == y) continue;
            const int wx = W[(int64_t)x * n + t], wy = W[(int64_t)y * n + t];
            f += sp_tf(wx, wy); p += sp_tp(wx, wy);
        }
        f = warp_sum_i64(f); p = warp_sum_i64(p);
        if (lane == 0) { sp.F[g] = f + max(w, 0); sp.P[g] = p + max(-w, 0); }
    }
}

// maxima of one leaf (64 slots), by one warp
__device__ __forceinline__ SpBest sp_leaf_maxima(const SpArrays& sp, const SpChain& ch, int leaf, int n_slots, int lane) {
    SpBest b; b.clear();
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const int s = leaf * 64 + u * 32 + lane;
        if (s < n_slots) b.consider(sp.key[ch.slot_off + s], sp.flag[ch.slot_off + s], sp.F[ch.slot_off + s], sp.P[ch.slot_off + s]);
    }
    return sp_warp_reduce(b);
}
__device__ __forceinline__ SpBest sp_sup_maxima(const SpArrays& sp, const SpChain& ch, int su, int lane) {
    SpBest b; b.clear();
#pragma unroll
    for (int u = 0; u < 2; u++) { const int l = su * 64 + u * 32 + lane; if (l < ch.n_leaf) b.merge(sp.leaf[ch.leaf_off + l]); }
    return sp_warp_reduce(b);
}

__global__ void __launch_bounds__(256) k_sp_leaves(SpArrays sp) {
    const int wpb = blockDim.x >> 5, lane = lane_id();
    for (int64_t g = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); g < sp.n_leaves; g += (int64_t)gridDim.x * wpb) {
        const int ci = sp_owner(sp, g, [](const SpChain& c) { return c.leaf_off; });
        const SpChain ch = sp.chains[ci];
        const SpBest b = sp_leaf_maxima(sp, ch, (int)(g - ch.leaf_off), sp.n_slots[ci], lane);
        if (lane == 0) sp.leaf[g] = b;
    }
}
__global__ void __launch_bounds__(256) k_sp_sups(SpArrays sp) {
    const int wpb = blockDim.x >> 5, lane = lane_id();
    for (int64_t g = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); g < sp.n_sups; g += (int64_t)gridDim.x * wpb) {
        const SpChain ch = sp.chains[sp_owner(sp, g, [](const SpChain& c) { return c.sup_off; })];
        const SpBest b = sp_sup_maxima(sp, ch, (int)(g - ch.sup_off), lane);
        if (lane == 0) sp.sup[g] = b;
    }
}

// ---------------------------------------------------------------- the greedy loop, one block per chain
__host__ __device__ inline size_t sp_smem_bytes(int nmax, int max_leaf) {
    size_t b = 32 * sizeof(SpBest) + 256;                           // reduction scratch, scalars
    b += ((size_t)(nmax + 31) / 32) * 4;                            // inS bit set
    b += ((size_t)(max_leaf + 31) / 32) * 4 + ((size_t)(max_leaf / 64 + 32) / 32) * 4;      // dirty leaves, dirty level-2 entries
    b += (size_t)nmax * 2;                                          // labels
    return (b + 15) & ~(size_t)15;
}

// chains [c_begin, c_end) of sp.chains (sorted by decreasing read count), one block of NT threads per chain at a time
template <int NT>
__global__ void __launch_bounds__(NT, NT >= 1024 ? 1 : 5) k_cluster_sparse(DB d, SpArrays sp, int c_begin, int c_end, int nmax, int max_leaf, int32_t* __restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char sp_sm[];
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    SpBest* red; int32_t* scal; uint32_t *inS, *dleaf, *dsup; uint16_t* label;
    const int w_ins = (nmax + 31) / 32, w_leaf = (max_leaf + 31) / 32, w_sup = (max_leaf / 64 + 32) / 32;
    {
        unsigned char* p = sp_sm;
        red = (SpBest*)p; p += 32 * sizeof(SpBest);
        scal = (int32_t*)p; p += 256;       // [0] item [1] |S| [2] flagged edges [3] queued level-2 entries [4] new list cursor
        inS = (uint32_t*)p; p += (size_t)w_ins * 4; dleaf = (uint32_t*)p; p += (size_t)w_leaf * 4; dsup = (uint32_t*)p; p += (size_t)w_sup * 4;
        label = (uint16_t*)p;
    }
    while (true) {
        __syncthreads();
        if (tid == 0) scal[0] = atomicAdd(work_counter, 1);
        __syncthreads();
        const int ci = c_begin + scal[0];
        if (ci >= c_end) break;
        const SpChain ch = sp.chains[ci];
        const int n = ch.n, n_slots = sp.n_slots[ci];
        int32_t* W = d.W + ch.w_off;
        uint32_t* key = sp.key + ch.slot_off; uint8_t* flag = sp.flag + ch.slot_off; long long* F = sp.F + ch.slot_off; long long* P = sp.P + ch.slot_off;
        uint32_t* pool = sp.pool + ch.list_off;
        long long* lptr = sp.lptr + ch.node_off; uint32_t* llen = sp.llen + ch.node_off; uint32_t* sa = sp.sa + ch.node_off; uint32_t* sb = sp.sb + ch.node_off;
        uint32_t* slist = sp.up + ch.node_off;                        // the row starts are not needed any more: the list S of a merge
        int32_t* wa = sp.wa + ch.node_off; int32_t* wb = sp.wb + ch.node_off; int32_t* nw = sp.nw + ch.node_off;
        long long* frF = sp.frF + ch.node_off; long long* frP = sp.frP + ch.node_off;
        SpBest* leaf = sp.leaf + ch.leaf_off; SpBest* sup = sp.sup + ch.sup_off;
        uint32_t* fl_slot = sp.fl_slot + ch.slot_off; int32_t* fl_old = sp.fl_old + ch.slot_off; uint32_t* supq = sp.supq + ch.sup_off; uint32_t* leafq = sp.leafq + ch.leaf_off;
        for (int x = tid; x < n; x += NT) label[x] = (uint16_t)x;
        for (int x = tid; x < w_ins; x += NT) inS[x] = 0;
        for (int x = tid; x < w_leaf; x += NT) dleaf[x] = 0;
        for (int x = tid; x < w_sup; x += NT) dsup[x] = 0;
        if (tid == 0) { scal[1] = 0; scal[2] = 0; scal[3] = 0; scal[4] = 0; scal[6] = 0; }
        int n_active = n;
        bool force_single = false;
        __syncthreads();
        auto mark = [&](uint32_t s) { atomicOr(&dleaf[s >> 11], 1u << ((s >> 6) & 31)); };
        // leaves of the changed slots, then their level-2 entries, then the block-wide maxima.  Three barriers.
        auto refresh = [&]() -> SpBest {
            __syncthreads();
            // dirty leaves come in runs (the slots of neighbouring nodes): they are queued first (a thread per 32-leaf word) and
            // then dealt round-robin over the warps — a run would otherwise land on the one warp that owns its word
            for (int w = tid; w < w_leaf; w += NT) {
                uint32_t bits = dleaf[w];
                if (!bits) continue;
                int q = atomicAdd(&scal[6], __popc(bits));
                for (; bits; bits &= bits - 1) leafq[q++] = (uint32_t)(w * 32 + __ffs(bits) - 1);
                dleaf[w] = 0;
            }
            __syncthreads();
            const int n_lq = scal[6];
            for (int qi = wid; qi < n_lq; qi += NW) {
                const int l = (int)leafq[qi];
                const SpBest b = sp_leaf_maxima(sp, ch, l, n_slots, lane);
                if (lane == 0) { leaf[l] = b; atomicOr(&dsup[l >> 11], 1u << ((l >> 6) & 31)); }
            }
            __syncthreads();
            if (tid == 0) scal[6] = 0;
            uint32_t mine_bits = 0;
            for (int j = wid; j < 32; j += NW) mine_bits |= 1u << j;
            for (int w = 0; w < w_sup; w++) {
                for (uint32_t bits = dsup[w] & mine_bits; bits; bits &= bits - 1) {
                    const int su = w * 32 + __ffs(bits) - 1;
                    const SpBest b = sp_sup_maxima(sp, ch, su, lane);
                    if (lane == 0) sup[su] = b;
                }
            }
            __syncthreads();
            for (int w = tid; w < w_sup; w += NT) dsup[w] = 0;
            SpBest b; b.clear();
            for (int su = tid; su < ch.n_sup; su += NT) b.merge(sup[su]);
            b = sp_warp_reduce(b);
            if (lane == 0) red[wid] = b;
            __syncthreads();
            SpBest r; r.clear();
            if (lane < NW) r = red[lane];
            r = sp_warp_reduce(r);
            return r;
        };
        SpBest so = refresh();
        while (so.M >= 0) {
            if (so.M >= so.maxP) {
                // ------------------------------------------------ merge (a,b) into a
                const int a = (int)(so.kF >> 16), b = (int)(so.kF & 0xffffu);
                // m1: S = nodes with an edge to a or b; their old weights and the slots of those edges
                for (int side = 0; side < 2; side++) {
                    const int u = side ? b : a;
                    const long long lp = lptr[u]; const int ll = (int)llen[u];
                    for (int i = tid; i < ll; i += NT) {
                        const uint32_t s = pool[lp + i];
                        if (flag[s] & SPF_DEAD) continue;
                        const int t = sp_other(key[s], u);
                        if (t == a || t == b) { if (!side) scal[5] = (int32_t)s; continue; }      // the edge (a,b) itself
                        if (side) sb[t] = s; else sa[t] = s;
                        if (!(atomicOr(&inS[t >> 5], 1u << (t & 31)) & (1u << (t & 31)))) slist[atomicAdd(&scal[1], 1)] = (uint32_t)t;
                    }
                }
                __syncthreads();
                const int n_s = scal[1];
                for (int i = tid; i < n_s; i += NT) {
                    const int t = (int)slist[i];
                    const int xa = W[(int64_t)a * n + t], xb = W[(int64_t)b * n + t];
                    wa[t] = xa; wb[t] = xb; nw[t] = (xa == CC_FORB || xb == CC_FORB) ? CC_FORB : xa + xb;
                }
                for (int t = tid; t < n; t += NT) if (label[t] == b) label[t] = (uint16_t)a;
                __syncthreads();
                // m2: one warp per x in S walks list(x): fresh induced costs of (a,x) over the common neighbours in S, and the
                // changed terms of the edges (x,y) inside S (done from the smaller end)
                for (int xi = wid; xi < n_s; xi += NW) {
                    const int x = (int)slist[xi];
                    const int xa = wa[x], xb = wb[x], xn = nw[x];
                    const long long lp = lptr[x]; const int ll = (int)llen[x];
                    long long f = 0, p = 0;
                    // four entries per lane in flight: the loads of a step (slot index -> flag, key -> weights -> costs) are issued for
                    // all four before the first is used — the walk is a chain of dependent L2 accesses otherwise
                    constexpr int U = NT >= 1024 ? 4 : 2;      // the 256-thread blocks run five to an SM: registers matter more there
                    // The walk also compacts the list: the slots of a node's partners die one by one as the partners merge into clusters,
                    // and late in a long chain most entries of list(x) are dead — they are dropped here (live entries move down, in order)
                    int kept = 0;                                          // warp-uniform: live entries written back so far
                    for (int base = 0; base < ll; base += 32 * U) {
                        const int i0 = base + lane;
                        uint32_t s[U], ky[U]; uint8_t fl[U]; bool ok[U], in[U]; int y[U], w_xy[U], ya[U], yb[U], yn[U];
#pragma unroll
                        for (int u = 0; u < U; u++) { in[u] = i0 + 32 * u < ll; s[u] = in[u] ? pool[lp + i0 + 32 * u] : 0u; }
#pragma unroll
                        for (int u = 0; u < U; u++) { fl[u] = in[u] ? flag[s[u]] : SPF_DEAD; ky[u] = in[u] ? key[s[u]] : 0u; }
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            y[u] = sp_other(ky[u], x);
                            ok[u] = !(fl[u] & SPF_DEAD) && y[u] != a && y[u] != b;
                            ok[u] = ok[u] && ((inS[y[u] >> 5] >> (y[u] & 31)) & 1u);
                        }
#pragma unroll
                        for (int u = 0; u < U; u++) if (ok[u]) { w_xy[u] = W[(int64_t)x * n + y[u]]; ya[u] = wa[y[u]]; yb[u] = wb[y[u]]; yn[u] = nw[y[u]]; }
                        __syncwarp();                                      // every lane has read its entries of this stretch
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            const bool live = !(fl[u] & SPF_DEAD);
                            const unsigned bal = __ballot_sync(0xffffffffu, live);
                            const int to = kept + __popc(bal & ((1u << lane) - 1u));
                            if (live && to != i0 + 32 * u) pool[lp + to] = s[u];       // to <= the entry's own index: never ahead of a read
                            kept += __popc(bal);
                        }
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            if (!ok[u]) continue;
                            f += sp_tf(yn[u], w_xy[u]); p += sp_tp(yn[u], w_xy[u]);
                            if (x < y[u] && !(fl[u] & SPF_FORB)) {
                                const long long df = sp_tf(xn, yn[u]) - sp_tf(xa, ya[u]) - sp_tf(xb, yb[u]), dp = sp_tp(xn, yn[u]) - sp_tp(xa, ya[u]) - sp_tp(xb, yb[u]);
                                if (df != 0 || dp != 0) { F[s[u]] += df; P[s[u]] += dp; mark(s[u]); }
                            }
                        }
                    }
                    if (lane == 0 && kept != ll) llen[x] = (uint32_t)kept;
                    f = warp_sum_i64(f); p = warp_sum_i64(p);
                    if (lane == 0) { frF[x] = f + max(xn, 0); frP[x] = p + max(-xn, 0); }
                }
                __syncthreads();
                // m3: the edges at a and b: (a,x) continues with the merged weight, (b,x) takes over as (a,x) if a had none;
                // list(a) is written afresh; rows / columns a, b of W
                if (tid == 0) {
                    scal[4] = 0;
                    flag[scal[5]] = SPF_DEAD; mark((uint32_t)scal[5]);                           // the slot of (a,b) dies
                    W[(int64_t)a * n + b] = 0; W[(int64_t)b * n + a] = 0;
                }
                const long long newl = sp.bump[ci];
                __syncthreads();
                for (int i = tid; i < n_s; i += NT) {
                    const int x = (int)slist[i];
                    const int xn = nw[x];
                    const uint32_t s_a = sa[x], s_b = sb[x];
                    uint32_t keep = s_a != SP_NONE ? s_a : s_b;
                    if (s_a != SP_NONE && s_b != SP_NONE) { flag[s_b] = SPF_DEAD; mark(s_b); }
                    if (xn == 0) { flag[keep] = SPF_DEAD; mark(keep); keep = SP_NONE; }
                    else {
                        key[keep] = a < x ? ((uint32_t)a << 16) | (uint32_t)x : ((uint32_t)x << 16) | (uint32_t)a;
                        if (xn == CC_FORB) flag[keep] = SPF_FORB;
                        else { flag[keep] = xn > 0 ? SPF_POS : 0; F[keep] = frF[x]; P[keep] = frP[x]; }
                        mark(keep);
                    }
                    if (keep != SP_NONE && newl + n_s <= ch.list_cap) pool[newl + atomicAdd(&scal[4], 1)] = keep;
                    W[(int64_t)a * n + x] = xn; W[(int64_t)x * n + a] = xn; W[(int64_t)b * n + x] = 0; W[(int64_t)x * n + b] = 0;
                    sa[x] = SP_NONE; sb[x] = SP_NONE;
                    atomicAnd(&inS[x >> 5], ~(1u << (x & 31)));
                }
                __syncthreads();
                if (newl + n_s > ch.list_cap) {                   // list pool exhausted: give the chain up (reported, never mis-clustered)
                    if (tid == 0) d.ch_status[ch.chain] = AHS_CHAIN_TOO_LARGE;
                    so.M = -1;
                    continue;
                }
                if (tid == 0) { lptr[a] = newl; llen[a] = (uint32_t)scal[4]; llen[b] = 0; sp.bump[ci] = newl + scal[4]; scal[1] = 0; }
                n_active--;
                force_single = false;
                so = refresh();
            } else if (force_single || so.maxPpos > so.M) {
                // ------------------------------------------------ one sequential forbid: the edge with the largest icp
                const int a = (int)(so.kP >> 16), b = (int)(so.kP & 0xffffu);
                const int old = W[(int64_t)a * n + b];
                for (int side = 0; side < 2; side++) {
                    const int u = side ? b : a, third = side ? a : b;
                    const long long lp = lptr[u]; const int ll = (int)llen[u];
                    for (int i = tid; i < ll; i += NT) {
                        const uint32_t s = pool[lp + i];
                        const uint8_t fl = flag[s];
                        if (fl & SPF_DEAD) continue;
                        const int o = sp_other(key[s], u);
                        if (o == third) { if (!side) { flag[s] = SPF_FORB; mark(s); } continue; }      // the edge itself: forbidden from now on
                        if (fl & SPF_FORB) continue;
                        const int wt = W[(int64_t)o * n + third];
                        if (wt != 0) { F[s] -= sp_tf(old, wt); P[s] += sp_tp(CC_FORB, wt) - sp_tp(old, wt); mark(s); }
                    }
                }
                __syncthreads();
                if (tid == 0) { W[(int64_t)a * n + b] = CC_FORB; W[(int64_t)b * n + a] = CC_FORB; }
                force_single = false;
                so = refresh();
            } else {
                // ------------------------------------------------ round: all negative candidates with icp > M at once
                // (exactness argument: k_chain.cuh).  They are found from the top of the tree.
                const long long M = so.M;
                for (int su = tid; su < ch.n_sup; su += NT) if (sup[su].maxPneg > M) supq[atomicAdd(&scal[3], 1)] = (uint32_t)su;
                __syncthreads();
                const int n_q = scal[3];
                for (int idx = wid; idx < n_q * 64; idx += NW) {           // (level-2 entry, leaf) pairs dealt over the warps
                    const int l = (int)supq[idx >> 6] * 64 + (idx & 63);
                    if (l >= ch.n_leaf || leaf[l].maxPneg <= M) continue;      // warp-uniform
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int s = l * 64 + u * 32 + lane;
                        if (s >= n_slots) continue;
                        const uint8_t fl = flag[s];
                        if ((fl & (SPF_DEAD | SPF_FORB | SPF_POS)) || P[s] <= M) continue;
                        const int q = atomicAdd(&scal[2], 1);              // never truncated: a partial round would not be exact
                        flag[s] = fl | SPF_FLAG;
                        fl_slot[q] = (uint32_t)s; fl_old[q] = W[(int64_t)(key[s] >> 16) * n + (key[s] & 0xffffu)];
                    }
                }
                __syncthreads();
                const int nflag = scal[2];
                // growth of the icp of the edges at both ends of every flagged edge (x,y): the term of (x,t) through y grows
                // from min(|w_xy|, w_ty) to w_ty when w_ty > 0.  sign = -1 undoes it.
                auto grow = [&](long long sign) {
                    for (int e = wid; e < nflag; e += NW) {
                        const uint32_t fs = fl_slot[e]; const int old = fl_old[e];
                        const int x = (int)(key[fs] >> 16), y = (int)(key[fs] & 0xffffu);
                        for (int side = 0; side < 2; side++) {
                            const int u = side ? y : x, third = side ? x : y;
                            const long long lp = lptr[u]; const int ll = (int)llen[u];
                            for (int i = lane; i < ll; i += 32) {
                                const uint32_t s = pool[lp + i];
                                if (flag[s] & (SPF_DEAD | SPF_FORB | SPF_FLAG)) continue;
                                const int t = sp_other(key[s], u);
                                const int g = max(max(W[(int64_t)third * n + t], 0) + old, 0);
                                if (g > 0) { atomicAdd((unsigned long long*)&P[s], (unsigned long long)(sign * (long long)g)); mark(s); }
                            }
                        }
                    }
                };
                grow(1);
                for (int e = tid; e < nflag; e += NT) mark(fl_slot[e]);
                // the flagged edges are no candidates while the round is judged
                __syncthreads();
                for (int e = tid; e < nflag; e += NT) flag[fl_slot[e]] = SPF_FORB | SPF_FLAG;
                const SpBest v = refresh();
                const bool ok = nflag == 1 || v.maxPpos < 0 || v.M < 0 || v.maxPpos <= v.M;
                if (ok) {
                    for (int e = tid; e < nflag; e += NT) {
                        const uint32_t fs = fl_slot[e];
                        flag[fs] = SPF_FORB;
                        const int x = (int)(key[fs] >> 16), y = (int)(key[fs] & 0xffffu);
                        W[(int64_t)x * n + y] = CC_FORB; W[(int64_t)y * n + x] = CC_FORB;
                    }
                    so = v;
                    __syncthreads();
                } else {
                    // roll back, then one sequential step on the unchanged `so`
                    for (int e = tid; e < nflag; e += NT) flag[fl_slot[e]] = SPF_FLAG;      // negative candidates again, still skipped by grow
                    __syncthreads();
                    grow(-1);
                    __syncthreads();
                    for (int e = tid; e < nflag; e += NT) { flag[fl_slot[e]] = 0; mark(fl_slot[e]); }
                    force_single = true;
                    (void)refresh();                                   // the tree is that of `so` again
                }
                if (tid == 0) { scal[2] = 0; scal[3] = 0; }
                __syncthreads();
            }
        }
        // ---- clusters: numbered by smallest member (= representative), ascending
        __syncthreads();
        const int64_t f0 = d.frow_off[ch.chain];
        // active[x] <=> label[x] == x; rank of a representative = number of representatives below it (block scan over n)
        {
            int* cnt = (int*)red;                                    // 32 warp totals
            int running = 0;
            for (int x0 = 0; x0 < n; x0 += NT) {
                const int x = x0 + tid;
                const bool rep = x < n && label[x] == x;
                const unsigned bal = __ballot_sync(0xffffffffu, rep);
                if (lane == 0) cnt[wid] = __popc(bal);
                __syncthreads();
                int before = running;
                for (int w = 0; w < wid; w++) before += cnt[w];
                int total = 0; for (int w = 0; w < NW; w++) total += cnt[w];
                if (rep) nw[x] = before + __popc(bal & ((1u << lane) - 1u));      // cluster id of representative x
                running += total;
                __syncthreads();
            }
            for (int x = tid; x < n; x += NT) d.fr_cluster[f0 + x] = nw[label[x]];
            if (tid == 0) d.ch_nclusters[ch.chain] = running;
        }
        (void)n_active;
    }
}

}  // namespace ahs
