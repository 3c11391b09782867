// k_cluster_big.cuh — cluster editing (rule R2) of chains ABOVE CC_MAXN final reads, one block of 256 threads per
// chain, state in HBM / L2.  The predecessor of k_cluster_sparse.cuh, which takes those chains now; selected
// with AHS_CLUSTER_BIG=1 for comparison runs (chains of 161-8,191 reads) and kept under test.
//
// Replaces ClusterEditingSolver(sim,false).run() (call site reference src/alignmentstoreadset.cpp:312-315;
// algorithm: oracle/core/phase_core.hpp rule R2) where the shared-memory kernel of k_chain.cuh does not fit.  Same
// step structure as k_cluster_chain (slots that only die, values + tie-break keys folded into every pass, exact
// batching of forbid rounds), with
//   * the dense weight matrix W[n][n] (int32, written by k_pair_scores) and the growth matrix D[n][n] (int64) in HBM;
//   * the candidate pairs as a slot list (key, icf, icp) in HBM, walked with a block stride and compacted when half
//     of it is dead; icf / icp are int64 (sum of |w| over the pairs of a long chain exceeds 2^31);
//   * the sparsity of long chains used everywhere: reads only overlap their neighbours along the chain, so a merge
//     touches the nodes adjacent to a or b (list S), induced costs are initialised over the common column range of
//     two rows, and a forbid round walks only the rows of its end nodes.
// Shared memory holds the per-node state (active list, labels, membership bit sets, the list S, the flagged edges).
#pragma once
#include "common.cuh"
#include "device_batch.cuh"
#include "k_chain.cuh"

namespace ahs {

constexpr int CB_THREADS = 256;                      // several chains per SM: a greedy step is a chain of memory latencies that only other chains hide
constexpr int CB_FLCAP = 2048;                      // flagged edges per forbid round (more -> the round is abandoned for a sequential step)
constexpr uint32_t CB_POS = 1u << 31, CB_FLAG = 1u << 30, CB_DEAD = 1u << 29, CB_GONE = CB_FLAG | CB_DEAD, CB_KEY = (1u << 26) - 1u;

__host__ __device__ inline size_t cb_smem_bytes(int nmax) {
    const size_t words = (size_t)(nmax + 31) / 32;
    size_t b = (size_t)nmax * 2 * 4;                // alist, apos, label, slist (u16 each)
    b += (size_t)nmax;                              // active
    b += words * 4 * 2;                             // inS, nodefl bit sets
    b += (size_t)CB_FLCAP * 8;                      // flagged edges: (a << 16 | b), old weight
    b += 32 * 48 + 256;                             // reduction scratch, scalars
    return (b + 15) & ~(size_t)15;
}

struct CBBest {
    long long M, maxP, maxPpos; uint32_t kF, kP; int live;
    __device__ __forceinline__ void clear() { M = -1; maxP = -1; maxPpos = -1; kF = CB_KEY; kP = CB_KEY; live = 0; }
    __device__ __forceinline__ void consider(uint32_t key, long long f, long long p) {
        const uint32_t kq = key & CB_KEY;
        if (f > M || (f == M && kq < kF)) { M = f; kF = kq; }
        if (p > maxP || (p == maxP && kq < kP)) { maxP = p; kP = kq; }
        if ((key & CB_POS) && p > maxPpos) maxPpos = p;
        live++;
    }
    __device__ __forceinline__ void merge(const CBBest& o) {
        if (o.M > M || (o.M == M && o.kF < kF)) { M = o.M; kF = o.kF; }
        if (o.maxP > maxP || (o.maxP == maxP && o.kP < kP)) { maxP = o.maxP; kP = o.kP; }
        if (o.maxPpos > maxPpos) maxPpos = o.maxPpos;
        live += o.live;
    }
};

__device__ __forceinline__ CBBest cb_shfl_xor(const CBBest& b, int o) {
    CBBest r;
    r.M = __shfl_xor_sync(0xffffffffu, b.M, o); r.maxP = __shfl_xor_sync(0xffffffffu, b.maxP, o); r.maxPpos = __shfl_xor_sync(0xffffffffu, b.maxPpos, o);
    r.kF = __shfl_xor_sync(0xffffffffu, b.kF, o); r.kP = __shfl_xor_sync(0xffffffffu, b.kP, o); r.live = __shfl_xor_sync(0xffffffffu, b.live, o);
    return r;
}

// block-wide combination; two barriers
template <int NW>
__device__ __forceinline__ CBBest cb_reduce(CBBest b, CBBest* red, int tid) {
    for (int o = 16; o > 0; o >>= 1) { const CBBest t = cb_shfl_xor(b, o); b.merge(t); }
    __syncthreads();                                                   // red is free again
    if ((tid & 31) == 0) red[tid >> 5] = b;
    __syncthreads();
    CBBest r; r.clear();
    if ((tid & 31) < NW) r = red[tid & 31];
    for (int o = 16; o > 0; o >>= 1) { const CBBest t = cb_shfl_xor(r, o); r.merge(t); }
    return r;
}

__device__ __forceinline__ long long cb_tf(int x, int y) { return (long long)max(min(x, y), 0); }
__device__ __forceinline__ long long cb_tp(int x, int y) { const int lo = min(x, y), hi = max(x, y); return (long long)max(min(hi, -lo), 0); }

__global__ void __launch_bounds__(CB_THREADS) k_cluster_big(DB d, const int32_t* __restrict__ chains, int n_list, int nmax,
                                                            int32_t* __restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char cb_sm[];
    constexpr int NT = CB_THREADS, NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int words = (nmax + 31) / 32;
    CBBest* red; int32_t* scal; uint32_t *inS, *nodefl, *fl_ab; int32_t* fl_old; uint16_t *alist, *apos, *label, *slist; uint8_t* active;
    {
        unsigned char* p = cb_sm;
        red = (CBBest*)p; p += 32 * 48;
        scal = (int32_t*)p; p += 256;               // [0] active nodes [1] slots in the list [2] |S| [3] work item [4] flagged edges [5] scratch counter
        fl_ab = (uint32_t*)p; p += CB_FLCAP * 4; fl_old = (int32_t*)p; p += CB_FLCAP * 4;
        inS = (uint32_t*)p; p += words * 4; nodefl = (uint32_t*)p; p += words * 4;
        alist = (uint16_t*)p; p += nmax * 2; apos = (uint16_t*)p; p += nmax * 2; label = (uint16_t*)p; p += nmax * 2; slist = (uint16_t*)p; p += nmax * 2;
        active = p; p += nmax;
    }
    static_assert(sizeof(CBBest) <= 48, "reduction scratch");
    while (true) {
        __syncthreads();
        if (tid == 0) scal[3] = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = scal[3];
        if (item >= n_list) break;
        const int c = chains[item];
        const int64_t f0 = d.frow_off[c];
        const int n = (int)(d.frow_off[c + 1] - f0);
        const int64_t nn = (int64_t)n * n;
        int32_t* W = d.W + d.cw_off[c];
        long long* SF = (long long*)(d.F + d.cf_off[c]); long long* SP = SF + nn / 2;            // slot icf / icp (<= n(n-1)/2 slots)
        long long* Dm = (long long*)(d.P + d.cf_off[c]);                                         // growth matrix of the rounds / packing area
        uint32_t* SK = d.big_key + d.cf_off[c];                                                  // slot keys
        int32_t* wa = d.ce_list + f0; int32_t* wb = d.ce_newrow + f0; int32_t* nwv = d.ce_label + f0;
        long long* frF = (long long*)(d.ce_rbF + f0); long long* frP = (long long*)(d.ce_rbP + f0);
        int32_t* lo = d.ce_rbFarg + f0; int32_t* hi = d.ce_rbParg + f0;
        for (int x = tid; x < n; x += NT) { active[x] = 1; label[x] = (uint16_t)x; alist[x] = (uint16_t)x; apos[x] = (uint16_t)x; lo[x] = n; hi[x] = -1; }
        for (int x = tid; x < words; x += NT) { inS[x] = 0; nodefl[x] = 0; }
        if (tid == 0) { scal[0] = n; scal[1] = 0; scal[2] = 0; scal[4] = 0; scal[5] = 0; }
        __syncthreads();
        // ---- slot list: the non-zero pairs of the upper triangle; column range of every row
        for (int64_t i0 = 0; i0 < nn; i0 += NT) {                  // uniform trip count: the ballots need every lane
            const int64_t idx = i0 + tid;
            const int x = (int)(idx / n), y = (int)(idx - (int64_t)x * n);
            const int w = (idx < nn && y > x) ? W[idx] : 0;
            const unsigned bal = __ballot_sync(0xffffffffu, w != 0);
            if (bal) {
                int base = 0;
                if (lane == __ffs(bal) - 1) base = atomicAdd(&scal[1], __popc(bal));
                base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
                if (w != 0) {
                    SK[base + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)((x << 13) | y) | (w > 0 ? CB_POS : 0u);
                    atomicMin(&lo[x], y); atomicMax(&hi[x], y); atomicMin(&lo[y], x); atomicMax(&hi[y], x);
                }
            }
        }
        __syncthreads();
        int E = scal[1];
        CBBest mine; mine.clear();
        // ---- initial induced costs over the common column range of the two rows (W[x][x] = 0)
        for (int s = tid; s < E; s += NT) {
            const uint32_t kq = SK[s];
            const int x = (int)((kq & CB_KEY) >> 13), y = (int)(kq & 0x1fffu);
            const int32_t* rx = W + (int64_t)x * n; const int32_t* ry = W + (int64_t)y * n;
            const int w = rx[y];
            long long f = max(w, 0), p = max(-w, 0);
            const int t0 = max(lo[x], lo[y]), t1 = min(hi[x], hi[y]);
            for (int t = t0; t <= t1; t++) { const int wx = rx[t], wy = ry[t]; f += cb_tf(wx, wy); p += cb_tp(wx, wy); }
            SF[s] = f; SP[s] = p;
            mine.consider(kq, f, p);
        }
        CBBest so = cb_reduce<NW>(mine, red, tid);
        int live_cap = so.live;
        bool force_single = false;
        while (so.M >= 0) {
            mine.clear();
            if (so.M >= so.maxP) {
                // ------------------------------------------------ merge (a,b) into a
                const int a = (int)(so.kF >> 13), b = (int)(so.kF & 0x1fffu);
                for (int x = tid; x < words; x += NT) inS[x] = 0;
                if (tid == 0) scal[2] = 0;
                __syncthreads();
                for (int t = tid; t < n; t += NT) {
                    int xa = W[(int64_t)a * n + t], xb = W[(int64_t)b * n + t];
                    if (t == a || t == b) { xa = 0; xb = 0; }
                    const int nwn = (xa == CC_FORB || xb == CC_FORB) ? CC_FORB : xa + xb;
                    wa[t] = xa; wb[t] = xb; nwv[t] = nwn;
                    if (xa != 0 || xb != 0) {                      // S: the nodes adjacent to a or b
                        W[(int64_t)a * n + t] = nwn; W[(int64_t)t * n + a] = nwn; W[(int64_t)b * n + t] = 0; W[(int64_t)t * n + b] = 0;
                        atomicOr(&inS[t >> 5], 1u << (t & 31));
                        slist[atomicAdd(&scal[2], 1)] = (uint16_t)t;
                    }
                    if (label[t] == b) label[t] = (uint16_t)a;
                    if (t == b) {
                        active[b] = 0;
                        const int nact = scal[0], pos = apos[b], lastn = alist[nact - 1];
                        alist[pos] = (uint16_t)lastn; apos[lastn] = (uint16_t)pos; scal[0] = nact - 1;
                        W[(int64_t)a * n + b] = 0; W[(int64_t)b * n + a] = 0;
                    }
                }
                __syncthreads();
                // fresh induced costs of the pairs (a,x), x in S: third nodes are the members of S
                {
                    const int ns_ = scal[2];
                    for (int xi = wid; xi < ns_; xi += NW) {
                        const int x = slist[xi];
                        const int w = nwv[x];
                        if (w == 0 || w == CC_FORB) continue;
                        const int32_t* rx = W + (int64_t)x * n;
                        long long f = 0, p = 0;
                        for (int vi = lane; vi < ns_; vi += 32) { const int v = slist[vi]; const int t1 = nwv[v], t2 = rx[v]; f += cb_tf(t1, t2); p += cb_tp(t1, t2); }
                        f = warp_sum_i64(f); p = warp_sum_i64(p);
                        if (lane == 0) { frF[x] = f + max(w, 0); frP[x] = p + max(-w, 0); }
                    }
                }
                __syncthreads();
                for (int s = tid; s < E; s += NT) {
                    uint32_t kq = SK[s];
                    if (kq & CB_GONE) continue;
                    const int x = (int)((kq & CB_KEY) >> 13), y = (int)(kq & 0x1fffu);
                    long long f = SF[s], p = SP[s];
                    if (x == a || y == a || x == b || y == b) {
                        const bool thru_b = x == b || y == b;
                        const int o = thru_b ? (x == b ? y : x) : (x == a ? y : x);
                        const int nwo = (o == a || o == b) ? 0 : nwv[o], wao = (o == a || o == b) ? 0 : wa[o];
                        if (nwo == 0 || nwo == CC_FORB || (thru_b && wao != 0)) { SK[s] = CB_DEAD; continue; }
                        kq = (uint32_t)(a < o ? (a << 13) | o : (o << 13) | a) | (nwo > 0 ? CB_POS : 0u);
                        f = frF[o]; p = frP[o];
                        SK[s] = kq; SF[s] = f; SP[s] = p;
                    } else if (((inS[x >> 5] >> (x & 31)) & (inS[y >> 5] >> (y & 31)) & 1u) != 0) {
                        const int xa = wa[x], xb = wb[x], xn = nwv[x], ya = wa[y], yb = wb[y], yn = nwv[y];
                        f += cb_tf(xn, yn) - cb_tf(xa, ya) - cb_tf(xb, yb);
                        p += cb_tp(xn, yn) - cb_tp(xa, ya) - cb_tp(xb, yb);
                        SF[s] = f; SP[s] = p;
                    }
                    mine.consider(kq, f, p);
                }
                force_single = false;
                so = cb_reduce<NW>(mine, red, tid);
                // pack the slot list once half of it is dead (through the growth matrix, free outside a round)
                if (so.live * 2 <= live_cap && so.live >= NT) {
                    unsigned char* pk = (unsigned char*)Dm;                      // icf | icp | keys of the survivors: 20 B each, <= 5 n^2 B
                    const int64_t L = so.live;
                    for (int s0 = 0; s0 < E; s0 += NT) {
                        const int s = s0 + tid;
                        uint32_t kq = CB_DEAD; if (s < E) kq = SK[s];
                        const bool lv = !(kq & CB_GONE);
                        const unsigned bal = __ballot_sync(0xffffffffu, lv);
                        int base = 0;
                        if (bal && lane == __ffs(bal) - 1) base = atomicAdd(&scal[5], __popc(bal));
                        if (bal) base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
                        if (lv) {
                            const int64_t pos = base + __popc(bal & ((1u << lane) - 1u));
                            *(long long*)(pk + pos * 8) = SF[s]; *(long long*)(pk + L * 8 + pos * 8) = SP[s];
                            *(uint32_t*)(pk + L * 16 + pos * 4) = kq;
                        }
                    }
                    __syncthreads();
                    const int total = scal[5];
                    for (int s = tid; s < total; s += NT) {
                        SF[s] = *(long long*)(pk + (int64_t)s * 8); SP[s] = *(long long*)(pk + L * 8 + (int64_t)s * 8);
                        SK[s] = *(uint32_t*)(pk + L * 16 + (int64_t)s * 4);
                    }
                    E = total; live_cap = total;
                    __syncthreads();
                    if (tid == 0) scal[5] = 0;
                }
            } else if (force_single || so.maxPpos > so.M) {
                // ------------------------------------------------ one sequential forbid: the edge with the largest icp
                const int a = (int)(so.kP >> 13), b = (int)(so.kP & 0x1fffu);
                const int old = W[(int64_t)a * n + b];
                for (int s = tid; s < E; s += NT) {
                    const uint32_t kq = SK[s];
                    if (kq & CB_GONE) continue;
                    if ((kq & CB_KEY) == so.kP) { SK[s] = CB_DEAD; continue; }
                    const int x = (int)((kq & CB_KEY) >> 13), y = (int)(kq & 0x1fffu);
                    long long f = SF[s], p = SP[s];
                    int o = -1, third = 0;
                    if (x == a || y == a) { o = x == a ? y : x; third = b; }
                    else if (x == b || y == b) { o = x == b ? y : x; third = a; }
                    if (o >= 0) {
                        const int wt = W[(int64_t)o * n + third];
                        if (wt != 0) { f -= cb_tf(old, wt); p += cb_tp(CC_FORB, wt) - cb_tp(old, wt); SF[s] = f; SP[s] = p; }
                    }
                    mine.consider(kq, f, p);
                }
                force_single = false;
                so = cb_reduce<NW>(mine, red, tid);
                if (tid == 0) { W[(int64_t)a * n + b] = CC_FORB; W[(int64_t)b * n + a] = CC_FORB; }
                __syncthreads();
            } else {
                // ------------------------------------------------ round: all negative candidates with icp > M at once
                // (exactness argument: k_chain.cuh)
                const long long M = so.M;
                for (int s = tid; s < E; s += NT) {
                    const uint32_t kq = SK[s];
                    if (kq & CB_FLAG) { SK[s] = CB_DEAD; continue; }               // forbidden for good in an earlier round
                    if ((kq & (CB_DEAD | CB_POS)) || SP[s] <= M) continue;
                    const int pos = atomicAdd(&scal[4], 1);
                    if (pos >= CB_FLCAP) { SK[s] = kq | CB_FLAG; continue; }       // buffer full: the round is abandoned below
                    const int x = (int)((kq & CB_KEY) >> 13), y = (int)(kq & 0x1fffu);
                    SK[s] = kq | CB_FLAG;
                    fl_ab[pos] = (uint32_t)((x << 16) | y); fl_old[pos] = W[(int64_t)x * n + y];
                    atomicOr(&nodefl[x >> 5], 1u << (x & 31)); atomicOr(&nodefl[y >> 5], 1u << (y & 31));
                }
                __syncthreads();
                if (scal[4] > CB_FLCAP) {
                    // more eligible edges than the buffer holds: a partial round is not exact (a positive edge could become the
                    // maximum between two parts), so nothing is forbidden here and one sequential step is taken instead
                    for (int s = tid; s < E; s += NT) if (SK[s] & CB_FLAG) SK[s] &= ~CB_FLAG;
                    for (int x = tid; x < words; x += NT) nodefl[x] = 0;
                    __syncthreads();
                    if (tid == 0) scal[4] = 0;
                    force_single = true;
                    __syncthreads();
                    continue;
                }
                const int nflag = scal[4];
                // zero the growth rows of the end nodes, then add the growth through every forbidden edge
                for (int xw = 0; xw < words; xw++)
                    for (uint32_t bits = nodefl[xw]; bits; bits &= bits - 1) {
                        const int x = xw * 32 + __ffs(bits) - 1;
                        for (int t = tid; t < n; t += NT) Dm[(int64_t)x * n + t] = 0;
                    }
                __syncthreads();
                for (int e = wid; e < nflag; e += NW) {
                    const int x = (int)(fl_ab[e] >> 16), y = (int)(fl_ab[e] & 0xffffu), old = fl_old[e];
                    const int32_t* rx = W + (int64_t)x * n; const int32_t* ry = W + (int64_t)y * n;
                    for (int t = lane; t < n; t += 32) {
                        // pair (x,t), third node y: grows by max(w_t,y - |old|, 0) when w_t,y > 0; same for (y,t) through x
                        const int gy = max(max(ry[t], 0) + old, 0), gx = max(max(rx[t], 0) + old, 0);
                        if (gy > 0) atomicAdd((unsigned long long*)&Dm[(int64_t)x * n + t], (unsigned long long)gy);
                        if (gx > 0) atomicAdd((unsigned long long*)&Dm[(int64_t)y * n + t], (unsigned long long)gx);
                    }
                }
                __syncthreads();
                for (int s = tid; s < E; s += NT) {
                    const uint32_t kq = SK[s];
                    if (kq & CB_DEAD) continue;
                    const int x = (int)((kq & CB_KEY) >> 13), y = (int)(kq & 0x1fffu);
                    if (kq & CB_FLAG) { W[(int64_t)x * n + y] = CC_FORB; W[(int64_t)y * n + x] = CC_FORB; continue; }
                    long long p = SP[s];
                    const bool fx = (nodefl[x >> 5] >> (x & 31)) & 1u, fy = (nodefl[y >> 5] >> (y & 31)) & 1u;
                    if (fx | fy) {
                        if (fx) p += Dm[(int64_t)x * n + y];
                        if (fy) p += Dm[(int64_t)y * n + x];
                        SP[s] = p;
                    }
                    mine.consider(kq, SF[s], p);
                }
                const CBBest v = cb_reduce<NW>(mine, red, tid);
                const bool ok = nflag == 1 || v.maxPpos < 0 || v.M < 0 || v.maxPpos <= v.M;
                if (ok) so = v;
                else {
                    force_single = true;                       // roll back, then one sequential step on the unchanged `so`
                    for (int s = tid; s < E; s += NT) {
                        const uint32_t kq = SK[s];
                        if (kq & CB_DEAD) continue;
                        const int x = (int)((kq & CB_KEY) >> 13), y = (int)(kq & 0x1fffu);
                        if (kq & CB_FLAG) { SK[s] = kq & ~CB_FLAG; continue; }
                        const bool fx = (nodefl[x >> 5] >> (x & 31)) & 1u, fy = (nodefl[y >> 5] >> (y & 31)) & 1u;
                        if (fx | fy) {
                            long long p = SP[s];
                            if (fx) p -= Dm[(int64_t)x * n + y];
                            if (fy) p -= Dm[(int64_t)y * n + x];
                            SP[s] = p;
                        }
                    }
                    for (int e = tid; e < nflag; e += NT) {
                        const int x = (int)(fl_ab[e] >> 16), y = (int)(fl_ab[e] & 0xffffu);
                        W[(int64_t)x * n + y] = fl_old[e]; W[(int64_t)y * n + x] = fl_old[e];
                    }
                }
                __syncthreads();
                for (int x = tid; x < words; x += NT) nodefl[x] = 0;
                if (tid == 0) scal[4] = 0;
                __syncthreads();
            }
        }
        // ---- clusters: numbered by smallest member (= representative), ascending
        __syncthreads();
        for (int x = tid; x < n; x += NT) {
            const int rep = label[x];
            int cid = 0;
            for (int y = 0; y < rep; y++) cid += active[y];
            d.fr_cluster[f0 + x] = cid;
        }
        if (tid == 0) d.ch_nclusters[c] = scal[0];
    }
}

}  // namespace ahs
