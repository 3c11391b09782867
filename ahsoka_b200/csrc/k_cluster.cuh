// k_cluster.cuh — cluster editing of the read-similarity graph, one thread block per chain.
//
// Replaces ClusterEditingSolver(sim,false).run() (call site reference
// src/alignmentstoreadset.cpp:312-315; algorithm: oracle/core/phase_core.hpp rule R2 — the
// induced-cost greedy heuristic with node merging).  All quantities are integers (Q10 weights,
// 64-bit induced costs), so the incremental updates below are exact and the result does not
// depend on the order in which threads apply them.
//
// Per chain with n final reads the workspace is three dense n x n matrices in HBM/L2:
//   W  int32  symmetric weights (0 = no edge, FORB = forbidden), written by k_pair_scores
//   F  int64  icf(a,b) for a<b      P  int64  icp(a,b) for a<b
// plus per-row caches of the best candidate (rbF/rbP) that are rescanned only when a row is
// marked dirty.  The greedy step itself is sequential (argmax -> merge or forbid); the work
// inside a step (neighbour lists, pair deltas, row rescans) is spread over the block.
#pragma once
#include "common.cuh"
#include "device_batch.cuh"

namespace ahs {

__device__ __forceinline__ int64_t ce_tf(int32_t x, int32_t y) { return (x > 0 && y > 0) ? (int64_t)min(x, y) : 0; }
__device__ __forceinline__ int64_t ce_abs(int32_t x) { return x == FORB ? INF64 : (x < 0 ? -(int64_t)x : (int64_t)x); }
__device__ __forceinline__ int64_t ce_tp(int32_t x, int32_t y) {
    if (x > 0 && y < 0) return min((int64_t)x, ce_abs(y));
    if (x < 0 && y > 0) return min(ce_abs(x), (int64_t)y);
    return 0;
}
__device__ __forceinline__ bool ce_better(int64_t v1, int a1, int64_t v2, int a2) { return v1 > v2 || (v1 == v2 && a1 < a2); }

constexpr int CE_THREADS = 256;

__global__ void __launch_bounds__(CE_THREADS) k_cluster_edit(DB d, const int32_t* __restrict__ chains, int n_list, int32_t* __restrict__ work_counter) {
    __shared__ int64_t s_val[CE_THREADS];
    __shared__ int32_t s_arg[CE_THREADS];
    __shared__ int64_t s_bestF, s_bestP;
    __shared__ int32_t s_aF, s_bF, s_aP, s_bP, s_cnt, s_chain;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    while (true) {
        if (tid == 0) s_chain = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_chain;
        __syncthreads();
        if (item >= n_list) break;
        const int c = chains[item];
        const int64_t f0 = d.frow_off[c];
        const int n = (int)(d.frow_off[c + 1] - f0);
        if (n == 0) continue;
        int32_t* W = d.W + d.cw_off[c];
        int64_t* F = d.F + d.cw_off[c];
        int64_t* P = d.P + d.cw_off[c];
        uint8_t* active = d.ce_active + f0; uint8_t* dirty = d.ce_dirty + f0;
        int32_t* list = d.ce_list + f0; int32_t* newrow = d.ce_newrow + f0; int32_t* label = d.ce_label + f0;
        int64_t* rbF = d.ce_rbF + f0; int64_t* rbP = d.ce_rbP + f0; int32_t* rbFarg = d.ce_rbFarg + f0; int32_t* rbParg = d.ce_rbParg + f0;
        auto Wat = [&](int a, int b) -> int32_t& { return W[(int64_t)a * n + b]; };

        for (int x = tid; x < n; x += nt) { active[x] = 1; dirty[x] = 1; label[x] = x; }
        __syncthreads();
        // initial induced costs: one warp per row a, lanes over b, third node loop
        for (int a = warp; a < n; a += nwarps)
            for (int b = a + 1 + lane; b < n; b += 32) {
                const int32_t w = Wat(a, b);
                if (w == 0) continue;
                int64_t f = w > 0 ? w : 0, p = w < 0 ? -(int64_t)w : 0;
                for (int t = 0; t < n; t++) if (t != a && t != b) {
                    const int32_t x = Wat(a, t); if (x == 0) continue;
                    const int32_t y = Wat(b, t);
                    f += ce_tf(x, y); p += ce_tp(x, y);
                }
                F[(int64_t)a * n + b] = f; P[(int64_t)a * n + b] = p;
            }
        __syncthreads();

        while (true) {
            // ---- rescan dirty rows (one warp per row)
            for (int a = warp; a < n; a += nwarps) {
                if (!dirty[a]) continue;
                int64_t bf = -1, bp = -1; int af = INT32_MAX, ap = INT32_MAX;
                if (active[a]) for (int b = a + 1 + lane; b < n; b += 32) {
                    if (!active[b]) continue;
                    const int32_t w = Wat(a, b);
                    if (w == 0 || w == FORB) continue;
                    const int64_t f = F[(int64_t)a * n + b], p = P[(int64_t)a * n + b];
                    if (ce_better(f, b, bf, af)) { bf = f; af = b; }
                    if (ce_better(p, b, bp, ap)) { bp = p; ap = b; }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    int64_t of = __shfl_xor_sync(0xffffffffu, bf, o); int oa = __shfl_xor_sync(0xffffffffu, af, o);
                    if (ce_better(of, oa, bf, af)) { bf = of; af = oa; }
                    int64_t op = __shfl_xor_sync(0xffffffffu, bp, o); int ob = __shfl_xor_sync(0xffffffffu, ap, o);
                    if (ce_better(op, ob, bp, ap)) { bp = op; ap = ob; }
                }
                if (lane == 0) { rbF[a] = bf; rbFarg[a] = af; rbP[a] = bp; rbParg[a] = ap; dirty[a] = 0; }
            }
            __syncthreads();
            // ---- block argmax over rows: icf
            {
                int64_t bv = -1; int ba = INT32_MAX;
                for (int a = tid; a < n; a += nt) { const int64_t v = rbF[a]; if (ce_better(v, a, bv, ba)) { bv = v; ba = a; } }
                s_val[tid] = bv; s_arg[tid] = ba;
                __syncthreads();
                for (int o = nt >> 1; o > 0; o >>= 1) {
                    if (tid < o && ce_better(s_val[tid + o], s_arg[tid + o], s_val[tid], s_arg[tid])) { s_val[tid] = s_val[tid + o]; s_arg[tid] = s_arg[tid + o]; }
                    __syncthreads();
                }
                if (tid == 0) { s_bestF = s_val[0]; s_aF = s_arg[0]; s_bF = s_val[0] >= 0 ? rbFarg[s_arg[0]] : -1; }
                __syncthreads();
                bv = -1; ba = INT32_MAX;
                for (int a = tid; a < n; a += nt) { const int64_t v = rbP[a]; if (ce_better(v, a, bv, ba)) { bv = v; ba = a; } }
                s_val[tid] = bv; s_arg[tid] = ba;
                __syncthreads();
                for (int o = nt >> 1; o > 0; o >>= 1) {
                    if (tid < o && ce_better(s_val[tid + o], s_arg[tid + o], s_val[tid], s_arg[tid])) { s_val[tid] = s_val[tid + o]; s_arg[tid] = s_arg[tid + o]; }
                    __syncthreads();
                }
                if (tid == 0) { s_bestP = s_val[0]; s_aP = s_arg[0]; s_bP = s_val[0] >= 0 ? rbParg[s_arg[0]] : -1; s_cnt = 0; }
                __syncthreads();
            }
            if (s_bestF < 0) break;                         // no candidate left
            if (s_bestF >= s_bestP) {
                // ================= merge (a,b) into a, a < b
                const int a = s_aF, b = s_bF;
                for (int x = tid; x < n; x += nt) {
                    if (!active[x] || x == a || x == b) continue;
                    const int32_t wa = Wat(a, x), wb = Wat(b, x);
                    if (wa == 0 && wb == 0) continue;
                    const int pos = atomicAdd(&s_cnt, 1);
                    list[pos] = x;
                    newrow[pos] = (wa == FORB || wb == FORB) ? FORB : wa + wb;
                }
                __syncthreads();
                const int cnt = s_cnt;
                // pairs inside S: the terms through a and b are replaced by the term through the merged node
                for (int64_t idx = tid; idx < (int64_t)cnt * cnt; idx += nt) {
                    const int u = (int)(idx / cnt), v = (int)(idx % cnt);
                    const int x = list[u], y = list[v];
                    if (x >= y) continue;
                    const int32_t w = Wat(x, y);
                    if (w == 0 || w == FORB) continue;
                    const int32_t xa = Wat(x, a), ya = Wat(y, a), xb = Wat(x, b), yb = Wat(y, b);
                    const int64_t df = ce_tf(newrow[u], newrow[v]) - ce_tf(xa, ya) - ce_tf(xb, yb);
                    const int64_t dp = ce_tp(newrow[u], newrow[v]) - ce_tp(xa, ya) - ce_tp(xb, yb);
                    if (df != 0 || dp != 0) { F[(int64_t)x * n + y] += df; P[(int64_t)x * n + y] += dp; dirty[x] = 1; }
                }
                __syncthreads();
                for (int u = tid; u < cnt; u += nt) {
                    const int x = list[u];
                    Wat(a, x) = newrow[u]; Wat(x, a) = newrow[u]; Wat(b, x) = 0; Wat(x, b) = 0;
                    dirty[x] = 1;
                }
                for (int x = tid; x < n; x += nt) if (label[x] == b) label[x] = a;
                if (tid == 0) { Wat(a, b) = 0; Wat(b, a) = 0; active[b] = 0; dirty[a] = 1; dirty[b] = 1; }
                __syncthreads();
                // fresh induced costs for the pairs (a,x): third nodes are exactly the members of S
                for (int u = warp; u < cnt; u += nwarps) {
                    const int x = list[u];
                    const int32_t w = newrow[u];
                    if (w == 0 || w == FORB) continue;
                    int64_t f = 0, p = 0;
                    for (int v = lane; v < cnt; v += 32) if (v != u) {
                        const int32_t t1 = newrow[v], t2 = Wat(x, list[v]);
                        f += ce_tf(t1, t2); p += ce_tp(t1, t2);
                    }
                    f = warp_sum_i64(f); p = warp_sum_i64(p);
                    if (lane == 0) {
                        f += w > 0 ? w : 0; p += w < 0 ? -(int64_t)w : 0;
                        const int lo = min(a, x), hi = max(a, x);
                        F[(int64_t)lo * n + hi] = f; P[(int64_t)lo * n + hi] = p; dirty[lo] = 1;
                    }
                }
                __syncthreads();
            } else {
                // ================= forbid (a,b)
                const int a = s_aP, b = s_bP;
                const int32_t old = Wat(a, b);
                for (int t = tid; t < n; t += nt) {
                    if (!active[t] || t == a || t == b) continue;
                    const int32_t tb = Wat(t, b), ta = Wat(t, a);
                    if (tb != 0 && ta != 0 && ta != FORB) {           // pair (a,t), third node b
                        const int64_t df = -ce_tf(old, tb), dp = ce_tp(FORB, tb) - ce_tp(old, tb);
                        if (df != 0 || dp != 0) { const int lo = min(a, t), hi = max(a, t); F[(int64_t)lo * n + hi] += df; P[(int64_t)lo * n + hi] += dp; dirty[lo] = 1; }
                    }
                    if (ta != 0 && tb != 0 && tb != FORB) {           // pair (b,t), third node a
                        const int64_t df = -ce_tf(old, ta), dp = ce_tp(FORB, ta) - ce_tp(old, ta);
                        if (df != 0 || dp != 0) { const int lo = min(b, t), hi = max(b, t); F[(int64_t)lo * n + hi] += df; P[(int64_t)lo * n + hi] += dp; dirty[lo] = 1; }
                    }
                }
                __syncthreads();
                if (tid == 0) { Wat(a, b) = FORB; Wat(b, a) = FORB; dirty[a] = 1; }
                __syncthreads();
            }
        }
        // ---- clusters: numbered by smallest member (= representative), ascending
        for (int x = tid; x < n; x += nt) {
            const int rep = label[x];
            int cid = 0;
            for (int y = 0; y < rep; y++) cid += active[y];
            d.fr_cluster[f0 + x] = cid;
        }
        if (tid == 0) { int k = 0; for (int y = 0; y < n; y++) k += active[y]; d.ch_nclusters[c] = k; }
        __syncthreads();
    }
}

}  // namespace ahs

// =====================================================================================
// int32 helpers for chains with n < 128 final reads (the common case: BASELINE configs average
// ~50 reads per chain).  With n < 128 and |w| <= 2^17 every induced cost is below
// n^2 * 2^17 < 2^31, so int32 sums are exact.
// =====================================================================================
namespace ahs {

__device__ __forceinline__ int cs_tf(int x, int y) { return (x > 0 && y > 0) ? min(x, y) : 0; }
__device__ __forceinline__ int cs_tp(int x, int y) {
    if (x > 0 && y < 0) return y == FORB ? x : min(x, -y);
    if (x < 0 && y > 0) return x == FORB ? y : min(-x, y);
    return 0;
}
}  // namespace ahs

// =====================================================================================
// Warp-per-chain variant (n < 128) — the production path for the BASELINE configs.
//
// The greedy loop of rule R2 is ~55 merges and ~420 single-edge forbids for a 55-read chain,
// and the forbids come in long runs.  Within a run the steps can be BATCHED EXACTLY:
//   * forbidding an edge of negative weight changes no icf (tf(x<0, .) = 0) and can only
//     INCREASE icp values (tp(x<0,y>0) = min(|x|,y) grows as |x| -> inf); removing it from the
//     candidates can only lower M = max icf;
//   * so every negative candidate whose icp exceeds M stays eligible until it is forbidden, and
//     the set forbidden before the next merge is the least fixed point of
//     "forbid all negative candidates with icp > M" — independent of the order;
//   * the argument breaks only if a POSITIVE edge would be picked inside the run.  A round is
//     therefore applied tentatively (icp deltas only, weights untouched), validated
//     (max icp over positive candidates <= new M) and, in the rare failing case, undone and
//     replaced by one sequential step.
// Results are bit-identical to the sequential definition (checked against the oracle, which
// stays strictly sequential).
//
// One warp owns one chain: no block barriers, only __syncwarp.  W / icf / icp are upper
// triangles in shared memory (3 x n(n-1)/2 int32) plus a compact list of candidate pairs, so a
// scan touches only live candidates.  ~9 chains per SM for n ~ 56.
// =====================================================================================
namespace ahs {

constexpr int CW_WARPS = 4;
constexpr uint16_t CW_FLAG = 0x8000;

__host__ __device__ inline size_t cw_slot_bytes(int nmax) {
    const size_t tri = (size_t)nmax * (nmax - 1) / 2;
    return 12 * tri + 2 * ((tri + 1) & ~(size_t)1) + 12 * (size_t)nmax + 8 * (size_t)nmax;
}

struct CW {
    int n, ncand;
    int32_t *Wt, *Ft, *Pt, *wa, *wb, *nw;
    uint16_t* cand;
    uint8_t *list, *label, *active, *posS;
    __device__ __forceinline__ int T(int x, int y) const { return ((x * (2 * n - x - 3)) >> 1) + y - 1; }     // x < y
    __device__ __forceinline__ int TT(int x, int y) const { return x < y ? T(x, y) : T(y, x); }
    __device__ __forceinline__ int w(int x, int y) const { return Wt[TT(x, y)]; }
};

__device__ __forceinline__ long long cw_max64(long long v) {
    for (int o = 16; o > 0; o >>= 1) { const long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}

// max icf, max icp, max icp over positive candidates; composite = value << 16 | (0xffff - key):
// largest value first, smallest (a,b) on ties.  Entries flagged for the tentative round are skipped.
__device__ __forceinline__ void cw_scan(const CW& s, int lane, long long& cF, long long& cP, long long& cPpos) {
    long long bf = -1, bp = -1, bpp = -1;
    for (int i = lane; i < s.ncand; i += 32) {
        const int k = s.cand[i];
        if (k & CW_FLAG) continue;
        const int t = s.T(k >> 8, k & 0xff);
        const int w = s.Wt[t];
        if (w == FORB) continue;
        const long long kk = 0xffff - k;
        const long long f = ((long long)s.Ft[t] << 16) | kk, p = ((long long)s.Pt[t] << 16) | kk;
        bf = f > bf ? f : bf; bp = p > bp ? p : bp;
        if (w > 0) bpp = p > bpp ? p : bpp;
    }
    cF = cw_max64(bf); cP = cw_max64(bp); cPpos = cw_max64(bpp);
}

// induced-cost deltas of forbidding edge (a,b) (weight `old`), applied with `sign` (+1 apply, -1 undo).
// Reads W only; the weight itself is set to FORB by the caller once the round is committed.
__device__ __forceinline__ void cw_forbid_deltas(const CW& s, int lane, int a, int b, int old, int sign) {
    for (int t0 = 0; t0 < s.n; t0 += 32) {
        const int t = t0 + lane;
        if (t >= s.n || !s.active[t] || t == a || t == b) continue;
        const int ita = s.TT(t, a), itb = s.TT(t, b);
        const int ta = s.Wt[ita], tb = s.Wt[itb];
        if (tb != 0 && ta != 0) {
            // pair (a,t), third node b            // pair (b,t), third node a
            s.Ft[ita] -= sign * cs_tf(old, tb);       s.Ft[itb] -= sign * cs_tf(old, ta);
            s.Pt[ita] += sign * (cs_tp(FORB, tb) - cs_tp(old, tb));
            s.Pt[itb] += sign * (cs_tp(FORB, ta) - cs_tp(old, ta));
        }
    }
}

__global__ void __launch_bounds__(CW_WARPS * 32) k_cluster_warp(DB d, const int32_t* __restrict__ chains, int n_list, int nmax,
                                                                int32_t* __restrict__ work_counter) {
    extern __shared__ int32_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tri_max = nmax * (nmax - 1) / 2;
    int32_t* base = sm + (size_t)warp * (cw_slot_bytes(nmax) / 4);
    CW s;
    s.Wt = base; s.Ft = s.Wt + tri_max; s.Pt = s.Ft + tri_max;
    s.wa = s.Pt + tri_max; s.wb = s.wa + nmax; s.nw = s.wb + nmax;
    s.cand = (uint16_t*)(s.nw + nmax);
    s.list = (uint8_t*)(s.cand + ((tri_max + 1) & ~1)); s.label = s.list + nmax; s.active = s.label + nmax; s.posS = s.active + nmax;
    const unsigned lt = (1u << lane) - 1u;
    while (true) {
        int item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_list) break;
        const int c = chains[item];
        const int64_t f0 = d.frow_off[c];
        const int n = (int)(d.frow_off[c + 1] - f0);
        s.n = n;
        const int32_t* Wg = d.W + d.cw_off[c];
        // load the upper triangle and build the candidate list (row-major order)
        int ncand = 0;
        for (int x = 0; x < n; x++) for (int y0 = x + 1; y0 < n; y0 += 32) {
            const int y = y0 + lane;
            int w = 0;
            if (y < n) { w = Wg[x * n + y]; s.Wt[s.T(x, y)] = w; }
            const unsigned m = __ballot_sync(0xffffffffu, w != 0);
            if (w != 0) s.cand[ncand + __popc(m & lt)] = (uint16_t)((x << 8) | y);
            ncand += __popc(m);
        }
        s.ncand = ncand;
        for (int x = lane; x < n; x += 32) { s.active[x] = 1; s.label[x] = (uint8_t)x; }
        __syncwarp();
        // initial induced costs, one candidate per lane
        for (int i = lane; i < ncand; i += 32) {
            const int k = s.cand[i], x = k >> 8, y = k & 0xff, ti = s.T(x, y);
            const int w = s.Wt[ti];
            int f = max(w, 0), p = max(-w, 0);
            for (int t = 0; t < n; t++) {
                if (t == x || t == y) continue;
                const int wx = s.w(x, t); if (wx == 0) continue;
                const int wy = s.w(y, t);
                f += cs_tf(wx, wy); p += cs_tp(wx, wy);
            }
            s.Ft[ti] = f; s.Pt[ti] = p;
        }
        __syncwarp();
        bool force_single = false;
        long long cF, cP, cPpos;
        cw_scan(s, lane, cF, cP, cPpos);
        while (cF >= 0) {
            const int M = (int)(cF >> 16), maxP = (int)(cP >> 16), maxPpos = cPpos < 0 ? -1 : (int)(cPpos >> 16);
            if (M >= maxP) {
                // ================= merge (a,b) into a
                const int kF = 0xffff - (int)(cF & 0xffff), a = kF >> 8, b = kF & 0xff;
                int cnt = 0;
                for (int t0 = 0; t0 < n; t0 += 32) {
                    const int t = t0 + lane;
                    int xa = 0, xb = 0;
                    if (t < n) { s.posS[t] = 0xff; if (s.active[t] && t != a && t != b) { xa = s.w(a, t); xb = s.w(b, t); } }
                    const bool in = (xa != 0) || (xb != 0);
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    if (in) {
                        const int pos = cnt + __popc(m & lt);
                        s.list[pos] = (uint8_t)t; s.posS[t] = (uint8_t)pos;
                        s.wa[pos] = xa; s.wb[pos] = xb; s.nw[pos] = (xa == FORB || xb == FORB) ? FORB : xa + xb;
                    }
                    cnt += __popc(m);
                }
                __syncwarp();
                // candidate pairs inside S: terms through a and b -> term through the merged node
                for (int i = lane; i < s.ncand; i += 32) {
                    const int k = s.cand[i], x = k >> 8, y = k & 0xff;
                    const int u = s.posS[x], v = s.posS[y];
                    if (u == 0xff || v == 0xff) continue;
                    const int ti = s.T(x, y);
                    if (s.Wt[ti] == FORB) continue;
                    const int nu = s.nw[u], nv = s.nw[v], au = s.wa[u], av = s.wa[v], bu = s.wb[u], bv = s.wb[v];
                    s.Ft[ti] += cs_tf(nu, nv) - cs_tf(au, av) - cs_tf(bu, bv);
                    s.Pt[ti] += cs_tp(nu, nv) - cs_tp(au, av) - cs_tp(bu, bv);
                }
                __syncwarp();
                // fresh induced costs for the pairs (a,x), x in S (third nodes are exactly the members of S)
                for (int u = lane; u < cnt; u += 32) {
                    const int x = s.list[u], w = s.nw[u];
                    if (w != 0 && w != FORB) {
                        int f = max(w, 0), p = max(-w, 0);
                        for (int v = 0; v < cnt; v++) if (v != u) { const int t1 = s.nw[v], t2 = s.w(x, s.list[v]); f += cs_tf(t1, t2); p += cs_tp(t1, t2); }
                        const int ti = s.TT(a, x);
                        s.Ft[ti] = f; s.Pt[ti] = p;
                    }
                }
                __syncwarp();
                for (int u = lane; u < cnt; u += 32) { const int x = s.list[u]; s.Wt[s.TT(a, x)] = s.nw[u]; s.Wt[s.TT(b, x)] = 0; }
                for (int x = lane; x < n; x += 32) if (s.label[x] == b) s.label[x] = (uint8_t)a;
                if (lane == 0) { s.Wt[s.T(a, b)] = 0; s.active[b] = 0; }
                __syncwarp();
                // candidate list: drop dead pairs (weight 0 / FORB, endpoint b), then add the pairs (a,x) that became edges
                int nc = 0;
                for (int i0 = 0; i0 < s.ncand; i0 += 32) {
                    const int i = i0 + lane;
                    int k = 0; bool keep = false;
                    if (i < s.ncand) { k = s.cand[i]; const int w = s.Wt[s.T(k >> 8, k & 0xff)]; keep = w != 0 && w != FORB; }
                    const unsigned m = __ballot_sync(0xffffffffu, keep);
                    if (keep) s.cand[nc + __popc(m & lt)] = (uint16_t)k;
                    nc += __popc(m);
                }
                for (int u0 = 0; u0 < cnt; u0 += 32) {
                    const int u = u0 + lane;
                    bool add = false; int x = 0;
                    if (u < cnt) { x = s.list[u]; const int w = s.nw[u]; add = s.wa[u] == 0 && w != 0 && w != FORB; }
                    const unsigned m = __ballot_sync(0xffffffffu, add);
                    if (add) s.cand[nc + __popc(m & lt)] = (uint16_t)(a < x ? (a << 8) | x : (x << 8) | a);
                    nc += __popc(m);
                }
                s.ncand = nc;
                __syncwarp();
                force_single = false;
            } else if (force_single || maxPpos > M) {
                // ================= one sequential forbid step: the edge with the largest icp
                const int kP = 0xffff - (int)(cP & 0xffff), a = kP >> 8, b = kP & 0xff, ti = s.T(a, b);
                const int old = s.Wt[ti];
                __syncwarp();
                cw_forbid_deltas(s, lane, a, b, old, 1);
                __syncwarp();
                if (lane == 0) s.Wt[ti] = FORB;
                __syncwarp();
                force_single = false;
            } else {
                // ================= tentative round: all negative candidates with icp > M
                int nflag = 0;
                for (int i0 = 0; i0 < s.ncand; i0 += 32) {
                    const int i = i0 + lane;
                    bool fl = false;
                    if (i < s.ncand) {
                        const int k = s.cand[i], t = s.T(k >> 8, k & 0xff), w = s.Wt[t];
                        fl = !(k & CW_FLAG) && w != FORB && w < 0 && s.Pt[t] > M;
                        if (fl) s.cand[i] = (uint16_t)(k | CW_FLAG);
                    }
                    nflag += __popc(__ballot_sync(0xffffffffu, fl));
                }
                __syncwarp();
                for (int pass = 0; pass < 2; pass++) {
                    const int sign = pass == 0 ? 1 : -1;
                    for (int i0 = 0; i0 < s.ncand; i0 += 32) {
                        const int i = i0 + lane;
                        const int k = i < s.ncand ? s.cand[i] : 0;
                        unsigned m = __ballot_sync(0xffffffffu, (k & CW_FLAG) != 0);
                        for (; m; m &= m - 1) {
                            const int kk = __shfl_sync(0xffffffffu, k, __ffs(m) - 1) & 0x7fff;
                            const int a = kk >> 8, b = kk & 0xff;
                            cw_forbid_deltas(s, lane, a, b, s.Wt[s.T(a, b)], sign);
                            __syncwarp();
                        }
                    }
                    if (pass == 1) break;
                    if (nflag == 1) break;                       // a single edge = the sequential step itself
                    long long vF, vP, vPpos;
                    cw_scan(s, lane, vF, vP, vPpos);            // flagged entries are skipped
                    const bool ok = vPpos < 0 || vF < 0 || (vPpos >> 16) <= (vF >> 16);
                    if (ok) break;
                    force_single = true;                         // undo (second pass) and fall back to one sequential step
                }
                // commit (weights -> FORB) or roll back (clear the flags)
                for (int i = lane; i < s.ncand; i += 32) {
                    const int k = s.cand[i];
                    if (k & CW_FLAG) { s.cand[i] = (uint16_t)(k & 0x7fff); if (!force_single) s.Wt[s.T((k & 0x7fff) >> 8, k & 0xff)] = FORB; }
                }
                __syncwarp();
            }
            cw_scan(s, lane, cF, cP, cPpos);
        }
        for (int x = lane; x < n; x += 32) {
            const int rep = s.label[x];
            int cid = 0;
            for (int y = 0; y < rep; y++) cid += s.active[y];
            d.fr_cluster[f0 + x] = cid;
        }
        if (lane == 0) { int k = 0; for (int y = 0; y < n; y++) k += s.active[y]; d.ch_nclusters[c] = k; }
        __syncwarp();
    }
}

}  // namespace ahs
