// k_cluster.cuh — cluster editing of the read-similarity graph for chains ABOVE CC_MAXN final reads (the
// shared-memory kernel of k_chain.cuh takes everything else), one thread block per chain, state in HBM/L2.
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

