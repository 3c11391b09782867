// k_chain.cuh — read-pair scoring FUSED with cluster editing, one thread block per chain, every
// intermediate in shared memory.
//
// Replaces ReadScoring::scoreReadsetLocal + ClusterEditingSolver::run (call sites reference
// src/alignmentstoreadset.cpp:308-315; algorithms: oracle/core/phase_core.hpp rules R1 and R2) for
// chains with at most CC_MAXN final reads — every chain of BASELINE configs 2-4.  Larger chains take
// the HBM-resident path (k_read_rates / k_pair_scores / k_cluster_edit).
//
// Per chain the block reads only the packed allele rows (code_bytes per cell) and 12 B of row
// descriptors per read from HBM and writes 4 B of cluster label per read: the pair scores never
// leave the SM ("fuse K2 into clustering input").
//
// Shared-memory layout for a class with at most nmax reads, tri = nmax(nmax-1)/2 pairs:
//   W   int32[tri]    Q10 pair weight, upper triangle row-major; 0 = no edge, FORB = forbidden.
//                     During scoring it holds (n << 16 | k): overlap and disagreement counts.
//   FP  int2 [tri]    x = icf, y = icp of the pair (rule R2); the rate-sort scratch during scoring.
//   cand u32 [tri]    live candidate pairs: (triangle index << 16) | (a << 8) | b, a < b.
//   + O(nmax) vectors (first/last position, es/ed rates, merge lists, labels).
// int32 is exact: every weight and induced cost is bounded by the sum of |w| over the chain's pairs
// <= tri * 2^17 < 2^31 for nmax <= 181.
//
// The greedy loop is sequential by definition (argmax -> merge or forbid); each step is spread over
// the block: strided candidate scans with redux.sync reductions, one barrier per reduction.  Runs
// of single-edge forbids are batched EXACTLY (see "round" below).
#pragma once
#include "common.cuh"
#include "device_batch.cuh"
#include "k_score.cuh"

namespace ahs {

constexpr int CC_MAXN = 176;
constexpr uint32_t CC_DEAD = 0xffffffffu;

__host__ __device__ inline int cc_flist_cap(int nmax) { return 4 * nmax; }

__host__ __device__ inline size_t cc_smem_bytes(int nmax, int nthreads) {
    const size_t tri = (size_t)nmax * (nmax - 1) / 2;
    size_t b = tri * 16;                                  // W, FP, cand
    b += (size_t)nmax * (4 * 5);                          // first, last, wa, wb, nw
    b += (size_t)nmax * (2 * 2);                          // es, ed
    b += (size_t)cc_flist_cap(nmax) * 8;                  // flagged (cand entry, old weight)
    b += (size_t)nmax * 4;                                // list, posS, label, active (u8 each)
    b += (size_t)(nthreads / 32) * 8 * 4 * 2;             // reduction scratch, double buffered
    b += 64;                                              // scalars
    return (b + 15) & ~(size_t)15;
}

struct CC {
    int n, tri;
    int32_t* W; int2* FP; uint32_t* cand;
    int32_t *first, *last, *wa, *wb, *nw;
    uint16_t *es, *ed;
    uint32_t* fl_c; int32_t* fl_old;
    uint8_t *list, *posS, *label, *active;
    int32_t* red; int32_t* scal;      // scal[0] = ncand, [1] = nflag / list count, [2] = pair count
    __device__ __forceinline__ int T(int x, int y) const { return ((x * (2 * n - x - 3)) >> 1) + y - 1; }     // x < y
    __device__ __forceinline__ int TT(int x, int y) const { return x < y ? T(x, y) : T(y, x); }
    __device__ __forceinline__ int w(int x, int y) const { return W[TT(x, y)]; }
};

struct CCScan { int M, kF, maxP, kP, maxPpos, live; };

// max icf (M, ties -> smallest pair key), max icp (same tie rule), max icp over positive-weight
// candidates, number of live candidates.  Every thread returns the same values.
template <int NT>
__device__ __forceinline__ CCScan cc_scan(const CC& s, int tid, int& phase) {
    constexpr int NW = NT / 32;
    int bf = -1, kf = 0xffff, bp = -1, kp = 0xffff, bpp = -1, live = 0;
    const int ncand = s.scal[0];
    for (int i = tid; i < ncand; i += NT) {
        const uint32_t c = s.cand[i];
        const int ti = (int)(c >> 16), key = (int)(c & 0xffffu);
        const int w = s.W[ti];
        if (w == 0 || w == FORB) continue;
        const int2 fp = s.FP[ti];
        live++;
        if (fp.x > bf || (fp.x == bf && key < kf)) { bf = fp.x; kf = key; }
        if (fp.y > bp || (fp.y == bp && key < kp)) { bp = fp.y; kp = key; }
        if (w > 0 && fp.y > bpp) bpp = fp.y;
    }
    const int mf = __reduce_max_sync(0xffffffffu, bf);
    const int mkf = __reduce_min_sync(0xffffffffu, bf == mf ? kf : 0xffff);
    const int mp = __reduce_max_sync(0xffffffffu, bp);
    const int mkp = __reduce_min_sync(0xffffffffu, bp == mp ? kp : 0xffff);
    const int mpp = __reduce_max_sync(0xffffffffu, bpp);
    const int lv = __reduce_add_sync(0xffffffffu, live);
    CCScan o;
    if (NW == 1) { o.M = mf; o.kF = mkf; o.maxP = mp; o.kP = mkp; o.maxPpos = mpp; o.live = lv; __syncwarp(); return o; }
    int32_t* r = s.red + phase * (NW * 8);
    if ((tid & 31) == 0) { int32_t* q = r + (tid >> 5) * 8; q[0] = mf; q[1] = mkf; q[2] = mp; q[3] = mkp; q[4] = mpp; q[5] = lv; }
    __syncthreads();
    o.M = -1; o.kF = 0xffff; o.maxP = -1; o.kP = 0xffff; o.maxPpos = -1; o.live = 0;
#pragma unroll
    for (int wv = 0; wv < NW; wv++) {
        const int32_t* q = r + wv * 8;
        const int a = q[0], ka = q[1], b = q[2], kb = q[3];
        if (a > o.M || (a == o.M && ka < o.kF)) { o.M = a; o.kF = ka; }
        if (b > o.maxP || (b == o.maxP && kb < o.kP)) { o.maxP = b; o.kP = kb; }
        o.maxPpos = max(o.maxPpos, q[4]); o.live += q[5];
    }
    phase ^= 1;
    return o;
}

// in-place compaction of the candidate list (drops pairs whose weight became 0 or FORB)
template <int NT, int MAXPER>
__device__ __forceinline__ void cc_compact(const CC& s, int tid, int& phase) {
    constexpr int NW = NT / 32;
    const int ncand = s.scal[0];
    const int lane = tid & 31, wid = tid >> 5;
    const int chunk = (ncand + NW - 1) / NW, lo = wid * chunk, hi = min(ncand, lo + chunk);
    uint32_t keep[MAXPER]; int cnt = 0;
#pragma unroll
    for (int x = 0; x < MAXPER; x++) {
        const int i = lo + x * 32 + lane;
        uint32_t c = CC_DEAD;
        if (i < hi) { c = s.cand[i]; const int w = s.W[c >> 16]; if (w == 0 || w == FORB) c = CC_DEAD; }
        keep[x] = c;
        cnt += __popc(__ballot_sync(0xffffffffu, c != CC_DEAD));
    }
    int32_t* r = s.red + phase * (NW * 8);
    if (lane == 0) r[wid * 8 + 6] = cnt;
    __syncthreads();                                        // all reads done, all counts visible
    int base = 0, total = 0;
#pragma unroll
    for (int wv = 0; wv < NW; wv++) { const int v = r[wv * 8 + 6]; if (wv < wid) base += v; total += v; }
#pragma unroll
    for (int x = 0; x < MAXPER; x++) {
        const uint32_t c = keep[x];
        const unsigned m = __ballot_sync(0xffffffffu, c != CC_DEAD);
        if (c != CC_DEAD) s.cand[base + __popc(m & ((1u << lane) - 1u))] = c;
        base += __popc(m);
    }
    if (tid == 0) s.scal[0] = total;
    phase ^= 1;
    __syncthreads();
}

__device__ __forceinline__ unsigned long long cc_globaltimer() {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
__device__ __forceinline__ int cc_tf(int x, int y) { return (x > 0 && y > 0) ? min(x, y) : 0; }
__device__ __forceinline__ int cc_tp(int x, int y) {
    if (x > 0 && y < 0) return y == FORB ? x : min(x, -y);
    if (x < 0 && y > 0) return x == FORB ? y : min(-x, y);
    return 0;
}

template <int BITS, int NT, int MAXPER>
__global__ void __launch_bounds__(NT) k_score_cluster(DB d, const int32_t* __restrict__ chains, int n_list, int nmax,
                                                       int32_t* __restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char cc_sm[];
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int tri_max = nmax * (nmax - 1) / 2;
    CC s;
    {
        unsigned char* p = cc_sm;
        s.FP = (int2*)p; p += (size_t)tri_max * 8;
        s.W = (int32_t*)p; p += (size_t)tri_max * 4;
        s.cand = (uint32_t*)p; p += (size_t)tri_max * 4;
        s.first = (int32_t*)p; p += nmax * 4; s.last = (int32_t*)p; p += nmax * 4;
        s.wa = (int32_t*)p; p += nmax * 4; s.wb = (int32_t*)p; p += nmax * 4; s.nw = (int32_t*)p; p += nmax * 4;
        s.fl_c = (uint32_t*)p; p += cc_flist_cap(nmax) * 4; s.fl_old = (int32_t*)p; p += cc_flist_cap(nmax) * 4;
        s.red = (int32_t*)p; p += NW * 8 * 4 * 2;
        s.scal = (int32_t*)p; p += 64;
        s.es = (uint16_t*)p; p += nmax * 2; s.ed = (uint16_t*)p; p += nmax * 2;
        s.list = p; p += nmax; s.posS = p; p += nmax; s.label = p; p += nmax; s.active = p; p += nmax;
    }
    const int FL_CAP = cc_flist_cap(nmax);
    int phase = 0;
    int64_t pairs_total = 0;
    unsigned long long t_score = 0, t_cluster = 0;      // thread 0: nanoseconds spent in the two halves (globaltimer)
    while (true) {
        __syncthreads();
        if (tid == 0) s.scal[3] = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s.scal[3];
        if (item >= n_list) break;
        const int c = chains[item];
        const int64_t f0 = d.frow_off[c];
        const int n = (int)(d.frow_off[c + 1] - f0);
        s.n = n; s.tri = n * (n - 1) / 2;
        const int tri = s.tri;
        const int words = d.ch_words[c];
        unsigned long long t_start = 0;
        if (tid == 0) t_start = cc_globaltimer();
        const uint32_t* rows = d.codes + d.code_off[c];
        // ================================================================ scoring (rule R1)
        for (int x = tid; x < n; x += NT) { s.first[x] = d.fr_first[f0 + x]; s.last[x] = d.fr_last[f0 + x]; s.active[x] = 1; s.label[x] = (uint8_t)x; }
        for (int x = tid; x < tri; x += NT) s.W[x] = 0;
        if (tid == 0) { s.scal[0] = 0; s.scal[2] = 0; s.scal[4] = 0; }
        __syncthreads();
        // overlap / disagreement counts, once per pair: warp per row, lanes over later reads of the band
        for (int i = wid; i < n - 1; i += NW) {
            const int first_i = s.first[i], last_i = s.last[i];
            const uint32_t* ri = rows + (int64_t)i * words;
            const int rowbase = s.T(i, 0);
            for (int j0 = i + 1; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                const bool valid = j < n && s.first[j] <= last_i;
                if (valid) {
                    int nn, kk; pair_nk<BITS>(ri, rows + (int64_t)j * words, max(first_i, s.first[j]), min(last_i, s.last[j]), nn, kk);
                    if (nn > 0) s.W[rowbase + j] = (nn << 16) | kk;
                }
                if (!__any_sync(0xffffffffu, valid)) break;          // reads are sorted by first position
            }
        }
        __syncthreads();
        // local rates per read: partners ordered by Hamming rate, pooled same / different rates
        {
            uint64_t* keys = (uint64_t*)s.FP + (size_t)wid * (tri / NW);     // tri/NW >= pow2(n-1) is checked on the host
            int pl = 0;
            for (int i = wid; i < n; i += NW) {
                int m = 0;
                for (int j0 = 0; j0 < n; j0 += 32) {
                    const int j = j0 + lane;
                    const int nk = (j < n && j != i) ? s.W[s.TT(i, j)] : 0;
                    const unsigned bal = __ballot_sync(0xffffffffu, nk != 0);
                    if (nk != 0) keys[m + __popc(bal & lt)] = rate_key(nk >> 16, nk & 0xffff);
                    m += __popc(bal);
                }
                int N = 1; while (N < m) N <<= 1;
                for (int x = m + lane; x < N; x += 32) keys[x] = ~0ull;
                __syncwarp();
                for (int kk = 2; kk <= N; kk <<= 1)
                    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                        for (int x = lane; x < N; x += 32) {
                            const int y = x ^ jj;
                            if (y > x) {
                                const uint64_t a = keys[x], b = keys[y];
                                const bool up = (x & kk) == 0;
                                if ((a > b) == up) { keys[x] = b; keys[y] = a; }
                            }
                        }
                        __syncwarp();
                    }
                uint32_t es = 0, ed = 0;
                if (m > 0) {
                    const int cut = max(1, m / d.ploidy);
                    int64_t Ks = 0, Ns = 0, Kd = 0, Nd = 0;
                    for (int x = lane; x < m; x += 32) {
                        const uint64_t key = keys[x];
                        const int64_t kk = (int64_t)(key & 0x7fff), nn = (int64_t)((key >> 15) & 0x7fff);
                        if (x < cut) { Ks += kk; Ns += nn; } else { Kd += kk; Nd += nn; }
                    }
                    Ks = warp_sum_i64(Ks); Ns = warp_sum_i64(Ns); Kd = warp_sum_i64(Kd); Nd = warp_sum_i64(Nd);
                    es = (uint32_t)((Ks * 1024 + Ns / 2) / Ns);
                    ed = Nd > 0 ? (uint32_t)((Kd * 1024 + Nd / 2) / Nd) : es;
                }
                __syncwarp();
                if (lane == 0) { s.es[i] = (uint16_t)es; s.ed[i] = (uint16_t)ed; }
                pl += m;
            }
            if (lane == 0) pairs_total += pl;
        }
        __syncthreads();
        // pair weights (fixed-point log likelihood ratio) and the candidate list
        for (int i = wid; i < n - 1; i += NW) {
            const int rowbase = s.T(i, 0);
            const int es_i = s.es[i], ed_i = s.ed[i];
            for (int j0 = i + 1; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                int w = 0, ti = 0;
                if (j < n) {
                    ti = rowbase + j;
                    const int nk = s.W[ti];
                    if (nk != 0) { w = pair_weight(d, nk >> 16, nk & 0xffff, es_i, ed_i, s.es[j], s.ed[j]); s.W[ti] = w; }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, w != 0);
                if (bal) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s.scal[0], __popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (w != 0) s.cand[base + __popc(bal & lt)] = ((uint32_t)ti << 16) | (uint32_t)((i << 8) | j);
                }
            }
        }
        __syncthreads();
        // ================================================================ cluster editing (rule R2)
        if (tid == 0) { const unsigned long long t1 = cc_globaltimer(); t_score += t1 - t_start; t_start = t1; }
        // initial induced costs, one candidate per thread
        {
            const int ncand = s.scal[0];
            for (int i = tid; i < ncand; i += NT) {
                const uint32_t cd = s.cand[i];
                const int ti = (int)(cd >> 16), x = (int)((cd >> 8) & 0xff), y = (int)(cd & 0xff);
                const int w = s.W[ti];
                int f = max(w, 0), p = max(-w, 0);
                for (int t = 0; t < n; t++) {
                    if (t == x || t == y) continue;
                    const int wx = s.w(x, t); if (wx == 0) continue;
                    const int wy = s.w(y, t);
                    f += cc_tf(wx, wy); p += cc_tp(wx, wy);
                }
                s.FP[ti] = make_int2(f, p);
            }
        }
        __syncthreads();
        bool force_single = false;
        int n_at_compact = s.scal[0];
        CCScan so = cc_scan<NT>(s, tid, phase);
        while (so.M >= 0) {
            if (so.M >= so.maxP) {
                // ------------------------------------------------ merge (a,b) into a
                const int a = so.kF >> 8, b = so.kF & 0xff;
                const int ncand = s.scal[0];           // read before anyone appends to the list below
                if (tid == 0) s.scal[1] = 0;
                __syncthreads();
                for (int t0 = wid * 32; t0 < n; t0 += NT) {
                    const int t = t0 + lane;
                    int xa = 0, xb = 0;
                    if (t < n) { s.posS[t] = 0xff; if (s.active[t] && t != a && t != b) { xa = s.w(a, t); xb = s.w(b, t); } }
                    const bool in = (xa != 0) || (xb != 0);
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    if (m) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&s.scal[1], __popc(m));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (in) {
                            const int pos = base + __popc(m & lt);
                            s.list[pos] = (uint8_t)t; s.posS[t] = (uint8_t)pos;
                            s.wa[pos] = xa; s.wb[pos] = xb; s.nw[pos] = (xa == FORB || xb == FORB) ? FORB : xa + xb;
                        }
                    }
                }
                __syncthreads();
                const int cnt = s.scal[1];
                // candidate pairs inside S: the terms through a and b become one term through the merged node
                for (int i = tid; i < ncand; i += NT) {
                    const uint32_t cd = s.cand[i];
                    const int x = (int)((cd >> 8) & 0xff), y = (int)(cd & 0xff);
                    const int u = s.posS[x], v = s.posS[y];
                    if (u == 0xff || v == 0xff) continue;
                    const int ti = (int)(cd >> 16);
                    const int w = s.W[ti];
                    if (w == FORB || w == 0) continue;
                    const int nu = s.nw[u], nv = s.nw[v], au = s.wa[u], av = s.wa[v], bu = s.wb[u], bv = s.wb[v];
                    int2 fp = s.FP[ti];
                    fp.x += cc_tf(nu, nv) - cc_tf(au, av) - cc_tf(bu, bv);
                    fp.y += cc_tp(nu, nv) - cc_tp(au, av) - cc_tp(bu, bv);
                    s.FP[ti] = fp;
                }
                // fresh induced costs of the pairs (a,x), x in S (their third nodes are exactly the members of S),
                // new weights of rows a and b, new candidates
                for (int u0 = wid * 32; u0 < cnt; u0 += NT) {
                    const int u = u0 + lane;
                    bool add = false; int x = 0, ti = 0;
                    if (u < cnt) {
                        x = s.list[u]; const int w = s.nw[u];
                        ti = s.TT(a, x);
                        if (w != 0 && w != FORB) {
                            int f = max(w, 0), p = max(-w, 0);
                            for (int v = 0; v < cnt; v++) if (v != u) { const int t1 = s.nw[v], t2 = s.w(x, s.list[v]); f += cc_tf(t1, t2); p += cc_tp(t1, t2); }
                            s.FP[ti] = make_int2(f, p);
                            add = s.wa[u] == 0;
                        } else if (w == 0 && s.wa[u] != 0 && s.wa[u] != FORB) {
                            s.scal[4] = 1;     // exact cancellation: the pair keeps a dead list entry that a later merge
                        }                      // could revive next to a fresh one -> compact before that can happen
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, add);
                    if (m) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&s.scal[0], __popc(m));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (add) s.cand[base + __popc(m & lt)] = ((uint32_t)ti << 16) | (uint32_t)(a < x ? (a << 8) | x : (x << 8) | a);
                    }
                }
                __syncthreads();       // every read of the old rows a, b is done
                for (int u = tid; u < cnt; u += NT) { const int x = s.list[u]; s.W[s.TT(a, x)] = s.nw[u]; s.W[s.TT(b, x)] = 0; }
                for (int x = tid; x < n; x += NT) if (s.label[x] == b) s.label[x] = (uint8_t)a;
                if (tid == 0) { s.W[s.T(a, b)] = 0; s.active[b] = 0; }
                __syncthreads();
                force_single = false;
                so = cc_scan<NT>(s, tid, phase);
                if (s.scal[4] || (so.live * 4 < n_at_compact * 3 && s.scal[0] > NT)) {
                    __syncthreads();
                    if (tid == 0) s.scal[4] = 0;
                    cc_compact<NT, MAXPER>(s, tid, phase); n_at_compact = s.scal[0];
                }
            } else if (force_single || so.maxPpos > so.M) {
                // ------------------------------------------------ one sequential forbid: the edge with the largest icp
                const int a = so.kP >> 8, b = so.kP & 0xff, tab = s.T(a, b);
                const int old = s.W[tab];
                __syncthreads();
                for (int t = tid; t < n; t += NT) {
                    if (!s.active[t] || t == a || t == b) continue;
                    const int ita = s.TT(t, a), itb = s.TT(t, b);
                    const int ta = s.W[ita], tb = s.W[itb];
                    if (ta != 0 && tb != 0) {
                        int2 fa = s.FP[ita], fb = s.FP[itb];
                        fa.x -= cc_tf(old, tb); fa.y += cc_tp(FORB, tb) - cc_tp(old, tb);        // pair (a,t), third node b
                        fb.x -= cc_tf(old, ta); fb.y += cc_tp(FORB, ta) - cc_tp(old, ta);        // pair (b,t), third node a
                        s.FP[ita] = fa; s.FP[itb] = fb;
                    }
                }
                if (tid == 0) s.W[tab] = FORB;
                __syncthreads();
                force_single = false;
                so = cc_scan<NT>(s, tid, phase);
            } else {
                // ------------------------------------------------ round: forbid negative candidates with icp > M at once.
                // Forbidding a negative edge changes no icf and only raises icp values, so every such edge stays
                // eligible until it is forbidden: the set forbidden before the next merge is a fixed point that does
                // not depend on the order, PROVIDED no positive edge would be picked in between.  That proviso is
                // checked on the state after the round (max icp over positive candidates <= new max icf); if it
                // fails the round is undone and one sequential step is taken instead.
                const int M = so.M;
                if (tid == 0) s.scal[1] = 0;
                __syncthreads();
                const int ncand = s.scal[0];
                for (int i = tid; i < ncand; i += NT) {
                    const uint32_t cd = s.cand[i];
                    const int ti = (int)(cd >> 16);
                    const int w = s.W[ti];
                    if (w < 0 && w != FORB && s.FP[ti].y > M) {
                        const int pos = atomicAdd(&s.scal[1], 1);
                        if (pos < FL_CAP) { s.fl_c[pos] = cd; s.fl_old[pos] = w; s.W[ti] = FORB; }
                    }
                }
                __syncthreads();
                const int nflag = min(s.scal[1], FL_CAP);
                for (int sign = 1; ; sign = -1) {
                    for (int idx = tid; idx < nflag * n; idx += NT) {
                        const int e = idx / n, t = idx - e * n;
                        const uint32_t cd = s.fl_c[e];
                        const int a = (int)((cd >> 8) & 0xff), b = (int)(cd & 0xff), old = s.fl_old[e];
                        if (!s.active[t] || t == a || t == b) continue;
                        const int ita = s.TT(t, a), itb = s.TT(t, b);
                        const int ta = s.W[ita], tb = s.W[itb];
                        if (ta != 0 && tb != 0) {
                            const int da = cc_tp(FORB, tb) - cc_tp(old, tb), db = cc_tp(FORB, ta) - cc_tp(old, ta);
                            if (da) atomicAdd(&s.FP[ita].y, sign * da);
                            if (db) atomicAdd(&s.FP[itb].y, sign * db);
                        }
                    }
                    __syncthreads();
                    if (sign < 0) break;
                    const CCScan v = cc_scan<NT>(s, tid, phase);
                    const bool ok = nflag == 1 || v.maxPpos < 0 || v.M < 0 || v.maxPpos <= v.M;
                    if (ok) { so = v; break; }
                    force_single = true;                   // undo below, then one sequential step on the unchanged `so`
                }
                if (force_single) {
                    for (int e = tid; e < nflag; e += NT) s.W[s.fl_c[e] >> 16] = s.fl_old[e];
                    __syncthreads();
                }
            }
        }
        // ---- clusters: numbered by smallest member (= representative), ascending
        for (int x = tid; x < n; x += NT) {
            const int rep = s.label[x];
            int cid = 0;
            for (int y = 0; y < rep; y++) cid += s.active[y];
            d.fr_cluster[f0 + x] = cid;
        }
        if (tid == 0) { int k = 0; for (int y = 0; y < n; y++) k += s.active[y]; d.ch_nclusters[c] = k; t_cluster += cc_globaltimer() - t_start; }
    }
    if (lane == 0 && pairs_total) atomicAdd((unsigned long long*)d.tot_pairs, (unsigned long long)pairs_total);
    if (tid == 0) { atomicAdd(d.t_phase, t_score); atomicAdd(d.t_phase + 1, t_cluster); }
}

}  // namespace ahs
