// k_score.cuh — read-pair agreement scoring on the packed allele matrix (K2).
//
// Replaces ReadScoring::scoreReadsetLocal (call site reference src/alignmentstoreadset.cpp:308-311;
// algorithm: oracle/core/phase_core.hpp rule R1).  Rows are dense by bubble id, so two reads of a
// chain are already aligned: overlap n and disagreement k are AND / XOR / popcount over the words
// of the shared span, no shifting.  Reads are sorted by first position, so the partners of read i
// form index bands.
//
//   k_read_rates : one warp per read — (n,k) against every partner, per-warp bitonic sort of the
//                  Hamming rates in shared memory, pooled same-/different-haplotype rates es, ed.
//   k_pair_scores: one warp per read i, lanes over partners j > i — fixed-point log-likelihood
//                  ratio, written straight into the cluster-editing weight matrix W (both
//                  triangles).  HBM traffic: packed rows + 4 B per scored pair.
#pragma once
#include "common.cuh"
#include "device_batch.cuh"

namespace ahs {

template <int BITS>
__device__ __forceinline__ void pair_nk(const uint32_t* __restrict__ ri, const uint32_t* __restrict__ rj,
                                        int lo_b, int hi_b, int& n, int& k) {
    n = 0; k = 0;
    if (hi_b < lo_b) return;
    constexpr int PW = 32 / BITS;
    for (int w = lo_b / PW; w <= hi_b / PW; w++) word_nk<BITS>(__ldg(ri + w), __ldg(rj + w), n, k);
}

// band of candidate partners of final read f (chain-local index i): reads with
// first <= last_i (upper side) and first >= first_i - maxspan (lower side)
__device__ __forceinline__ void partner_band(const DB& d, int c, int i, int n_c, const int32_t* __restrict__ first,
                                             int first_i, int last_i, int& lo, int& hi) {
    int a = i, b = n_c;                       // hi = first index > i with first > last_i
    while (a + 1 < b) { int m = (a + b) >> 1; if (first[m] <= last_i) a = m; else b = m; }
    hi = a;                                   // inclusive
    const int bound = first_i - d.ch_maxspan[c];
    a = -1; b = i;                            // lo = first index with first >= bound
    while (a + 1 < b) { int m = (a + b) >> 1; if (first[m] >= bound) b = m; else a = m; }
    lo = b;
}

__device__ __forceinline__ uint64_t rate_key(int n, int k) {
    // monotone in (k/n, n, k): distinct rationals with n < 2^15 differ by > 2^-30
    const uint64_t rk = ((uint64_t)k << 32) / (uint64_t)n;       // <= 2^32
    return (rk << 30) | ((uint64_t)n << 15) | (uint64_t)k;
}

constexpr int RATE_SMEM_KEYS = 256;           // per warp; more partners -> global scratch

template <int BITS>
__global__ void __launch_bounds__(256) k_read_rates(DB d) {
    __shared__ uint64_t skeys[8][RATE_SMEM_KEYS];
    const int wpb = blockDim.x >> 5, lane = lane_id(), wib = threadIdx.x >> 5;
    int64_t pairs_local = 0;
    for (int64_t f = blockIdx.x * (int64_t)wpb + wib; f < d.NF; f += (int64_t)gridDim.x * wpb) {
        const int c = d.fr_chain[f];
        if (d.ch_small[c]) continue;
        const int64_t f0 = d.frow_off[c];
        const int n_c = (int)(d.frow_off[c + 1] - f0), i = (int)(f - f0);
        const int32_t* first = d.fr_first + f0; const int32_t* lastp = d.fr_last + f0;
        const int first_i = first[i], last_i = lastp[i];
        int lo, hi; partner_band(d, c, i, n_c, first, first_i, last_i, lo, hi);
        const int cap = hi - lo;                                   // candidates (self excluded)
        uint64_t* keys = skeys[wib];
        int kcap = RATE_SMEM_KEYS;
        if (cap > RATE_SMEM_KEYS) { keys = d.key_scratch + d.key_scratch_off[f]; kcap = 1; while (kcap < cap) kcap <<= 1; }
        const int words = d.ch_words[c];
        const uint32_t* rows = d.codes + d.code_off[c];
        const uint32_t* ri = rows + (int64_t)i * words;
        int m = 0;
        for (int jb = lo; jb <= hi; jb += 32) {
            const int j = jb + lane;
            int n = 0, k = 0;
            if (j <= hi && j != i) pair_nk<BITS>(ri, rows + (int64_t)j * words, max(first_i, first[j]), min(last_i, lastp[j]), n, k);
            const unsigned bal = __ballot_sync(0xffffffffu, n > 0);
            if (n > 0) keys[m + __popc(bal & ((1u << lane) - 1u))] = rate_key(n, k);
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
        if (lane == 0) { d.es[f] = (uint16_t)es; d.ed[f] = (uint16_t)ed; pairs_local += m; if (m) atomicAdd(&d.ch_pairs2[c], (unsigned long long)m); }
    }
    if (lane == 0 && pairs_local) atomicAdd((unsigned long long*)d.tot_pairs, (unsigned long long)pairs_local);
}

__device__ __forceinline__ int32_t pair_weight(const DB& d, int n, int k, int es_i, int ed_i, int es_j, int ed_j) {
    int es = (es_i + es_j) >> 1, ed = (ed_i + ed_j) >> 1;
    es = min(max(es, 10), 460);
    ed = min(max(ed, es + 51), 972);
    const int64_t s20 = (int64_t)k * (d.ln[es] - d.ln[ed]) + (int64_t)(n - k) * (d.ln1[es] - d.ln1[ed]);
    int64_t w = floordiv1024(s20);
    w = w > W_CLAMP ? W_CLAMP : (w < -W_CLAMP ? -W_CLAMP : w);
    return (int32_t)w;
}

template <int BITS>
__global__ void __launch_bounds__(256) k_pair_scores(DB d) {
    const int wpb = blockDim.x >> 5, lane = lane_id();
    for (int64_t f = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); f < d.NF; f += (int64_t)gridDim.x * wpb) {
        const int c = d.fr_chain[f];
        if (d.ch_status[c] != AHS_CHAIN_OK || d.ch_small[c]) continue;
        const int64_t f0 = d.frow_off[c];
        const int n_c = (int)(d.frow_off[c + 1] - f0), i = (int)(f - f0);
        const int32_t* first = d.fr_first + f0; const int32_t* lastp = d.fr_last + f0;
        const int first_i = first[i], last_i = lastp[i];
        int a = i, b = n_c;
        while (a + 1 < b) { int m = (a + b) >> 1; if (first[m] <= last_i) a = m; else b = m; }
        const int hi = a;
        const int words = d.ch_words[c];
        const uint32_t* rows = d.codes + d.code_off[c];
        const uint32_t* ri = rows + (int64_t)i * words;
        int32_t* W = d.W + d.cw_off[c];
        const int es_i = d.es[f], ed_i = d.ed[f];
        for (int j = i + 1 + lane; j <= hi; j += 32) {
            int n, k; pair_nk<BITS>(ri, rows + (int64_t)j * words, first[j], min(last_i, lastp[j]), n, k);
            if (n > 0) {
                const int32_t w = pair_weight(d, n, k, es_i, ed_i, d.es[f0 + j], d.ed[f0 + j]);
                W[(int64_t)i * n_c + j] = w; W[(int64_t)j * n_c + i] = w;
            }
        }
    }
}

}  // namespace ahs
