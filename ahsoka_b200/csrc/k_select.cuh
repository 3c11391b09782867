// k_select.cuh — rule R1's pooling of one read's partner pairs by radix selection (used by k_score_chain, one thread
// per read).  Self-contained (no CUDA headers) so that tests/native/select_harness.cpp can run the very same
// statements on the host against a sort-based restatement of rule R1 (oracle/core/phase_core.hpp).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AHS_HD __host__ __device__ __forceinline__
#else
#define AHS_HD inline
#endif

namespace ahs {

constexpr uint32_t CS_INVALID = 0xffffffffu;

// rcp[n] = floor(2^32 / n) + 1 for 2 <= n <= 255: high word of a * rcp[n] = floor(a / n) for a < 2^24
// (the error a * (rcp[n] n - 2^32) / (n 2^32) < 2^-8 is below 1/n, the smallest gap to the next integer)
AHS_HD uint32_t cs_rcp(uint32_t n) { return n >= 2 ? 0xffffffffu / n + 1u : 0u; }
// counter b (8 bits, 4 per word) += 1.  On the device a shared-memory reduction: no load-to-use dependency in the scan loop.
AHS_HD void cs_hist_inc(uint32_t* hw, uint32_t b) {
#if defined(__CUDA_ARCH__)
    atomicAdd(&hw[b >> 2], 1u << ((b & 3u) * 8u));
#else
    hw[b >> 2] += 1u << ((b & 3u) * 8u);
#endif
}
AHS_HD uint32_t cs_mulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// ORDER KEY of a pair with overlap 1 <= n <= 255 and k <= n disagreements: floor(65534 k / n) << 16 | n << 8 | k.
// Two distinct rates with denominators <= 255 differ by at least 1/(255*254) > 1/65534, so their 16-bit rates differ:
// the key orders exactly like rule R1's (k/n exact, then n, then k).  Always below 0xffff0000, hence never CS_INVALID.
AHS_HD uint32_t cs_order_key(uint32_t n, uint32_t k, uint32_t rcp_n) {
    const uint32_t a = k * 65534u;
    const uint32_t rate = n == 1 ? a : cs_mulhi(a, rcp_n);
    return (rate << 16) | (n << 8) | k;
}

// Pools the valid keys of row[0..len) (CS_INVALID entries are skipped): the cut = max(1, m / ploidy) smallest keys as
// same-haplotype pairs (Ks, Ns), the others as different-haplotype pairs (Kd, Nd); m = number of valid keys.  Equal
// keys are interchangeable (same n and k), so only the partition at rank cut matters, not the order: radix selection,
// 6 bits per level from the top — byte histogram of 64 buckets in hw[16] (must be all zero on entry, is all zero on
// exit), the bucket holding the cut-th key is compacted to the front of the row and refined.  The row is overwritten.
AHS_HD void cs_pool_select(uint32_t* row, int len, int ploidy, uint32_t* hw, int& Ks, int& Ns, int& Kd, int& Nd, int& m) {
    Ks = Ns = Kd = Nd = m = 0;
    int cc = len, need = 0;
    for (int shift = 26; ; shift -= 6) {
        const int sh = shift > 0 ? shift : 0; const uint32_t bm = shift >= 0 ? 63u : 3u;      // last level: the two lowest bits
        int cnt = 0;
        for (int j = 0; j < cc; j++) { const uint32_t key = row[j]; if (key != CS_INVALID) { cs_hist_inc(hw, (key >> sh) & bm); cnt++; } }
        if (shift == 26) { m = cnt; if (m == 0) return; need = m / ploidy; if (need < 1) need = 1; }
        // threshold bucket tb: the one holding the need-th smallest key (byte prefix sums of a word by one multiply: counts <= 255 in total)
        int cum = 0, fw = 16, fcum = 0; uint32_t fx = 0;
        for (int w = 0; w < 16; w++) {
            const uint32_t x = hw[w]; hw[w] = 0;
            const int tot = (int)((x * 0x01010101u) >> 24);
            if (fw == 16 && cum + tot >= need) { fw = w; fx = x; fcum = cum; }
            cum += tot;
        }
        const int b0 = (int)(fx & 255u), b1 = (int)((fx >> 8) & 255u), b2 = (int)((fx >> 16) & 255u);
        int tb = fw * 4, before = fcum;
        if (before + b0 < need) { tb++; before += b0; if (before + b1 < need) { tb++; before += b1; if (before + b2 < need) { tb++; before += b2; } } }
        need -= before;
        // below the threshold bucket -> same, above -> diff, inside -> kept (compacted to the front of the row).  Branch free:
        // (k, n) travel as k | n << 16 (the sums stay below 2^16: at most 255 keys of n, k <= 255); the store is unconditional,
        // the write cursor only moves for a kept key and never passes the read cursor.
        int wr = 0; uint32_t Ss = 0, Sd = 0;
        for (int j = 0; j < cc; j++) {
            const uint32_t key = row[j];
            const bool valid = key != CS_INVALID;
            const uint32_t b = (key >> sh) & bm;
            const uint32_t v = valid ? ((key & 255u) | ((key & 0xff00u) << 8)) : 0u;
            Ss += b < (uint32_t)tb ? v : 0u; Sd += b > (uint32_t)tb ? v : 0u;
            row[wr] = key;
            wr += (valid && b == (uint32_t)tb) ? 1 : 0;
        }
        Ks += (int)(Ss & 0xffffu); Ns += (int)(Ss >> 16); Kd += (int)(Sd & 0xffffu); Nd += (int)(Sd >> 16);
        cc = wr;                                                      // >= need >= 1 keys share the threshold bucket
        if (need == cc || shift < 0) {                                // all of them are pooled as same, or all are equal: split by count
            for (int j = 0; j < cc; j++) { const uint32_t key = row[j]; const int kq = (int)(key & 255u), nq = (int)((key >> 8) & 255u); if (j < need) { Ks += kq; Ns += nq; } else { Kd += kq; Nd += nq; } }
            return;
        }
        if (cc <= 8) {                                                // few keys left: rank them
            for (int a = 0; a < cc; a++) {
                const uint32_t ka = row[a]; int r = 0;
                for (int b = 0; b < cc; b++) { const uint32_t kb = row[b]; r += (kb < ka || (kb == ka && b < a)) ? 1 : 0; }
                const int kq = (int)(ka & 255u), nq = (int)((ka >> 8) & 255u);
                if (r < need) { Ks += kq; Ns += nq; } else { Kd += kq; Nd += nq; }
            }
            return;
        }
    }
}

}  // namespace ahs
