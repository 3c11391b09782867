// host_plan.hpp — how one call's chains are shared out over several devices.  Plain C++, no CUDA: used by
// phase_batch.cu (ahs_phase_batch_multi) and exported as ahs_plan_shares for the host tools / tests.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace ahs {

// Cuts 0 = cut[0] <= cut[1] <= ... <= cut[G] = C of the chains into G contiguous shares such that the largest share cost is as
// small as contiguous cuts allow (to the precision of the bisection, 1e-6 relative).  pre = prefix sums of the chain costs
// (C + 1 entries, pre[0] = 0, non-decreasing).  Bisection on the bound M; a bound is feasible if packing greedily from the
// left — every share takes chains while it stays <= M, and at least one chain — covers all chains with G shares.
inline std::vector<int64_t> balanced_contiguous_cuts(const std::vector<double>& pre, int G) {
    const int64_t C = (int64_t)pre.size() - 1;
    std::vector<int64_t> cuts((size_t)std::max(G, 1) + 1, 0);
    if (G < 1 || C <= 0) { for (auto& c : cuts) c = std::max<int64_t>(C, 0); cuts[0] = 0; return cuts; }
    double cmax = 0;
    for (int64_t c = 0; c < C; c++) cmax = std::max(cmax, pre[c + 1] - pre[c]);
    auto pack = [&](double M, std::vector<int64_t>* out) {
        int64_t c = 0;
        for (int g = 0; g < G; g++) {
            if (c < C) {
                int64_t e = (int64_t)(std::upper_bound(pre.begin() + c + 1, pre.end(), pre[c] + M) - pre.begin()) - 1;     // last prefix <= pre[c] + M
                c = std::min(std::max(e, c + 1), C);              // a chain above the bound still goes somewhere
            }
            if (out) (*out)[g + 1] = c;
        }
        return c >= C;
    };
    double lo = std::max(cmax, pre[C] / G), hi = pre[C] + cmax;
    if (pack(lo, nullptr)) hi = lo;
    for (int it = 0; it < 60 && hi - lo > 1e-6 * hi; it++) { const double mid = 0.5 * (lo + hi); if (pack(mid, nullptr)) hi = mid; else lo = mid; }
    pack(hi, &cuts);
    cuts[0] = 0; cuts[G] = C;
    return cuts;
}

}  // namespace ahs
