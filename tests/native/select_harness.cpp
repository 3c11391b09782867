// Host harness for ahsoka_b200/csrc/k_select.cuh (TEST CODE): runs the product's radix selection on random and
// degenerate partner lists and compares the pooled sums with rule R1 restated by a full sort
// (oracle/core/phase_core.hpp: rate ascending by exact cross products, then n, then k).  Prints "ok <cases>".
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../ahsoka_b200/csrc/k_select.cuh"

int main(int argc, char** argv) {
    const int cases = argc > 1 ? atoi(argv[1]) : 200000;
    std::mt19937_64 rng(12345);
    long long done = 0;
    // every (n, k) maps to an order-preserving key
    {
        std::vector<std::pair<std::pair<int, int>, uint32_t>> all;
        for (int n = 1; n <= 255; n++) for (int k = 0; k <= n; k++) all.push_back({{n, k}, ahs::cs_order_key(n, k, ahs::cs_rcp(n))});
        for (auto& a : all) {
            if (a.second >= 0xffff0000u) { printf("key out of range n=%d k=%d\n", a.first.first, a.first.second); return 1; }
            if ((a.second >> 16) != (uint32_t)((long long)a.first.second * 65534 / a.first.first)) { printf("rate wrong n=%d k=%d\n", a.first.first, a.first.second); return 1; }
        }
        std::vector<size_t> idx(all.size()); for (size_t i = 0; i < idx.size(); i++) idx[i] = i;
        auto rule = [&](size_t x, size_t y) {
            const long long nx = all[x].first.first, kx = all[x].first.second, ny = all[y].first.first, ky = all[y].first.second;
            const long long l = kx * ny, r = ky * nx;
            if (l != r) return l < r; if (nx != ny) return nx < ny; return kx < ky;
        };
        std::sort(idx.begin(), idx.end(), rule);
        for (size_t i = 1; i < idx.size(); i++) if (!(all[idx[i - 1]].second < all[idx[i]].second)) { printf("key order differs from rule R1\n"); return 1; }
    }
    for (int it = 0; it < cases; it++) {
        const int mode = it % 6;
        const int len = 1 + (int)(rng() % 160);
        const int ploidy = 1 + (int)(rng() % 6);
        std::vector<uint32_t> row(len); std::vector<std::pair<int, int>> kn;
        for (int j = 0; j < len; j++) {
            if (rng() % 5 == 0) { row[j] = ahs::CS_INVALID; continue; }
            int n, k;
            if (mode == 0) { n = 1 + rng() % 255; k = rng() % (n + 1); }
            else if (mode == 1) { n = 1 + rng() % 40; k = 0; }                              // all rates zero
            else if (mode == 2) { n = 8; k = 4; }                                           // all keys equal
            else if (mode == 3) { n = 2 * (1 + rng() % 100); k = n / 2; }                   // one rate, many n
            else if (mode == 4) { n = 1 + rng() % 30; k = (rng() % 2) ? n / 10 : n / 2; }   // two groups
            else { n = 200 + rng() % 56; k = rng() % 4; }
            row[j] = ahs::cs_order_key(n, k, ahs::cs_rcp(n)); kn.push_back({k, n});
        }
        std::sort(kn.begin(), kn.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) {
            const long long l = (long long)a.first * b.second, r = (long long)b.first * a.second;
            if (l != r) return l < r; if (a.second != b.second) return a.second < b.second; return a.first < b.first;
        });
        const int m = (int)kn.size(); const int cut = m ? std::max(1, m / ploidy) : 0;
        long long wKs = 0, wNs = 0, wKd = 0, wNd = 0;
        for (int x = 0; x < m; x++) { if (x < cut) { wKs += kn[x].first; wNs += kn[x].second; } else { wKd += kn[x].first; wNd += kn[x].second; } }
        uint32_t hw[16] = {0};
        int Ks, Ns, Kd, Nd, gm;
        ahs::cs_pool_select(row.data(), len, ploidy, hw, Ks, Ns, Kd, Nd, gm);
        for (int w = 0; w < 16; w++) if (hw[w]) { printf("histogram not cleared\n"); return 1; }
        if (gm != m || Ks != wKs || Ns != wNs || Kd != wKd || Nd != wNd) {
            printf("MISMATCH case %d mode %d len %d ploidy %d m %d/%d: got %d %d %d %d want %lld %lld %lld %lld\n", it, mode, len, ploidy, gm, m, Ks, Ns, Kd, Nd, wKs, wNs, wKd, wNd);
            return 1;
        }
        done++;
    }
    printf("ok %lld\n", done);
    return 0;
}
