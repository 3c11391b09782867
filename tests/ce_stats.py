"""Step statistics of cluster editing (rule R2) on a cfg2 sample: merges, single forbids, batched
rounds, failed rounds, candidate counts.  Design aid for the cluster-editing kernel; uses the CPU
oracle only to obtain the pair weights (test infrastructure, not product code)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ahsoka_b200 import synth
from tests.oracle_binding import oracle_phase, oracle_score, oracle_cluster

FORB = -(1 << 31)


def tf(x, y):
    return np.where((x > 0) & (y > 0), np.minimum(x, y), 0)


def absw(x):
    return np.where(x == FORB, 1 << 40, np.abs(x))


def tp(x, y):
    r = np.zeros(np.broadcast(x, y).shape, dtype=np.int64)
    m1 = (x > 0) & (y < 0); m2 = (x < 0) & (y > 0)
    r = np.where(m1, np.minimum(x, absw(y)), r)
    r = np.where(m2, np.minimum(absw(x), y), r)
    return r


def full_FP(W, active):
    n = W.shape[0]
    idx = np.where(active)[0]
    Wa = W[np.ix_(idx, idx)].astype(np.int64)
    k = len(idx)
    F = np.zeros((k, k), dtype=np.int64); P = np.zeros((k, k), dtype=np.int64)
    for a in range(k):
        x = Wa[a][None, :]            # W[a,c]
        Fa = tf(np.broadcast_to(Wa[a][None, :], (k, k)), Wa)   # [b,c] = tf(W[a,c], W[b,c])
        Pa = tp(np.broadcast_to(Wa[a][None, :], (k, k)), Wa)
        Fa[:, a] = 0; Pa[:, a] = 0
        fa = Fa.sum(1) - Fa[np.arange(k), np.arange(k)]
        pa = Pa.sum(1) - Pa[np.arange(k), np.arange(k)]
        F[a] = fa + np.maximum(Wa[a], 0); P[a] = pa + np.maximum(-np.where(Wa[a] == FORB, 0, Wa[a]), 0)
    return idx, Wa, F, P


def simulate(n, pi, pj, pw):
    W = np.zeros((n, n), dtype=np.int64)
    W[pi, pj] = pw; W[pj, pi] = pw
    active = np.ones(n, dtype=bool)
    st = dict(n=n, ncand0=int((pw != 0).sum()), merges=0, singles=0, rounds=0, round_edges=0, fails=0, scans=0, cand_visits=0, S_sum=0)
    while True:
        idx, Wa, F, P = full_FP(W, active)
        k = len(idx)
        cand = (Wa != 0) & (Wa != FORB) & np.triu(np.ones((k, k), dtype=bool), 1)
        nc = int(cand.sum())
        st["scans"] += 1; st["cand_visits"] += nc
        if nc == 0:
            break
        Fm = np.where(cand, F, -1); Pm = np.where(cand, P, -1)
        M = Fm.max(); maxP = Pm.max()
        if M >= maxP:
            a, b = np.unravel_index(np.argmax(Fm), Fm.shape)   # first max in row-major = smallest (a,b)
            ga, gb = idx[a], idx[b]
            S = active & ((W[ga] != 0) | (W[gb] != 0)); S[ga] = S[gb] = False
            st["S_sum"] += int(S.sum())
            new = np.where((W[ga] == FORB) | (W[gb] == FORB), FORB, W[ga] + W[gb])
            new[~S] = 0
            W[ga] = new; W[:, ga] = new; W[gb] = 0; W[:, gb] = 0; W[ga, ga] = 0
            active[gb] = False
            st["merges"] += 1
        else:
            maxPpos = np.where(cand & (Wa > 0), P, -1).max()
            if maxPpos > M:
                a, b = np.unravel_index(np.argmax(Pm), Pm.shape)
                W[idx[a], idx[b]] = FORB; W[idx[b], idx[a]] = FORB
                st["singles"] += 1
                continue
            # batched round: fixed point of "forbid negative candidates with icp > M"
            fl = cand & (Wa < 0) & (P > M)
            Wt = W.copy()
            ii, jj = np.where(fl)
            Wt[idx[ii], idx[jj]] = FORB; Wt[idx[jj], idx[ii]] = FORB
            idx2, Wa2, F2, P2 = full_FP(Wt, active)
            cand2 = (Wa2 != 0) & (Wa2 != FORB) & np.triu(np.ones((k, k), dtype=bool), 1)
            st["scans"] += 1; st["cand_visits"] += int(cand2.sum())
            ok = True
            if len(ii) > 1 and cand2.any():
                M2 = np.where(cand2, F2, -1).max(); pp2 = np.where(cand2 & (Wa2 > 0), P2, -1).max()
                ok = pp2 <= M2
            if ok:
                W = Wt; st["rounds"] += 1; st["round_edges"] += len(ii)
            else:
                st["fails"] += 1
                a, b = np.unravel_index(np.argmax(Pm), Pm.shape)
                W[idx[a], idx[b]] = FORB; W[idx[b], idx[a]] = FORB
                st["singles"] += 1
    # labels
    return st, W, active


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.001
    workload = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
    b = synth.generate(synth.config(workload, scale))
    r = oracle_phase(b)
    tot = []
    for c in range(b.n_chains):
        f0, f1 = int(r.read_off[c]), int(r.read_off[c + 1])
        n = f1 - f0
        if n < 2:
            continue
        rows = []
        for f in range(f0, f1):
            lo, hi = int(r.cell_off[f]), int(r.cell_off[f + 1])
            rows.append((r.cell_pos[lo:hi], r.cell_allele[lo:hi].astype(np.int32)))
        sc = oracle_score(rows, int(b.ploidy))
        st, W, active = simulate(n, sc["i"], sc["j"], sc["w"].astype(np.int64))
        k, label = oracle_cluster(n, sc["i"], sc["j"], sc["w"])
        st["nclusters"] = int(active.sum()); st["oracle_k"] = k
        tot.append(st)
        print(st, flush=True)
    keys = [k for k in tot[0] if k != "n"]
    print("chains", len(tot), "mean n", np.mean([t["n"] for t in tot]))
    for k in keys:
        print(k, np.mean([t[k] for t in tot]))


if __name__ == "__main__":
    main()
