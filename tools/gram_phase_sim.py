"""Host-side estimate of how much of the work-list Gram's panel traffic can be shared through L2.

Replays the schedule of csrc/gram_wl.cu (tiles laid end to end by cost over the CTAs) on a common clock and counts, per
time step, how many DISTINCT (panel, row chunk) pairs the CTAs request within a residency window, with and without the
phase-aligned cyclic walk.  Model only: every CTA advances at exactly its tile's cost rate."""
import sys
import numpy as np


def schedule(m, n, ncta, BK, phase, load_pct=70):
    nt = (m + 127) // 128
    tiles = []
    for tj in range(nt):
        for ti in range(tj + 1):
            diag = ti == tj
            mma = 18 / 32 if diag else 1.0
            load = 0.01 * load_pct * (128 + (0 if diag else 128)) / 256
            tiles.append((ti, tj, max(mma, load)))
    total = sum(c for _, _, c in tiles) * n
    L = total / ncta
    ctas = [[] for _ in range(ncta)]
    U = 0.0
    for ti, tj, cost in tiles:
        span = cost * n
        b_lo, b_hi = int(U // L), min(int((U + span) // L), ncta - 1)

        def boundary(b):
            if b <= b_lo:
                return 0
            if b > b_hi:
                return n
            r = int(round((b * L - U) / cost / BK)) * BK
            if r < 16 * BK:
                r = 0
            if n - r < 16 * BK:
                r = n
            return min(max(r, 0), n)
        for b in range(b_lo, b_hi + 1):
            r0, r1 = boundary(b), boundary(b + 1)
            if r1 <= r0:
                continue
            nch = (r1 - r0 + BK - 1) // BK
            c0 = 0
            if phase:
                ln = max(1, int(round(L / cost / BK)))
                c0 = ((ln - (r0 // BK) % ln) % ln) % nch
            ctas[b].append((ti, tj, cost, r0 // BK, nch, c0))
        U += span
    return ctas, L


def simulate(m, n, ncta=148, BK=32, phase=1, window=64, samples=400):
    ctas, L = schedule(m, n, ncta, BK, phase)
    T = L                                    # every CTA is busy for L cost units
    tot = uniq = 0
    for t in np.linspace(0.001 * T, 0.999 * T, samples):
        reads = set()
        cnt = 0
        for items in ctas:
            tt = t
            for ti, tj, cost, a, nch, c0 in items:
                dur = cost * nch * BK
                if tt < dur:
                    c = (c0 + int(tt / (cost * BK))) % nch
                    row = (a + c) // window            # residency window: chunks this close count as shared
                    reads.add((ti, row)); cnt += 1
                    if tj != ti:
                        reads.add((tj, row)); cnt += 1
                    break
                tt -= dur
        tot += cnt
        uniq += len(reads)
    return uniq / tot


if __name__ == "__main__":
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 896
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096000
    for ph in (0, 1):
        for w in (16, 64, 256):
            print(f"m={m} n={n} phase={ph} window={w} chunks: DRAM share of panel requests ~ {simulate(m, n, phase=ph, window=w):.3f}")
