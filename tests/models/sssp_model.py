"""Lane-level Python model of the warp-per-source bucketed shortest-path search
in flamingpy_b200/csrc (k_paths).  Test infrastructure: it lets the algorithm
(pass A = compact + collect, pass B = relax, empty-bucket skipping) be checked
against plain Dijkstra on CPU before spending GPU time."""

import numpy as np

INF = np.inf
CURCAP = 64


def sssp_bucketed(ct, wslot, seeds, step):
    """seeds: list of (cube, dist0). Returns (dist[S], stats)."""
    S = ct.S
    dist = np.full(S, INF)
    flag = np.zeros(S, np.uint8)
    L = []
    for c, d0 in seeds:
        dist[c] = d0
        flag[c] = 1
        L.append(c)
    t_lo, t_hi = 0.0, step
    passes = relax_batches = 0
    while True:
        # ---- pass A
        out, cur, more = [], [], False
        minv = INF
        for base in range(0, len(L), 32):
            chunk = L[base:base + 32]
            for v in chunk:
                dv = dist[v]
                if dv < t_lo:
                    continue  # settled in an earlier bucket
                out.append(v)
                if dv < t_hi:
                    if flag[v] & 1:
                        if len(cur) < CURCAP:
                            cur.append(v)
                            flag[v] &= ~np.uint8(1)
                        else:
                            more = True
                else:
                    minv = min(minv, dv)
        L = out
        passes += 1
        if not cur:
            if minv == INF:
                break
            t_lo = t_hi
            t_hi = (np.floor(minv / step) + 1.0) * step
            if not t_hi > minv:  # guard against rounding at a bucket edge
                t_hi = minv + step
            continue
        # ---- pass B
        again = more
        min_b = INF
        for cb in range(0, len(cur), 32):
            batch = cur[cb:cb + 32]
            dvs = [dist[v] for v in batch]  # read at batch start, like the kernel
            relax_batches += 1
            for d in range(6):
                for v, dv in zip(batch, dvs):
                    kind = (int(ct.ninfo[v]) >> (2 * d)) & 3
                    if kind >= 2:
                        continue
                    u = v + int(ct.d_off[d, kind])
                    w = wslot[int(ct.sbase[v]) + int(ct.s_off[d, kind])]
                    nd = dv + w
                    old = dist[u]
                    if nd < old:
                        dist[u] = nd
                        flag[u] = 1 | (d << 1)
                        if old == INF:
                            L.append(u)
                        if nd < t_hi:
                            again = True
                        else:
                            min_b = min(min_b, nd)
        if not again:
            # bucket finished without a confirming pass
            mn = min(minv, min_b)
            if mn == INF:
                break
            t_lo = t_hi
            t_hi = (np.floor(mn / step) + 1.0) * step
            if not t_hi > mn:
                t_hi = mn + step
    return dist, {"passes": passes, "relax_batches": relax_batches}
