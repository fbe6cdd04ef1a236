"""Lane-level Python model of the warp-per-source bucketed shortest-path search
in flamingpy_b200/csrc (warp_search in fpb_kernels.cuh).  Test infrastructure:
it lets the algorithm (one bucket byte per cube, dense scan for the current
bucket, batched relaxation with store-then-verify conflict resolution, parking
of cubes beyond the ring horizon) be checked against plain Dijkstra on CPU
before spending GPU time."""

import math

import numpy as np

INF = np.inf
NB = 32  # ring slots (bucket index mod NB is stored in one byte per cube)
NONE = 0xFF
CURCAP = 256  # staging list capacity


def sssp_bucketed(ct, wslot, seeds, step, nb=NB, curcap=CURCAP):
    """seeds: list of (cube, dist0).  Returns (dist[S], stats)."""
    S = ct.S
    inv = 1.0 / step
    dist = np.full(S, INF)
    bkt = np.full(S, NONE, np.uint8)  # ring slot of the cube's pending entry
    present = 0  # superset of the non-empty ring slots
    stats = {"steps": 0, "scans": 0, "batches": 0, "repark": 0, "pushes": 0, "lanes": 0}

    def bucket_of(x):
        return int(math.floor(x * inv)) if x < INF else 2 ** 31 - 1

    cur = 0  # absolute index of the bucket being processed

    def push(u, bi):
        nonlocal present
        bi = min(bi, cur + nb - 1)  # beyond the horizon: park in the farthest slot
        r = bi % nb
        bkt[u] = r
        present |= 1 << r
        stats["pushes"] += 1

    def relax_batch(batch):
        """batch: list of (v, dv).  All loads first, stores in direction order, then
        verification rounds; pushes use the verified distance."""
        stats["batches"] += 1
        stats["lanes"] += len(batch)
        cand = []  # (d, lane, u, nd)
        for lane, (v, dv) in enumerate(batch):
            for d in range(6):
                kind = (int(ct.ninfo[v]) >> (2 * d)) & 3
                if kind >= 2:
                    continue
                u = v + int(ct.d_off[d, kind])
                w = wslot[int(ct.sbase[v]) + int(ct.s_off[d, kind])]
                nd = dv + w
                if nd < dist[u]:  # all `old` values are read before any store
                    cand.append((d, lane, u, nd))
        cand.sort(key=lambda c: (c[0], c[1]))
        for d, lane, u, nd in cand:  # stores in direction order: last writer wins
            dist[u] = nd
        while True:  # verify with hoisted loads, re-store in direction order
            seen = {u: dist[u] for _, _, u, _ in cand}
            redo = [(d, lane, u, nd) for d, lane, u, nd in cand if seen[u] > nd]
            if not redo:
                break
            for d, lane, u, nd in redo:
                dist[u] = nd
        for d, lane, u, nd in cand:
            push(u, bucket_of(dist[u]))

    for c, d0 in seeds:
        dist[c] = d0
    for c, d0 in seeds:
        push(c, bucket_of(d0))

    while present:
        # next slot in ring order from the head
        k = next(k for k in range(nb) if present >> ((cur + k) % nb) & 1)
        cur += k
        r = cur % nb
        stats["steps"] += 1
        while present >> r & 1:
            present &= ~(1 << r)
            stats["scans"] += 1
            start = 0
            while start < S:
                # dense scan for cubes filed under slot r, at most `curcap` per chunk
                staged = []
                v = start
                while v < S and len(staged) < curcap:
                    if bkt[v] == r:
                        staged.append(v)
                        bkt[v] = NONE
                    v += 1
                start = v
                for base in range(0, len(staged), 32):
                    batch = []
                    for v2 in staged[base:base + 32]:
                        bt = bucket_of(dist[v2])
                        if bt == cur:
                            batch.append((v2, dist[v2]))
                        elif bt > cur:
                            stats["repark"] += 1
                            push(v2, bt)  # a parked cube: move it towards its real bucket
                    if batch:
                        relax_batch(batch)
    return dist, stats
