"""TEST INFRASTRUCTURE ONLY - plain-Python restatement of the reference's Union-Find decoder.

Follows ``flamingpy/decoders/unionfind/algos.py`` (``initialize_cluster_trees`` :58-129, ``union_find``
:132-193, ``obtain_spanning_forest`` :196-226, ``peeling`` / ``trim_forest`` :260-311) and
``uf_classes.py`` (``Support.grow`` :134-142, ``Root`` :81-98, ``Boundary`` :183-232) on the edge list of
``flamingpy_b200.uf_native.uf_graph`` (cubes, then boundary points; one edge per syndrome qubit), with
sets and dicts like the reference, no attempt at speed.  Pinned against the live reference's own
Support tables in ``tests/test_uf.py`` (``tests/golden/uf_*.npz``); then used as the oracle for the
C++ decoder (``csrc/uf.cpp``) on seeded inputs at sizes the golden files do not cover.

Deterministic part = the Support table (which edges are fully grown when no odd cluster is left).
The recovery depends on the spanning forest chosen (the reference: rustworkx + set order); this
restatement peels a depth-first forest rooted at the lowest-numbered boundary point of each tree.
"""

from __future__ import annotations

import numpy as np


def grow_clusters(n_nodes, n_cubes, edge_a, edge_b, parity, erased=None):
    """Support table: uint8[Q], 0 empty / 1 half-grown / 2 grown, after ``union_find`` has finished."""
    Q = len(edge_a)
    status = np.zeros(Q, np.uint8)
    incident = [[] for _ in range(n_nodes)]
    for q in range(Q):
        if edge_a[q] >= 0:
            incident[edge_a[q]].append(q)
            incident[edge_b[q]].append(q)
    parent = list(range(n_nodes))

    def find(u):
        while parent[u] != u:
            parent[u] = parent[parent[u]]
            u = parent[u]
        return u

    members = {u: {u} for u in range(n_nodes)}  # root -> nodes of the cluster
    par = {u: int(parity[u]) if u < n_cubes else 0 for u in range(n_nodes)}
    bnd = {u: u >= n_cubes for u in range(n_nodes)}

    def union(a, b):
        a, b = find(a), find(b)
        if a == b:
            return
        if len(members[a]) < len(members[b]):
            a, b = b, a
        parent[b] = a
        members[a] |= members.pop(b)
        par[a] ^= par.pop(b)
        bnd[a] = bnd[a] or bnd.pop(b)

    if erased is not None:
        for q in range(Q):
            if edge_a[q] >= 0 and erased[q]:
                status[q] = 2  # an erased edge starts fully grown (uf_classes.py:128-131)
                union(edge_a[q], edge_b[q])
    # algos.py:124-126: the first list holds every cluster of odd parity, boundary or not
    odd = {r for r in members if par[r]}
    while odd:
        fusion = []
        changed = False
        for r in odd:
            for u in members[r]:  # Boundary.nodes: nodes with an edge that is not grown; the others change nothing
                for q in incident[u]:
                    if status[q] == 0:
                        status[q] = 1
                        changed = True
                    elif status[q] == 1:
                        status[q] = 2
                        fusion.append(q)
                        changed = True
        for q in fusion:
            union(edge_a[q], edge_b[q])
        odd = {find(r) for r in odd}
        odd = {r for r in odd if par[r] and not bnd[r]}  # algos.py:186-191
        if odd and not changed:  # the reference would loop forever (algos.py:155); nothing can grow any more
            raise ValueError("an odd cluster can neither grow nor reach a boundary point")
    return status


def peel(n_nodes, n_cubes, edge_a, edge_b, parity, status):
    """Recovery uint8[Q] from a spanning forest of the grown edges (algos.py:196-311)."""
    Q = len(edge_a)
    adj = [[] for _ in range(n_nodes)]
    for q in range(Q):
        if edge_a[q] >= 0 and status[q] == 2:
            adj[edge_a[q]].append((q, edge_b[q]))
            adj[edge_b[q]].append((q, edge_a[q]))
    flips = np.zeros(Q, np.uint8)
    seen = [False] * n_nodes
    npar = [int(parity[u]) if u < n_cubes else 0 for u in range(n_nodes)]
    for root in list(range(n_cubes, n_nodes)) + list(range(n_cubes)):
        if seen[root]:
            continue
        order, up = [root], {root: None}
        seen[root] = True
        stack = [root]
        while stack:
            u = stack.pop()
            for q, v in adj[u]:
                if not seen[v]:
                    seen[v] = True
                    up[v] = (q, u)
                    order.append(v)
                    stack.append(v)
        for u in reversed(order[1:]):  # leaves first: an odd subtree sends its parity up through its edge
            if npar[u]:
                q, p = up[u]
                flips[q] ^= 1
                npar[p] ^= 1
        if root < n_cubes and npar[root]:
            raise ValueError("odd tree without a boundary point")
    return flips


def decode(n_nodes, n_cubes, edge_a, edge_b, parity, erased=None):
    status = grow_clusters(n_nodes, n_cubes, edge_a, edge_b, parity, erased)
    return peel(n_nodes, n_cubes, edge_a, edge_b, parity, status), (status == 2).astype(np.uint8)
