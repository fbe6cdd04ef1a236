import sys; sys.path.insert(0,'.')
import numpy as np
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.engine import TrialEngine
from oracle import restate
lat = build_lattice(5, "primal", "open"); ct = lat.complexes["primal"]
delta, ps = 0.09, 0.25
wopts = {"method": "blueprint", "delta": delta}
eng = TrialEngine(lat, delta, ps, wopts, mode="fresh")
T = 4
batch = eng.run_rng(11, 0, T, want_draws=True)
bad = 0; n = 0
for t in range(T):
    res = restate.trial_from_draws(lat, batch.state[t], batch.x[t], delta, wopts, want_graph=True, want_paths=True)
    md = res.per_complex["primal"]["graph"]; K = len(md.odd)
    for u in range(K):
        for v in list(range(u + 1, K)) + [-1]:
            path = md.pair_paths[(u, v)] if v >= 0 else md.bnd_paths[u]
            exp = np.zeros(ct.Q, np.uint8)
            for q in path: exp[q] ^= 1
            fl = eng.paths_toggle([t], [0], [u], [v])[t]
            n += 1
            if not np.array_equal(fl, exp):
                bad += 1
                w = res.per_complex["primal"]["weights"]
                if bad < 6:
                    print("len expected", w[np.nonzero(exp)[0]].sum(), "got", w[np.nonzero(fl)[0]].sum())
                    print("MISMATCH t", t, "u", u, "v", v, "cubes", md.odd[u], md.odd[v] if v >= 0 else None,
                          "expected", np.nonzero(exp)[0], "got", np.nonzero(fl)[0], "side", md.bnd_side[u] if v < 0 else None)
print("queries", n, "bad", bad)
