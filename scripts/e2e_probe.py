import sys, time, threading; sys.path.insert(0, ".")
import numpy as np, torch
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.engine import TrialEngine
from flamingpy_b200.noise import sample_batch
lat = build_lattice(13, "both", "periodic"); delta = 0.09
wo = {"method": "blueprint", "delta": delta}
Be = 512; N = lat.N
pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
x = pin((Be, 2 * N), torch.float64); st = pin((Be, N), torch.uint8)
sample_batch(np.random.default_rng(0), N, delta, 0, Be, out_x=x, out_state=st)
def mk():
    e = TrialEngine(lat, delta, 0, wo)
    e.run_injected(x, st, fetch=False); sz = e.sizes()
    bufs = {"hom": pin((Be * e.Qtot,), torch.float64), "weights": pin((Be * e.Qtot,), torch.float64),
            "bits": pin((Be * e.bit_words,), torch.int32).view(np.uint32)}
    for i, ec in enumerate(lat.ec):
        ct = lat.complexes[ec]
        bufs[f"parity{i}"] = pin((Be * ((ct.S + 31) // 32),), torch.int32).view(np.uint32)
        bufs[f"odd_count{i}"] = pin((Be,), torch.int32)
        bufs[f"odd_ids{i}"] = pin((int(sz.total_odd[i]) + 1024,), torch.int32)
        bufs[f"pair_len{i}"] = pin((int(sz.total_pairs[i]) + 1024,), torch.float64)
    return e, bufs
e, bufs = mk()
for rep in range(3):
    t0 = time.perf_counter(); e.run_injected(x, st, fetch=False); t1 = time.perf_counter(); e.fetch_into(bufs); t2 = time.perf_counter()
    nbytes = sum(b.nbytes for b in bufs.values())
    print(f"single lane: run {1e3*(t1-t0):.1f} ms (device {e.timings()['total_ms']:.1f}), fetch {1e3*(t2-t1):.1f} ms = {nbytes/(t2-t1)/1e9:.1f} GB/s")
for lanes in (2, 3):
    es = [mk() for _ in range(lanes)]
    def work(j, n):
        for _ in range(n):
            es[j][0].run_injected(x, st, fetch=False); es[j][0].fetch_into(es[j][1])
    for n in (2, 6):
        ths = [threading.Thread(target=work, args=(j, n)) for j in range(lanes)]
        t0 = time.perf_counter(); [t.start() for t in ths]; [t.join() for t in ths]; dt = time.perf_counter() - t0
        print(f"{lanes} lanes x {n} steps: {lanes*n*Be/dt:.0f} trials/s, {1e3*dt/(lanes*n):.1f} ms/step")
print("--- timeline, 2 lanes")
es = [mk() for _ in range(2)]
log = []
T0 = time.perf_counter()
def work2(j, n):
    for i in range(n):
        a = time.perf_counter(); es[j][0].run_injected(x, st, fetch=False); b = time.perf_counter(); es[j][0].fetch_into(es[j][1]); c = time.perf_counter()
        log.append((j, i, 1e3*(a-T0), 1e3*(b-T0), 1e3*(c-T0), es[j][0].timings()['total_ms'], es[j][0].timings()['graph_ms']))
ths = [threading.Thread(target=work2, args=(j, 4)) for j in range(2)]
[t.start() for t in ths]; [t.join() for t in ths]
for r in sorted(log, key=lambda r: r[2]): print("lane %d step %d: run %.1f -> %.1f (device total %.1f graph %.1f), fetch -> %.1f" % r)
