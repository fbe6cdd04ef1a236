"""Aggregate an ncu SASS source page per CUDA source line (developer tool).
usage: ncu_lines.py <report.ncu-rep> <kernel-regex> [lib.so]"""
import csv, re, subprocess, sys, os, tempfile, collections
rep, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else "flamingpy_b200/libfpb.so"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = rows[hi + 1:]
col = {h: i for i, h in enumerate(hdr)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# walk the function of interest
# pick the one function whose instruction count equals the report's
funcs = {}; name = None
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        name = m.group(1); funcs[name] = 0; continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): funcs[name] += 1
target = [f for f, n in funcs.items() if re.search(kern, f) and n == len(body)]
target = target[0] if target else None
print("function", target, file=sys.stderr)
lines = []; cur = None; infn = False
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = (m.group(1) == target) if target else re.search(kern, m.group(1)) is not None
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print("sass instrs in report", len(body), "in disasm", len(lines), file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0])
n = min(len(body), len(lines))
for r, ln in zip(body[:n], lines[:n]):
    a = agg[ln]
    a[0] += int(r[col["Instructions Executed"]] or 0)
    a[1] += int(r[col["Warp Stall Sampling (All Samples)"]] or 0)
    a[2] += 1
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
src = {}
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    if ln and ln[0] not in src:
        for root in ("flamingpy_b200/csrc", "."):
            p = os.path.join(root, ln[0])
            if os.path.exists(p): src[ln[0]] = open(p).read().splitlines(); break
    text = src.get(ln[0], [""] * 100000)[ln[1] - 1].strip() if ln else "?"
    print(f"{ln[0] if ln else '?'}:{ln[1] if ln else 0:4d} inst {100*a[0]/tot_i:5.1f}% stall {100*a[1]/tot_s:5.1f}% sass {a[2]:4d} | {text[:100]}")
