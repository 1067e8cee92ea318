"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export per CUDA source line / function."""
import csv, re, sys
path, srcfile = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if r and r[0] == "Line No")
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows:
    if len(r) > iI and r[0].isdigit():
        num = lambda v: int(v) if v.strip().isdigit() else 0
        data.append((int(r[0]), r[1].strip(), num(r[iI]), num(r[iS])))
tot, tots = sum(d[2] for d in data), sum(d[3] for d in data)
print("total warp-instr", tot, "samples", tots)
src = open(srcfile).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:__device__|__global__|template).*?\b(\w+)\(", l)
    if l.startswith("__device__") or l.startswith("__global__"):
        m = re.search(r"(\w+)\(", l.split("__forceinline__")[-1])
        if m: funcs.append((i, m.group(1)))
def fn(line):
    name = "?"
    for s, n in funcs:
        if line >= s: name = n
    return name
agg = {}
for ln, s, ins, sm in data:
    a = agg.setdefault(fn(ln), [0, 0]); a[0] += ins; a[1] += sm
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-22s instr %5.1f%%  samples %5.1f%%" % (k, 100 * v[0] / tot, 100 * v[1] / tots))
print("top lines by samples")
for ln, s, ins, sm in sorted(data, key=lambda d: -d[3])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print("%4d %5.1f%% smp %5.1f%% ins  %s" % (ln, 100 * sm / tots, 100 * ins / tot, s[:120]))
