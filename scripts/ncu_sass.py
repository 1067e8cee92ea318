"""Group the SASS of an `ncu --page source --csv --print-source sass` export by execution count (= code region) and opcode.
usage: ncu_sass.py export.csv [n_regions]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
iS, iE, iW, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("# Samples")
byexec = collections.defaultdict(collections.Counter)
wav = collections.Counter(); smp = collections.Counter(); total = 0
for r in rows[hi + 1:]:
    try: e = int(r[iE])
    except Exception: continue
    t = r[iS].strip().split()
    op = t[1] if t[0].startswith("@") else t[0]
    byexec[e][op.split(".")[0]] += 1
    wav[e] += int(r[iW] or 0); smp[e] += int(r[iSm] or 0); total += e
ts = sum(smp.values())
print("total warp instructions", total, "samples", ts, "shared wavefronts", sum(wav.values()))
reg = sorted(byexec.items(), key=lambda kv: -kv[0] * sum(kv[1].values()))
for e, c in reg[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    n = sum(c.values())
    print("exec %9d x %4d static = %5.1f%% instr, %5.1f%% samples, %6.1f wavefronts/exec | %s" % (
        e, n, 100.0 * e * n / total, 100.0 * smp[e] / ts, wav[e] / max(e, 1), " ".join("%s:%d" % kv for kv in c.most_common(16))))
