"""Print the SASS of one execution-count region of an `ncu --page source --csv --print-source sass` export in address order,
with samples and the dominant stall reasons.  usage: ncu_region.py export.csv exec_count"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
want = int(sys.argv[2]); tot = 0
for r in rows[hi + 1:]:
    try: e = int(r[iE])
    except Exception: continue
    if e != want: continue
    st = sorted(((int(r[i] or 0), n) for i, n in stall), reverse=True)[:2]
    tot += int(r[iSm] or 0)
    print("%5s %-62s %s" % (r[iSm], r[iS][:62], " ".join("%s:%d" % (n, v) for v, n in st if v)))
print("samples in region", tot)
