"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump: stall samples per CUDA source line
(all source files of the kernel)."""
import csv, sys, collections, os
rows = list(csv.reader(open(sys.argv[1])))
def I(x):
    try: return int(x)
    except Exception: return 0
agg = collections.OrderedDict()
fname, hdr = "?", None
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r
        isamp, iexec = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= isamp or r[2] != '-': continue
    k = (fname, r[0])
    a = agg.setdefault(k, {"src": r[1], "samp": 0, "exec": 0, "st": collections.Counter()})
    a["samp"] += I(r[isamp]); a["exec"] += I(r[iexec])
    for i in stall_cols: a["st"][hdr[i]] += I(r[i])
tot = sum(a["samp"] for a in agg.values())
print("total samples", tot)
lo = int(sys.argv[3]) if len(sys.argv) > 3 else None
hi = int(sys.argv[4]) if len(sys.argv) > 4 else None
items = agg.items()
if lo is not None:
    items = [(k, a) for k, a in items if lo <= I(k[1]) <= hi]
    items = sorted(items, key=lambda kv: I(kv[0][1]))
else:
    items = sorted(items, key=lambda kv: -kv[1]["samp"])
for k, a in list(items)[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    st = ", ".join("%s %d" % (n[6:], v) for n, v in a["st"].most_common(3) if v)
    print("%-18s %5s %6d %5.1f%% exec %8d | %-64s | %s" % (k[0][:18], k[1], a["samp"], 100.0 * a["samp"] / max(tot, 1), a["exec"], a["src"].strip()[:64], st))
