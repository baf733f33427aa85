"""Per-phase stall breakdown of one kernel from an `ncu --page source --print-source cuda,sass --csv` dump.

    python tools/ncu_phase_stalls.py dump.csv FILE:LO-HI:name [FILE:LO-HI:name ...]

Source-line rows only (the SASS rows repeat them).  A phase = a range of source lines of one file (inlined helpers
are attributed to their own file/lines)."""
import csv
import sys
import collections

rows = list(csv.reader(open(sys.argv[1])))
phases = []
for a in sys.argv[2:]:
    f, rng, name = a.split(":")
    lo, hi = rng.split("-")
    phases.append((f, int(lo), int(hi), name))
hdr = None
cur = None
stall_cols = {}
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
        continue
    if len(r) >= 8 and r[0].isdigit() and hdr:
        line = int(r[0])
        name = None
        for f, lo, hi, n in phases:
            if f == cur and lo <= line <= hi:
                name = n
                break
        if name is None:
            name = "other:" + cur
        try:
            agg[name]["samples"] += int(r[6]); agg[name]["inst"] += int(r[7])
            for i, h in stall_cols.items():
                agg[name][h] += int(r[i] or 0)
        except ValueError:
            pass
ts = sum(v["samples"] for v in agg.values()) or 1
ti = sum(v["inst"] for v in agg.values()) or 1
for name, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    st = sorted(((h[6:], c) for h, c in v.items() if h.startswith("stall_") and c), key=lambda x: -x[1])[:5]
    print("%-26s samp %5.1f%% inst %5.1f%%  %s" % (name, 100 * v["samples"] / ts, 100 * v["inst"] / ti,
                                                  " ".join("%s:%.0f%%" % (h, 100 * c / max(1, v["samples"])) for h, c in st)))
