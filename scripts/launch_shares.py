"""Dev helper: per-kernel share of device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t = collections.Counter(); n = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[ui], 1.0)
    k = r[ki].split("(")[0][:90]
    t[k] += v; n[k] += 1
tot = sum(t.values())
print("# share of device time per kernel under ncu (cold-cache, serialised: compare SHARES); command: see header of", path.split("/")[-1])
for k, v in t.most_common(14):
    print("%5.1f%%  %4d launches  %10.1f us  %s" % (100 * v / tot, n[k], v, k))
