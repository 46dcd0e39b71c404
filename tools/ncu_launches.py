"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / mean device time."""
import csv, sys, collections
fn = sys.argv[1]; skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0; take = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
rows = [r for r in csv.reader(open(fn, errors="replace")) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict(); tot = 0.0; n = 0
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    n += 1
    if n <= skip or n > skip + take: continue
    v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    v = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)  # -> us
    name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v); tot += v
print("launches %d (skip %d) total %.1f us" % (sum(a[0] for a in agg.values()), skip, tot))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s n=%6d total=%10.1f us  mean=%8.2f  max=%8.2f  share=%5.1f%%" % (k[:28], a[0], a[1], a[1] / a[0], a[2], 100 * a[1] / tot))
