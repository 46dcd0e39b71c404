"""Summarise an `ncu --page source --csv` export (SASS view): stall breakdown + hottest instructions of the first kernel."""
import csv, gzip, io, sys
fn = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
f = io.TextIOWrapper(gzip.open(fn)) if fn.endswith(".gz") else open(fn)
rows = list(csv.reader(f))
kern = []  # split into kernels at 'Kernel Name' rows
for r in rows:
    if r and r[0] == "Kernel Name":
        kern.append({"name": r[1], "hdr": None, "body": []})
    elif r and r[0] == "Address":
        kern[-1]["hdr"] = r
    elif kern and kern[-1]["hdr"] and len(r) == len(kern[-1]["hdr"]):
        kern[-1]["body"].append(r)
print("kernels:", [k["name"].split("(")[0][-30:] for k in kern])
k = kern[which]; hdr = k["hdr"]; body = k["body"]
def iv(x):
    try: return int(x)
    except Exception: return 0
tot = sum(iv(r[2]) for r in body)
print(k["name"].split("(")[0], "total samples", tot, "ninstr", len(body))
sc = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
agg = {hdr[i]: sum(iv(r[i]) for r in body) for i in sc}
print([(a, b) for a, b in sorted(agg.items(), key=lambda kv: -kv[1])[:8]])
top = sorted(range(len(body)), key=lambda i: -iv(body[i][2]))[:ntop]
for i in sorted(top):
    r = body[i]
    st = {hdr[j][6:]: iv(r[j]) for j in sc if iv(r[j]) > 0}
    print(str(i).rjust(5), r[1].strip()[:64].ljust(64), r[2].rjust(6), st)
