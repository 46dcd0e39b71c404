"""Per-kernel summary of an `ncu --page raw --csv` export: duration, DRAM bytes read / written, DRAM throughput %, achieved occupancy."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
def col(*names):
    for n in names:
        if n in ix:
            return ix[n]
    return None
c_name, c_dur = col("Kernel Name"), col("gpu__time_duration.sum")
c_rd, c_wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
c_thr, c_occ = col("dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), col("sm__warps_active.avg.pct_of_peak_sustained_active")
units = rows[1]
def f(r, c):
    try:
        return float(r[c].replace(",", ""))
    except Exception:
        return float("nan")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("%-26s dur %10.2f %s  dram read %12.0f %s  write %12.0f %s  dram %% %6.1f  occ %% %5.1f" % (
        r[c_name].split("(")[0].split("::")[-1][:26], f(r, c_dur), units[c_dur], f(r, c_rd), units[c_rd], f(r, c_wr), units[c_wr],
        f(r, c_thr) if c_thr is not None else float("nan"), f(r, c_occ) if c_occ is not None else float("nan")))
