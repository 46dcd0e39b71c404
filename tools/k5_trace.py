"""GPU box: chunk-level timeline of one preconditioner application (K5), DOTGPU_SOLVE_TRACE=1.
Stamps per (CTA, chunk): 0 posted (TMA issued), 1 vector ready, 2 data + vector seen by the consumers, 3 products done, 4 published, 5 queue slot."""
import os, sys
os.environ["DOTGPU_SOLVE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import dot_b200 as D
from bench import load_workload
wl = load_workload(sys.argv[1] if len(sys.argv) > 1 else "bar1M")
fm = D.Anim(wl["anim"], wl["V"]).fixed_mask()
stp = D.Stepper(wl["V"], wl["T"], wl["epart"], fm, energy=wl["energy"], k=wl["k"], dt=wl["dt"])
print("MS", stp.time_kernels(5, 5))
t = stp.solve_trace().astype(np.int64)
np.save("gpurun_out/k5_chunk_trace_%s.npy" % wl["name"], t)
valid = t[:, :, 0] > 0
t0 = t[:, :, 0][valid].min()
us = lambda a: (a - t0) / 1e3
print("CTAs with chunks:", int(valid.any(axis=1).sum()), "chunks traced per CTA (median):", int(np.median(valid.sum(axis=1))))
post, vr, seen, done, pub = (us(t[:, :, i]) for i in range(5))
m = valid & (t[:, :, 3] > 0) & (t[:, :, 2] > 0)
def stat(name, a):
    a = a[np.isfinite(a)]
    print("%-46s median %7.2f  mean %7.2f  p90 %7.2f us" % (name, np.median(a), a.mean(), np.percentile(a, 90)))
stat("posted -> consumers see data+vector", (seen - post)[m])
stat("posted -> vector ready (gatherers)", (vr - post)[m & (t[:, :, 1] > 0)])
stat("consumers: seen -> products done", (done - seen)[m])
cyc0 = (t[:, :, 6] & 0xffffffff).astype(np.float64)
nrows_ = ((t[:, :, 6] >> 32) & 0xffff)
W_ = ((t[:, :, 6] >> 48) & 0xff)
kind_ = ((t[:, :, 6] >> 56) & 0xff)
cyc3 = t[:, :, 7].astype(np.float64)
mk = m & (t[:, :, 6] > 0)
stat("   warp 0 cycles in the product loop (clock64)", cyc0[mk])
stat("   warp 3 cycles in the product loop (clock64)", cyc3[mk])
for kd in (0, 1):
    for Wv in (1, 2, 4, 8, 16, 32):
        sel = mk & (kind_ == kd) & (W_ == Wv)
        if sel.sum() > 50:
            print("      kind %d W %2d: chunks %6d  rows median %4d  warp0 cycles median %6.0f  warp3 %6.0f  seen->done %.2f us" % (
                kd, Wv, sel.sum(), np.median(nrows_[sel]), np.median(cyc0[sel]), np.median(cyc3[sel]), np.median((done - seen)[sel])))
stat("products done -> published (signaller)", (pub - done)[m & (t[:, :, 4] > 0)])
per = np.diff(done, axis=1)
mm = m[:, 1:] & m[:, :-1]
stat("period between consecutive chunks of a CTA", per[mm])
idle = (seen[:, 1:] - done[:, :-1])
stat("consumer idle between chunks (next seen - prev done)", idle[mm])
pp = np.diff(post, axis=1)
stat("producer period (post to post)", pp[mm])
# a few CTAs in detail
for c in (0, 1, 200):
    if c >= t.shape[0] or not valid[c].any():
        continue
    print("CTA", c)
    for i in range(min(14, int(valid[c].sum()))):
        print("   chunk %2d slot %6d: post %8.2f vready %8.2f seen %8.2f done %8.2f pub %8.2f" % (i, t[c, i, 5], post[c, i], vr[c, i], seen[c, i], done[c, i], pub[c, i]))
