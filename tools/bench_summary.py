import json, sys
for fn in sys.argv[1:]:
    try:
        d = json.load(open(fn))
    except Exception as e:
        print(fn, "ERR", e); continue
    print(fn.split("/")[-1], "fps %.2f e2e %.2f ms/step %.2f iters %d conv %s launches %d solve_ms %.2f refresh_ms %.2f" % (
        d["value"], d["e2e"]["value"], d["ms_per_step"], d["inner_iters"], d["all_frames_converged"], d["gpu_launches"], d["solve_ms_per_step"], d["refresh_ms_per_step"]))
    for k, v in d["kernels"].items():
        print("     %-14s" % k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != "algorithmic_bytes"})
    r = d.get("roofline", {})
    print("     roofline frac %.3f  K5 ms in-frame %.4f isolated %.4f  share %.2f" % (r.get("frac", 0), r.get("avg_launch_ms_in_timed_region", 0), r.get("isolated_ms", 0), r.get("share_of_frame", 0)))
    for key in ("parity", "secondary", "cpu_baseline"):
        if key in d:
            print("     %s:" % key, json.dumps(d[key])[:600])
