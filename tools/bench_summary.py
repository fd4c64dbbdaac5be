"""prints the numbers of a bench.py JSON line that matter while iterating"""
import json, sys
for ln in open(sys.argv[1]):
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    c = d.get("config", {})
    print("value %.2f %s  ms/step %.2f  e2e %s  launches %s  clocks %s" % (d["value"], d["unit"], d.get("ms_per_step", 0), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), d.get("clocks")))
    print("  stages", [round(x, 2) for x in c.get("stage_ms_last_step", [])], "iters", c.get("pcg_iters_last_step"), "resid", c.get("pcg_residual_last_step"))
    print("  iter ms", c.get("pcg_iteration_ms_kernels"), "iter hbm frac", c.get("pcg_iteration_hbm_frac"), "dense", c.get("step_dense_model_frac"), "touched", c.get("step_hbm_frac_touched"))
    for k, v in (c.get("kernels") or {}).items():
        print("   %-40s n=%-4d avg %.4f ms %s" % (k, v["launches"], v["avg_ms"], ("%.0f GB/s" % v["gbs"]) if "gbs" in v else ""))
    print("  roofline", d.get("roofline"))
    if c.get("parity_checked"):
        print("  parity", c["parity_checked"])
    if d.get("cpu_baseline"):
        print("  cpu", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("sample"))
