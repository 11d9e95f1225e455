"""print the headline numbers and the per-kernel table of a bench.py JSON line"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 4), "| e2e", round(d["e2e"]["value"]),
      "ms", round(d["e2e"]["ms_per_step"], 3), "| launches", d.get("gpu_launches"))
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), "| cpu", d.get("cpu_baseline", {}).get("value"))
for k in d.get("kernels", []):
    print("  %-44s %6.1f x %8.4f ms %s" % (k["name"], k["launches_per_step"], k["ms_per_step"],
                                          ("%.3f of hbm" % k["frac_of_hbm_peak"]) if "frac_of_hbm_peak" in k else ""))
