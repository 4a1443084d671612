"""Print the headline numbers and the per-kernel-class table of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f %s  ms/step %.3f  e2e %s  launches %s" % (d["value"], d["unit"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), d.get("gpu_launches")))
print("roofline", d.get("roofline"))
for k in d.get("kernels", []):
    print("  %-22s n=%5.0f  ms=%7.3f  share=%5.1f%%  %8.1f TF/s %8.1f GB/s" % (k["kernel"], k["launches_per_step"], k["ms_per_step"], 100 * k["share"], k["tflops"], k["gbs"]))
print("serialised sum: %.3f ms" % sum(k["ms_per_step"] for k in d.get("kernels", [])))
if d.get("generation"): print("generation", d["generation"])
if d.get("cpu_baseline"): print("cpu_baseline", d["cpu_baseline"])
