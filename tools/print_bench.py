"""One-screen digest of a bench.py JSON line:  python tools/print_bench.py gpurun_out/bench.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.2f %s | e2e %.2f | e2e_detected %s | n_gpus %d | ms/step %.2f" % (
    d["value"], d["unit"], d["e2e"]["value"], (d.get("e2e_detected") or {}).get("value"), d["n_gpus"], d["ms_per_step"]))
r = d.get("roofline") or {}
print("roofline %s: %.1f %s = %.3f of peak (%s)" % (r.get("kernel"), r.get("achieved", 0), r.get("unit"), r.get("frac", 0), r.get("peak_source")))
for k, v in (r.get("per_kernel") or {}).items():
    if v:
        print("   %-20s %.1f TF/s frac %.3f pipe %.3f share %.3f" % (k, v["achieved"], v["frac"], v["tensor_pipe_frac"], v["share_of_step"]))
print("clocks", d.get("clocks"))
print("stages", json.dumps(d.get("stages"))[:900])
for k, v in list((d.get("kernels") or {}).items())[:12]:
    print("   %-28s share %.3f  %s" % (k, v["share"], {a: round(b, 3) for a, b in v.items() if a.startswith(("frac", "tensor_pipe", "algorithmic"))}))
print("cpu_baseline", d.get("cpu_baseline"))
