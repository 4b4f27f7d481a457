"""Print the interesting parts of a bench.py JSON line:  python profiles/show_bench.py gpurun_out/bench_X.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("N=%d value %.1f f/s (%.3f ms/step)  e2e %.1f f/s (%.3f ms)  launches %d  host issue %.2f ms" % (
    d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["host_issue_ms_per_step"]))
print("repeats", ["%.2f" % x for x in d["timing"]["repeats_ms_per_step"]], "e2e", ["%.2f" % x for x in d["e2e"]["repeats_ms_per_step"]])
r = d["roofline"]
print("roofline: achieved %.1f TF/s frac %.3f share %.3f issued frac %.3f" % (r["achieved"], r["frac"], r["share_of_step"], r["tensor_pipe_frac_issued"]))
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1].get("share_ms_per_step", 0)):
    if "share_ms_per_step" not in v:
        print("  %-30s %.3f ms" % (k, v["ms_per_step"]))
        continue
    extra = ""
    if "hbm_frac" in v:
        extra += " hbm %.3f" % v["hbm_frac"]
    if "TFLOPps" in v:
        extra += " %.0f TF/s" % v["TFLOPps"]
    print("  %-30s %6.3f ms  n=%5.1f avg %7.1f us%s" % (k, v["share_ms_per_step"], v["calls_per_step"], v["avg_us"], extra))
    for a, b in v.items():
        if a.startswith("isolated"):
            print("        %-28s %s" % (a, {x: round(y, 3) for x, y in b.items()}))
ex = d.get("extras") or {}
for k, v in ex.items():
    print("extra", k, json.dumps({a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if not isinstance(b, dict)}))
    for a, b in v.items():
        if isinstance(b, dict):
            print("      ", a, {x: (round(y, 3) if isinstance(y, float) else y) for x, y in b.items()})
print("clocks", d.get("clocks"))
if d.get("cpu_baseline"):
    print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
