"""one-screen summary of a bench.py JSON line"""
import json
import sys

for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if "dump_overlap" in d:
        print(json.dumps(d, indent=1))
        continue
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "steps", "n_gpus", "gpu_launches")})
    print(" config:", d["config"].get("workload", "")[:60], d["config"].get("kernel"), d["config"].get("transport"),
          "fallback" if d["config"].get("fallback") else "")
    print(" roofline: frac144=%.3f frac96=%s traffic/min=%s dram=%s" % (
        r.get("frac", 0), r.get("frac_on_minimum"), r.get("traffic_over_minimum"), r.get("dram_gbs_from_traffic")))
    print(" e2e:", e.get("value"), e.get("seconds"), e.get("matches_selfcheck"), e.get("skipped"))
    print(" selfcheck:", d.get("selfcheck"))
    print(" cpu:", (d.get("cpu_baseline") or {}).get("value"), " clocks:", d.get("clocks"))
