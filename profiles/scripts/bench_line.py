"""prints value / ms / per-kernel times of one bench.py JSON line read from stdin (A/B helper)"""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
ks = [(k["kernel"], round(k["avg_ms"], 3)) for k in d.get("roofline", {}).get("kernels", [])]
print(tag, "value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.4g" % (d.get("e2e") or {}).get("value", 0), ks)
