#!/usr/bin/env python
"""gpurun_out/parity_r02.jsonl (one record per parity comparison, written by tests/parity_util.py during `pytest -m gpu`)
-> profiles/parity_r02.json: {"summary": ..., "records": [...]}"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_r02.jsonl")
recs = [json.loads(l) for l in open(src) if l.strip()]
summary = {"comparisons": len(recs), "queries_compared": int(sum(r.get("n", 0) for r in recs)),
           "mismatches_listed": int(sum(r.get("n_mismatches", 0) for r in recs)),
           "unexplained": int(sum(r.get("unexplained", 0) for r in recs)),
           "tests": sorted({r.get("test", "?") for r in recs})}
with open(os.path.join(ROOT, "profiles", "parity_r02.json"), "w") as f:
    json.dump({"summary": summary, "records": recs}, f, indent=1)
print(summary)
