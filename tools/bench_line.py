"""Condense a bench.py JSON line (stdin) to one short line: python bench.py ... | python tools/bench_line.py [label]"""
import json
import sys

d = json.loads(sys.stdin.read())
g = d.get("roofline", {}).get("all_groups", {})
e = d.get("e2e", {})
print(sys.argv[1] if len(sys.argv) > 1 else "", "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "e2e", round(e.get("value", 0), 1),
      [round(x) for x in e.get("runs", [])], {k[:14]: round(v["ms_per_step"], 2) for k, v in g.items()})
