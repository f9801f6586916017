"""Condense bench.py's JSON line (a file given as the argument, else stdin) to the numbers watched while tuning."""
import json
import sys

for line in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin):
    if line.startswith("{"):
        d = json.loads(line)
        print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3),
              {k: round(v["ms_per_step"], 3) for k, v in d.get("stages", {}).items()},
              "cpu", d.get("cpu_baseline", {}).get("value"))
