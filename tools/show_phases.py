"""Print value / ms_per_step / per-kernel phase times of bench.py JSON lines (files given on the command line)."""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"], 1), round(d["ms_per_step"], 4),
          {a: round(b, 4) for a, b in (d.get("phases_ms") or {}).items()})
