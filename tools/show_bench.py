import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 4), d.get("step_ms"), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["clocks"])
