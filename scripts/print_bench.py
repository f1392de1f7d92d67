"""Print the headline numbers of a bench.py JSON line (usage: python scripts/print_bench.py file.json)."""
import json, sys
d = json.load(open(sys.argv[1]))
print("n_gpus", d["n_gpus"], "value %.0f" % d["value"], d["unit"], "e2e %.0f" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], "clocks", d["clocks"])
so = d.get("string_ops") or {}
print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in so.items() if not isinstance(v, (dict, str))})
for r in d.get("other_parameter_sets") or []:
    if isinstance(r, dict) and "params" in r:
        print(r["params"], "%.0f KS-PBS/s" % r["ks_pbs_per_s"])
