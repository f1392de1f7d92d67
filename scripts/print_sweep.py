"""Print the rows of a `bench.py --param-sweep` JSON line (usage: python scripts/print_sweep.py file.json)."""
import json, sys
for r in json.load(open(sys.argv[1]))["sets"]:
    print(r["params"], r["N"], r["k"], r["pbs_level"], r["batch"], "ms %.1f ks %.2f pbs %.1f" % (r["ms"], r["keyswitch_ms"], r["pbs_ms"]),
          "gpu %.0f/s tflops %.2f" % (r["ks_pbs_per_s"], r["pbs_tflops"]), "cpu %.1f/s" % r.get("cpu_port_ks_pbs_per_s", 0))
