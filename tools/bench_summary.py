import json
import sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable:", e)
        continue
    print(path, "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "kernel_ms", {k: round(v, 2) for k, v in d.get("kernel_ms", {}).items()},
          "solved", d["config"].get("solved_fraction"), "iters", d["config"].get("mean_ipm_iters"), "clk", d.get("clocks", {}).get("sm_mhz"))
