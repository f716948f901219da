"""Print the headline numbers of the given bench JSON lines."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f.split('/')[-1], "value", round(d["value"], 3), "e2e", round(d["e2e"]["value"], 3), "ms", round(d["ms_per_step"], 1), "cores", d.get("host_cores_busy"))
    r = d["roofline"]
    print("   solo", r.get("all_kernels_ms"))
    print("   lockstep", r.get("lockstep_stage_ms"))
