"""Print the headline numbers of gpurun_out/b*.json bench lines."""
import glob, json, sys
for f in sorted(glob.glob((sys.argv[1] if len(sys.argv) > 1 else "gpurun_out") + "/b*_w*_c*.json")):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), "ms", round(d["ms_per_step"], 1),
              "launches", d["gpu_launches"], "cores_busy", d.get("host_cores_busy"))
        print("  ", {k: round(v, 1) for k, v in d["stage_ms_per_step"].items()})
    except Exception as e:
        print(f, "ERR", e)
