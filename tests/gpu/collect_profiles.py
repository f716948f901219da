"""Build-box helper: copy the round's evidence from gpurun_out/ (scratch) into profiles/ (tracked) under a round prefix.
usage: python tests/gpu/collect_profiles.py r2"""
import json
import os
import shutil
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ncu_summary  # noqa: E402
from gpu import ncu_report  # noqa: E402,F401


def main(prefix):
    g, p = "gpurun_out", "profiles"
    for src, dst in [("bench_default.json", "bench_default_10k.json"), ("bench_reference.json", "bench_reference_arm.json"),
                     ("launches_bench.csv", "launches_bench_2000reads.csv")]:
        if os.path.exists(os.path.join(g, src)):
            shutil.copy(os.path.join(g, src), os.path.join(p, "%s_%s" % (prefix, dst)))
    if os.path.exists(os.path.join(g, "launches_bench.csv")):
        open(os.path.join(p, "%s_launches_bench_2000reads_summary.csv" % prefix), "w").write(
            "\n".join(ncu_summary.summarise(os.path.join(g, "launches_bench.csv"))) + "\n")
    full = {}
    for tag in ("fill", "edupper", "reseed", "chain", "seed", "extract"):
        rep = os.path.join(g, "prof_%s.ncu-rep" % tag)
        if os.path.exists(rep):
            full[tag] = ncu_report.summarise(rep)
    if full:
        json.dump(full, open(os.path.join(p, "%s_ncu_full_summary.json" % prefix), "w"), indent=1)
        # DRAM traffic of the fill launches of the 2000-read pass (roofline.traffic is scaled from this)
        rd = wr = 0.0
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for d in full.get("fill", []):
            for k, v in d.items():
                if k.startswith("rd["):
                    rd += v * scale[k[3:-1]]
                if k.startswith("wr["):
                    wr += v * scale[k[3:-1]]
        print("fill DRAM bytes in the capture: read %.3e write %.3e" % (rd, wr))
        # roofline.traffic of bench.py is scaled from this file: DRAM bytes per full-matrix-equivalent fill cell of the capture
        import re
        log = os.path.join(g, "ncu_fill.log")
        if os.path.exists(log) and rd + wr > 0:
            m = re.findall(r"'n_fill_cells': ([0-9.]+)", open(log).read())
            if m:
                cells = float(m[-1])
                json.dump({"k_fill": {"dram_bytes_per_unit": (rd + wr) / cells, "unit_count": "n_fill_cells", "dram_read_bytes": rd,
                                      "dram_write_bytes": wr, "units_in_capture": cells,
                                      "source": "ncu --set full --clock-control none, python tests/quick_gpu.py 2000 1 (one lock-step pass of 2000 "
                                                "reads; the %d banded + full-matrix class launches of the pass); profiles/%s_ncu_full_summary.json"
                                                % (len(full["fill"]), prefix)}},
                          open(os.path.join(p, "%s_kernel_traffic.json" % prefix), "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2")
