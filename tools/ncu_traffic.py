"""Extract per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of every profiled kernel from an
ncu --set full report and write profiles/r2_traffic.json (read by bench.py for roofline.traffic).
Usage (build container): python tools/ncu_traffic.py gpurun_out/s5_prof.ncu-rep"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    acc = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("hmvit::", "").split("<")[0]
        b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        a = acc.setdefault(name, {"launches": 0, "bytes": 0.0, "dur": 0.0})
        a["launches"] += 1
        a["bytes"] += b
        a["dur"] += float(r[it])
    res = {k: {"dram_bytes_per_launch": v["bytes"] / v["launches"], "launches_profiled": v["launches"],
               "profiled_duration_" + units[it]: v["dur"] / v["launches"]} for k, v in acc.items()}
    res["_source"] = {"report": os.path.basename(rep),
                      "command": sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none --import-source on ... python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train"}
    json.dump(res, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
