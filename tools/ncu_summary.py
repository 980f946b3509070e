"""Summarise an .ncu-rep (ncu --set full capture brought back in gpurun_out/) as markdown.
Usage (build container, no GPU needed):  python tools/ncu_summary.py gpurun_out/prof.ncu-rep "<title>" "<command>" > profiles/x.md"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main():
    rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n\nCapture command (B200 box, under gpurun):\n\n```\n{cmd}\n```\n")
    print("Numbers below are from the profiler's replay passes (cold cache, serialised): use them for traffic, "
          "pipe utilisation and stall reasons, NOT as bench timings.\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        print(f"## `{name}`\n\n| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {label} | {r[i]} {units[i]} |")
        st = sorted(((float(r[i]), k) for i, k in enumerate(hdr)
                     if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")), reverse=True)
        top = ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
                        for v, k in st[:5])
        print(f"| top stall reasons (warps per issue) | {top} |\n")


if __name__ == "__main__":
    main()
