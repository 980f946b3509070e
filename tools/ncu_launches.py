"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file ...) per kernel as markdown.
Usage (build container): python tools/ncu_launches.py gpurun_out/launches.csv "<title>" "<command>" > profiles/x.md"""
import csv
import sys


def main():
    path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = {}
    unit = "ns"
    for r in rows[1:]:
        name = r[ik].split("(")[0][:70]
        a = acc.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
        unit = r[iu]
    tot = sum(v[1] for v in acc.values())
    print(f"# {title}\n\nCommand (on the B200 box, under gpurun):\n\n```\n{cmd}\n```\n")
    print(f"Per-launch times are cold-cache and serialised (compare SHARES, not absolutes). Unit: {unit}. "
          f"First {sum(v[0] for v in acc.values())} launches of the process.\n")
    print("| kernel | launches | total | share |\n|---|---:|---:|---:|")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.0f} | {100 * v[1] / tot:.1f}% |")


if __name__ == "__main__":
    main()
