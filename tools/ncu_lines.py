"""Per-source-line stall attribution of one profiled launch (no GPU needed).

ncu's CSV source page is SASS-level without line numbers; nvdisasm -g of the same cubin carries
`//## File "...", line N` markers.  Both list the kernel's instructions in the same order, so they are
joined by position.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep group_attn_kernel [launch_skip] [top] [mangled section pattern]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "hm-vit_b200", "libhmvit_b200.so")


def sass_lines(kernel_pat):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    out, cur_line, cur_file, active = [], None, None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            active = re.search(kernel_pat, m.group(1)) is not None and not out
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_file, cur_line = os.path.basename(m.group(1)), int(m.group(2))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            out.append((cur_file, cur_line, ln.split("*/", 1)[1].strip().rstrip(";")))
    return out


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    res = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(res.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    k = int(skip)
    hdr = rows[starts[k] + 1]
    data = rows[starts[k] + 2:starts[k + 1]]
    sl = sass_lines(sys.argv[5] if len(sys.argv) > 5 else kern)   # optional: mangled-name pattern of the instance
    if len(sl) != len(data):
        print(f"warning: {len(sl)} disassembled instructions vs {len(data)} profiled rows; joining by position anyway")
    i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    tot_s = tot_e = 0
    for k, r in enumerate(data):
        f, l, _ = sl[k] if k < len(sl) else ("?", 0, "")
        key = (f, l)
        a = agg.setdefault(key, {"samples": 0, "exec": 0, "stalls": {}})
        s, e = int(r[i_s] or 0), int(r[i_e] or 0)
        a["samples"] += s
        a["exec"] += e
        tot_s += s
        tot_e += e
        for i, h in stall_cols:
            v = int(r[i] or 0)
            if v:
                a["stalls"][h] = a["stalls"].get(h, 0) + v
    print(f"kernel {kern} (launch skip {skip}): {tot_s} samples, {tot_e} warp instructions")
    srcs = {}
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        if f not in srcs:
            p = os.path.join(ROOT, "hm-vit_b200", "csrc", f or "")
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:90] if l and l <= len(srcs[f]) else ""
        st = ", ".join(f"{h[6:]} {v}" for h, v in sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3])
        print(f"{100.0 * a['samples'] / max(tot_s, 1):5.1f}% smp {100.0 * a['exec'] / max(tot_e, 1):5.1f}% inst  {f}:{l:<4d} {text}\n        [{st}]")


if __name__ == "__main__":
    main()
