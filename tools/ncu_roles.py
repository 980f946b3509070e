"""Per-ROLE summary of a profiled warp-specialised kernel: the SASS is split at its USETMAXREG instructions (role entry
points) and samples / executed instructions / stall reasons are summed per section, plus the top instructions per role.
  python tools/ncu_roles.py gpurun_out/x.ncu-rep fused_attn2 [launch_skip]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
res = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(res.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
hdr = rows[starts[skip] + 1]
data = rows[starts[skip] + 2:starts[skip + 1]]
i_src = hdr.index("Source"); i_s = hdr.index("# Samples"); i_e = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
sec, secs = 0, []
cur = {"n": 0, "samples": 0, "exec": 0, "stalls": {}, "top": []}
for k, r in enumerate(data):
    if "USETMAXREG" in r[i_src]:
        secs.append(cur); cur = {"n": 0, "samples": 0, "exec": 0, "stalls": {}, "top": []}
    s, e = int(r[i_s] or 0), int(r[i_e] or 0)
    cur["n"] += 1; cur["samples"] += s; cur["exec"] += e
    st = {}
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v:
            cur["stalls"][h[6:]] = cur["stalls"].get(h[6:], 0) + v; st[h[6:]] = v
    cur["top"].append((s, e, k, r[i_src][:70], st))
secs.append(cur)
tot = sum(c["samples"] for c in secs)
for n, c in enumerate(secs):
    print(f"== section {n}: {c['n']} instr, {100.0*c['samples']/tot:.1f}% samples, {c['exec']/1e6:.1f} M warp instr")
    print("   stalls:", ", ".join(f"{h} {100.0*v/max(c['samples'],1):.0f}%" for h, v in sorted(c["stalls"].items(), key=lambda kv: -kv[1])[:6]))
    for s, e, k, src, st in sorted(c["top"], key=lambda t: -t[0])[:12]:
        print(f"   {100.0*s/tot:5.2f}% {e/1e6:7.2f}M  #{k:<5d} {src}  {max(st, key=st.get) if st else ''}")
