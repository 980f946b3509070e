"""Static evidence of the built library (no GPU needed): per kernel, registers / spills / shared memory from
`cuobjdump -res-usage` and the count of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path
(UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor copies,
UBLKCP = cp.async.bulk, SYNCS = mbarrier, HMMA = mma.sync) -- B200_PROFILING.md's mnemonic list.

    python tools/sass_summary.py > profiles/r1_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hm-vit_b200", "libhmvit_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS",
             "LDSM", "MUFU.EX2", "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in out:
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"\(.*$", "", n)
        short.append(n.replace("hmvit::", ""))
    return dict(zip(names, short))


def main():
    cu = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
    res = subprocess.run([cu, "-res-usage", LIB], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in ln:
            usage[cur] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", ln)}
            cur = None
    sass = subprocess.run([cu, "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.defaultdict(collections.Counter), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + ".") or (mn == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                    counts[cur][mn] += 1
    names = demangle(sorted(usage))
    print("# Static SASS / resource summary of hm-vit_b200/libhmvit_b200.so (sm_100a, final build of round 2)\n")
    print("`python tools/sass_summary.py` (cuobjdump -res-usage / -sass; no GPU).  UTCHMMA = `tcgen05.mma`, LDTM / STTM = "
          "`tcgen05.ld` / `tcgen05.st`, UTMALDG / UTMASTG = TMA tensor copies, UBLKCP = `cp.async.bulk`, SYNCS = mbarrier, "
          "HMMA = `mma.sync`, STL / LDL = local-memory (spill) traffic.\n")
    cols = ["REG", "STACK", "SHARED"] + MNEMONICS
    print("| kernel | SASS instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for mangled in sorted(usage, key=lambda k: names[k]):
        u, c = usage[mangled], counts.get(mangled, {})
        row = [str(u.get(k, 0)) for k in ("REG", "STACK", "SHARED")] + [str(c.get(mn, 0) or "") for mn in MNEMONICS]
        print(f"| `{names[mangled]}` | {c.get('_total', 0)} | " + " | ".join(row) + " |")


if __name__ == "__main__":
    sys.exit(main())
