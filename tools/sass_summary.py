"""per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import collections, re, subprocess, sys, os
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "robot-3dlotus_b200", "liblotus.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "HMMA", "LDGSTS", "SYNCS", "LDS", "STS", "LD.E", "ST.E"]
cur, counts, regs = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    mm = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if mm:
        op = mm.group(1)
        for k in keys:
            if op.startswith(k):
                counts[cur][k] += 1
                break
        counts[cur]["_n"] += 1
print("# Round 2 — SASS evidence per kernel (`cuobjdump -sass robot-3dlotus_b200/liblotus.so`, sm_100a)\n")
print("`UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG` / `UTMASTG` / `UTMAREDG` = TMA load / store / reduce-add, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier, "
      "`LDGSTS` = cp.async, `HMMA` = legacy mma.sync.  `LDS` / `STS` are shared-space accesses; the generic `LD.E` / `ST.E` columns are GLOBAL accesses "
      "(round 1's tensor-core kernels reached shared memory through generic pointers; that is gone).\n")
print("| kernel | SASS lines | " + " | ".join(keys) + " |")
print("|---|---:|" + "---:|" * len(keys))
for k, c in counts.items():
    if any(c[x] for x in ("UTCHMMA", "LDTM", "UTMALDG", "HMMA", "UTMASTG")):
        print(f"| `{k}` | {c['_n']} | " + " | ".join(str(c[x]) for x in keys) + " |")
