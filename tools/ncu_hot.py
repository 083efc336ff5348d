"""top stall sites of a kernel from an ncu report (dev tool): python tools/ncu_hot.py report.ncu-rep [kernel-substring] [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2; data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            data.append(rows[j]); j += 1
        if pat in name:
            ci = hdr.index("Warp Stall Sampling (All Samples)"); ie = hdr.index("Instructions Executed")
            tot = sum(int(r[ci]) for r in data if len(r) > ci and r[ci].isdigit())
            print("==", name, "samples", tot, "sass lines", len(data))
            idx = sorted(range(len(data)), key=lambda k: -(int(data[k][ci]) if data[k][ci].isdigit() else 0))[:topn]
            for k in sorted(idx):
                r = data[k]
                # the instruction a stall is waiting FOR is usually a few lines above: print 1 line of context
                print(f"{int(r[ci]):7d} {100 * int(r[ci]) / max(tot, 1):5.1f}%  exec {r[ie]:>9s}  [{k:5d}] {r[1].strip()[:110]}")
            break
        i = j
    else:
        i += 1
