#!/usr/bin/env python
"""ncu traffic CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) -> profiles/ncu_traffic.json:
mean DRAM bytes per launch of every kernel bench.py reports a roofline for, plus the fwd+bwd patch-attention pair.
usage: tools/traffic_json.py gpurun_out/traffic.csv "source note" > profiles/ncu_traffic.json"""
import csv, json, re, sys
from collections import defaultdict

path, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lines = [l for l in open(path) if l.startswith('"')]
per = defaultdict(lambda: defaultdict(float))        # launch id -> metric -> value (bytes / ns)
name_of = {}
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
    per[r["ID"]][r["Metric Name"]] = v
    name_of[r["ID"]] = r["Kernel Name"]

def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "")
    return re.sub(r"<.*", "", n)

agg = defaultdict(lambda: [0, 0.0, 0.0])
for i, m in per.items():
    k = short(name_of[i])
    a = agg[k]
    a[0] += 1; a[1] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0); a[2] += m.get("gpu__time_duration.sum", 0)
out = {}
for k, (n, b, ns) in agg.items():
    out[k] = {"dram_bytes_per_launch": b / n, "launches": n, "ncu_us_per_launch": ns / n / 1e3, "source": note}
pf, pb = agg.get("paf::k_pattn_fwd"), agg.get("paf::k_pattn_bwd")
if pf and pb:
    out["patch_attention(fwd+bwd)"] = {"dram_bytes_per_launch": (pf[1] + pb[1]) / (pf[0] + pb[0]), "launches": pf[0] + pb[0], "source": note}
json.dump(out, sys.stdout, indent=1)
