#!/usr/bin/env python
"""Non-GEMM network kernels of the LAST forward in `ncu --csv --metrics gpu__time_duration.sum,dram__bytes.sum` logs of scripts/profile_forward.py."""
import csv, collections, sys
N_GLUE = 10          # conv_first, att_gate launches (6: four nbp_att_scale + two level-5 gates), head<8> per forward
def load(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]; h = rows[hi]
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > h.index("Metric Value"):
            per.setdefault(r[h.index("ID")], {"n": r[h.index("Kernel Name")]})[r[h.index("Metric Name")]] = float(r[h.index("Metric Value")].replace(",", ""))
    v = list(per.values())
    first = max(i for i, k in enumerate(v) if "conv_first" in k["n"])
    return v[first:]
tabs = [load(p) for p in sys.argv[1:]]
for i in range(len(tabs[0])):
    line = f"{tabs[0][i]['n'].split('(')[0][-28:]:28s}"
    for t in tabs:
        k = t[i]; us = k["gpu__time_duration.sum"] / 1000
        line += f" | {us:7.1f} us {k.get('dram__bytes.sum', 0) / 1e6:7.1f} MB {k.get('dram__bytes.sum', 0) / k['gpu__time_duration.sum']:6.0f} GB/s"
    print(line)
print("total".ljust(28) + "".join(f" | {sum(k['gpu__time_duration.sum'] for k in t) / 1000:7.1f} us" + " " * 26 for t in tabs))
