#!/usr/bin/env python
"""Per-layer table of the LAST forward's conv_gemm launches from `ncu --csv --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...`
logs of scripts/profile_forward.py (one column pair per log)."""
import collections, csv, re, sys
ORDER = ["Conv1.b"] + [f"Conv{l}.{ab}" for l in range(2, 6) for ab in "ab"]
for dec, lvls in ((1, (5, 4)), (2, (5, 4, 3, 2))):
    for l in lvls:
        ORDER += [f"Up{l}_{dec}", f"Att{l}_{dec}", f"Up_conv{l}_{dec}.a", f"Up_conv{l}_{dec}.b"]
def load(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]; h = rows[hi]
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > h.index("Metric Value") and "conv_gemm" in r[h.index("Kernel Name")]:
            e = per.setdefault(r[h.index("ID")], {"n": re.search(r"<([^>]*)>", r[h.index("Kernel Name")]).group(1)})
            v = float(r[h.index("Metric Value")].replace(",", ""))
            if "time" in r[h.index("Metric Name")]:
                v = v / 1000 if r[h.index("Metric Unit")] in ("ns", "nsecond") else v
            e[r[h.index("Metric Name")]] = v
    return list(per.values())[-len(ORDER):]
tabs = [load(p) for p in sys.argv[1:]]
print("layer".ljust(14) + "".join(f" | {p.split('/')[-1][:22]:>22s}      us  tensor%" for p in sys.argv[1:]))
for i, name in enumerate(ORDER):
    line = name.ljust(14)
    for t in tabs:
        k = t[i]
        line += f" | <{k['n']:>20s}> {k['gpu__time_duration.sum']:7.1f} {k.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):7.1f}"
    print(line)
print("total".ljust(14) + "".join(f" | {'':22s} {sum(k['gpu__time_duration.sum'] for k in t):7.1f}" + " " * 8 for t in tabs))
