#!/usr/bin/env python
"""Per-layer table from `ncu --csv --metrics ...` logs of scripts/profile_forward.py: the last 33 conv_gemm launches = one forward."""
import csv, sys, collections
def load(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]; idc, kn, mn, mv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv: continue
        per.setdefault(r[idc], {"name": r[kn]})[r[mn]] = float(r[mv].replace(",", "")) if r[mv].replace(",", "").replace(".", "").replace("e+", "").isdigit() else r[mv]
    return list(per.values())
order = ["Conv1.b"] + [f"Conv{l}.{ab}" for l in range(2, 6) for ab in "ab"]
for dec, lvls in ((1, (5, 4)), (2, (5, 4, 3, 2))):
    for l in lvls:
        order += [f"Up{l}_{dec}", f"Att{l}_{dec}", f"Up_conv{l}_{dec}.a", f"Up_conv{l}_{dec}.b"]
tabs = [load(p)[-len(order):] for p in sys.argv[1:]]
print(f"{'layer':16s}" + "".join(f" | {p.split('/')[-1][:22]:>22s} us  tensor%" for p in sys.argv[1:]))
tot = [0.0] * len(tabs)
for i, name in enumerate(order):
    line = f"{name:16s}"
    for j, t in enumerate(tabs):
        k = t[i]; us = k["gpu__time_duration.sum"] / 1000.0; tot[j] += us
        tmpl = k["name"][k["name"].find("<"):k["name"].find(">") + 1]
        line += f" | {tmpl:>14s} {us:8.1f} {k.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):7.1f}"
    print(line)
print("total".ljust(16) + "".join(f" | {'':14s} {t:8.1f}        " for t in tot))
