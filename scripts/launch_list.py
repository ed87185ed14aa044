#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` log: per-kernel totals, or (--convs N) the last N conv_gemm launches in order."""
import csv, collections, re, sys
path = sys.argv[1]
rows = list(csv.reader(open(path, errors="ignore")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
data = [r for r in rows[hi + 1:] if len(r) > mv and r[mv].replace(",", "").replace(".", "").isdigit()]
def us(r):
    v = float(r[mv].replace(",", ""))
    return v / 1000 if r[mu] in ("ns", "nsecond") else v * 1000 if r[mu] in ("ms", "msecond") else v
if len(sys.argv) > 3 and sys.argv[2] == "--convs":
    n = int(sys.argv[3])
    convs = [r for r in data if "conv_gemm" in r[kn]][-n:]
    print(" ".join(f"{us(r):.0f}" for r in convs)); print("total", sum(us(r) for r in convs))
else:
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        k = re.sub(r"\(.*", "", r[kn])[:70]; agg[k][0] += 1; agg[k][1] += us(r)
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}% x{v[0]:5d} {k}")
    print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
