#!/usr/bin/env python
"""Print one row per kernel of an `ncu --metrics ... --csv` log (first launch of each kernel name + launch count)."""
import collections
import csv
import io
import sys

txt = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(txt))))
agg = collections.OrderedDict()
for r in rows:
    agg.setdefault((r["ID"], r["Kernel Name"].split("(")[0], r["Grid Size"]), {})[r["Metric Name"]] = r["Metric Value"]
seen = collections.OrderedDict()
for (i, name, grid), m in agg.items():
    seen.setdefault(name, []).append((grid, m))
print("| kernel | launches | grid | us | MB read | MB written | dram % | sm % | regs | warps active % |")
print("|---|---|---|---|---|---|---|---|---|---|")
for name, ls in seen.items():
    grid, m = ls[-1]
    f = lambda k: float(m.get(k, "nan").replace(",", ""))
    print(f"| `{name}` | {len(ls)} | {grid} | {f('gpu__time_duration.sum') / 1e3:.1f} | {f('dram__bytes_read.sum') / 1e6:.1f} | "
          f"{f('dram__bytes_write.sum') / 1e6:.1f} | {f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{f('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {f('launch__registers_per_thread'):.0f} | "
          f"{f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} |")
