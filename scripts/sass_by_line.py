#!/usr/bin/env python
"""Attribute the per-instruction counts of an ncu source-page CSV (SASS view) to source lines.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> > sass.csv
    cuobjdump -xelf <file> lib.so; nvdisasm -g -c <file>.cubin > dis.txt
    python scripts/sass_by_line.py sass.csv dis.txt '<mangled function substring>' source.cu

Instructions are matched in order (the n-th instruction of the function in both listings); inlined code is attributed
to the innermost line nvdisasm reports."""
import collections
import csv
import re
import sys


def main():
    sass_csv, dis, fn, src = sys.argv[1:5]
    r = list(csv.reader(open(sass_csv)))
    h = r[1]
    ia, ism = h.index("Instructions Executed"), h.index("# Samples")
    rows = r[2:]
    lines, cur, infn = [], None, False
    for ln in open(dis):
        if ln.startswith(".text."):
            infn = fn in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File ".*?", line (\d+)', ln)
        if m:
            cur = int(m.group(1))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    assert len(lines) == len(rows), (len(lines), len(rows))
    cnt, smp = collections.Counter(), collections.Counter()
    for l, x in zip(lines, rows):
        cnt[l] += int(x[ia])
        smp[l] += int(x[ism])
    tot, tots = sum(cnt.values()), sum(smp.values())
    text = open(src).read().split("\n")
    print(f"total instructions {tot}, samples {tots}")
    for l, n in cnt.most_common(40):
        print(f"{l:5d} {n:12d} {100 * n / tot:5.1f}%  smp {100 * smp[l] / tots:5.1f}%  {text[l - 1].strip()[:90] if l else ''}")


if __name__ == "__main__":
    main()
