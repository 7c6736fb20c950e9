"""Summarise an ncu launch-list csv: pre-decode (vision+prefill) and decode-step shares."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    seq.append((name, v))
last = max((i for i, (n, v) in enumerate(seq) if "flash_varlen_kernel<128>" in n), default=-1)
for label, part in (("pre-decode (preprocess + vision + prefill)", seq[:last + 1]), ("decode steps", seq[last + 1:])):
    agg = collections.OrderedDict()
    for n, v in part:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1
    print(f"{label}: {len(part)} launches, {tot:.0f} us")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"   {k[:58]:58s} n={c:4d} avg={t / c:8.2f} us  share={t / tot * 100:5.1f} %")
