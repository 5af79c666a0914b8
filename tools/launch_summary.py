"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: count, mean duration, share."""
import collections
import csv
import re
import sys


def main(path):
    rows = []
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        nm = r["Kernel Name"]
        m = re.search(r"(\w+_kernel)", nm)
        t = re.search(r"_kernel<([^>]*)>", nm)
        nm = (m.group(1) if m else nm[:40]) + ("<" + t.group(1) + ">" if t else "")
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        rows.append((nm, v))
    agg = collections.OrderedDict()
    for nm, v in rows:
        agg.setdefault(nm, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{path}: {len(rows)} launches, {tot:.1f} us")
    for nm, v in agg.items():
        print(f"  {nm:62s} n={len(v):3d} mean={sum(v) / len(v):8.1f} us share={sum(v) / tot:5.1%}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
