"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file x.csv) per kernel: total, launches, share, average.
usage: python tools/ncu_launch_list.py gpurun_out/launches.csv [skip_first_n_launches] > profiles/rNN_ncu_launches.txt"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = rows[0]; ci = {n: i for i, n in enumerate(h)}
unit = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
agg = collections.OrderedDict()
for r in rows[1 + skip:]:
    name = r[ci["Kernel Name"]].split("(")[0][-90:]
    ms = float(r[ci["Metric Value"]].replace(",", "")) * unit.get(r[ci["Metric Unit"]], 1e-6)
    a = agg.setdefault(name, [0.0, 0]); a[0] += ms; a[1] += 1
tot = sum(a[0] for a in agg.values())
print(f"# {len(rows) - 1 - skip} launches, {tot:.3f} ms of kernel time in total (cold-cache, serialised under ncu: compare SHARES, not absolutes)")
for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms:10.3f} ms {n:5d} launches {100 * ms / tot:6.1f}%  avg {ms / n:9.4f} ms  {name}")
