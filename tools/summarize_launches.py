"""ncu launch list (CSV of `--metrics gpu__time_duration.sum`) -> markdown table per kernel.
usage: python tools/summarize_launches.py launches.csv "<command that was profiled>" > summary.md"""
import collections, csv, sys

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = []
with open(path) as f:
    for line in f:
        if line.startswith('"ID"'):
            break
    for r in csv.reader(f):
        if len(r) > 5:
            rows.append(r)
agg = collections.OrderedDict()
for r in rows:
    val = float(r[-1].replace(",", ""))
    unit = r[-2]
    val = val / 1000 if unit == "ns" else (val * 1000 if unit == "ms" else val)
    name = r[4].split("(")[0].replace("void ", "").replace("gf::<unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += val
tot = sum(a[1] for a in agg.values())
print("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n")
print("Command: `%s`\n" % cmd)
print("%d launches captured. Times are cold-cache and serialised under the profiler: compare SHARES, "
      "not absolutes; numbers printed by bench.py under ncu are not bench values.\n" % len(rows))
print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k[:72], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
spmv = sum(a[1] for k, a in agg.items() if k.startswith("spmv"))
print("\nAll SpMV kernels (all multigrid levels): %.1f%% of the captured device time." % (100 * spmv / tot))
