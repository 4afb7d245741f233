"""Per-kernel device time from an `ncu --metrics gpu__time_duration.sum --csv` log: count, total, average, max."""
import csv, io, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rd = csv.reader(io.StringIO("".join(lines)))
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
per = collections.defaultdict(list)
for r in rd:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    sc = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ix["Metric Unit"]], 1.0)
    per[r[ix["Kernel Name"]].split("(")[0]].append(v * sc)
tot = sum(sum(v) for v in per.values())
for k, v in sorted(per.items(), key=lambda x: -sum(x[1]))[:14]:
    print(f"{k[:44]:44s} n={len(v):4d} total {sum(v) / 1e3:9.2f} ms {100 * sum(v) / tot:5.1f}%  avg {sum(v) / len(v) / 1e3:8.3f} ms  max {max(v) / 1e3:8.3f} ms")
