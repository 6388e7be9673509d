"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) of
`bench.py --train-only`: one training step = the launches between two obb_frontend_kernel launches.

Usage: python tools/launch_list_summary.py gpurun_out/launches_train.csv
"""
import collections
import csv
import sys

TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "s": 1e3, "second": 1e3}
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    iK, iM, iV, iID, iU = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    byid = collections.OrderedDict()
    for r in rows[1:]:
        d = byid.setdefault(r[iID], {"k": r[iK].replace("void ", "").replace("durf::", "")[:40]})
        v = float(r[iV].replace(",", ""))
        if r[iM].startswith("gpu__time"):
            v *= TIME.get(r[iU], 1.0)
        elif r[iM].startswith("dram"):
            v *= BYTES.get(r[iU], 1.0)
        d[r[iM]] = v
    launches = list(byid.values())
    marks = [i for i, d in enumerate(launches) if d["k"].startswith("obb_frontend_kernel")]
    if len(marks) < 2:
        raise SystemExit("need two obb_frontend_kernel launches to delimit a step")
    step = launches[marks[0]:marks[1]]
    total = sum(d["gpu__time_duration.sum"] for d in step)
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in step:
        g = agg[d["k"]]
        g[0] += 1
        g[1] += d["gpu__time_duration.sum"]
        g[2] += d.get("dram__bytes_read.sum", 0.0)
        g[3] += d.get("dram__bytes_write.sum", 0.0)
    print(f"one train step: {len(step)} launches, {total:.3f} ms of kernels (ncu: serialised, cold caches - compare shares)")
    for k, g in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        bw = (g[2] + g[3]) / g[1] / 1e9 if g[1] else 0.0
        print(f"{k:42s} n={g[0]:3d} {g[1]:8.3f} ms {g[1] / total:6.1%}  dram read {g[2] / 1e9:7.2f} GB  write {g[3] / 1e9:7.2f} GB  {bw:6.2f} TB/s")


if __name__ == "__main__":
    main()
