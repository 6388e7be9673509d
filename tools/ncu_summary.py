"""Summarise an .ncu-rep (ncu --set full capture) into the table committed under profiles/.

Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_<what>_summary.txt
Runs `ncu -i <rep> --page raw --csv` (no GPU needed) and keeps the columns the DESIGN/roofline discussion uses.
"""
import csv
import io
import subprocess
import sys

COLS = [
    "Kernel Name",
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__cluster_dim_x",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units, data = rows[0], rows[1], rows[2:]
    # some metrics appear several times, with a unit prefix ("TPC.TriageCompute.sm__pipe_tensor...") and possibly empty:
    # match by suffix and take the first non-empty value per kernel
    hits = []
    for c in COLS:
        h = [i for i, name in enumerate(header) if name == c] + [i for i, name in enumerate(header) if name.endswith("." + c)]
        if h:
            hits.append((c, h))
    print([c for c, _ in hits])
    print([units[h[0]] for _, h in hits])
    for r in data:
        print([next((r[i][:48] for i in h if r[i] != ""), "") for _, h in hits])


if __name__ == "__main__":
    main()
