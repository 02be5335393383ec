"""Read an `ncu --page raw --csv` export and print, per kernel launch, duration, DRAM bytes and the
figures bench.py's traffic table wants (bytes per cell needs the grid, given as --cells)."""
import argparse
import csv
import json
import sys

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--cells", type=float, default=1024 * 1024 * 256)
args = ap.parse_args()
rows = list(csv.reader(open(args.csv)))
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr = i
        break
if hdr is None:
    sys.exit("no header row in " + args.csv)
names, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(names)}


def num(r, key):
    try:
        return float(r[col[key]].replace(",", ""))
    except (KeyError, ValueError):
        return None


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


for r in rows[hdr + 2:]:
    if len(r) < len(names):
        continue
    rd = to_bytes(num(r, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]])
    wr = to_bytes(num(r, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]])
    dur = num(r, "gpu__time_duration.sum")
    dur_unit = units[col["gpu__time_duration.sum"]]
    dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(dur_unit, 1)
    out = {"kernel": r[col["Kernel Name"]], "grid": r[col.get("Grid Size", 0)], "block": r[col.get("Block Size", 0)],
           "duration_ms": dur_ms, "dram_read_GB": rd / 1e9, "dram_write_GB": wr / 1e9,
           "bytes_per_cell": (rd + wr) / args.cells, "dram_GBps": (rd + wr) / 1e9 / (dur_ms * 1e-3),
           "registers": num(r, "launch__registers_per_thread"),
           "dram_throughput_pct": num(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
           "issue_active_pct": num(r, "sm__inst_issued.avg.pct_of_peak_sustained_active") or num(r, "smsp__issue_active.avg.pct")}
    print(json.dumps(out))
