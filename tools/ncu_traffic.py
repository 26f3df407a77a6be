"""DRAM traffic per launch of the hot kernels from a committed `ncu --set full` capture.

    ncu -i gpurun_out/<report>.ncu-rep --page raw --csv > profiles/<name>_raw.csv
    python tools/ncu_traffic.py profiles/<name>_raw.csv --workload c2 --shard-seconds 20 \
        --command "python bench.py --steps 1 --warmup 1" > profiles/ncu_traffic.json

Reads the wide `--page raw --csv` export (one row per profiled launch, first row metric names, second row units), adds
dram__bytes_read.sum + dram__bytes_write.sum per launch and averages over the launches of each kernel class.  bench.py
reports the figure of its dominant kernel as `roofline.traffic` (per launch, like `roofline.achieved`) and names this
file as the source; without a capture of the benchmarked workload it reports null.
"""
import argparse
import csv
import json
import re
import sys

CLASSES = [("gram", r"k_gram"), ("chol_update", r"k_chol_update"), ("chol_solve", r"k_chol_solve|k_chol_real"),
           ("basis", r"k_basis"), ("interp", r"k_interp"), ("synth", r"k_synth")]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--shard-seconds", type=float, default=20.0)
    ap.add_argument("--command", default="")
    ap.add_argument("--chunks-per-adaptation", type=int, default=1,
                    help="launches per kernel and adaptation in the captured run (the analysed frames are processed in chunks "
                         "sized by the scratch budget): bench.py rescales the per-launch figure to its own chunking")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    name_c = col["Kernel Name"]
    rd, wr, tm = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"], col["gpu__time_duration.sum"]
    out = {}
    for key, pat in CLASSES:
        sel = [r for r in data if re.search(pat, r[name_c])]
        if not sel:
            continue
        # adaptive launches only: adaptation 0 builds its matrices from lag tables (other kernels)
        by = [float(r[rd]) * SCALE[units[rd]] + float(r[wr]) * SCALE[units[wr]] for r in sel]
        ms = [float(r[tm]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[tm]] for r in sel]
        out[key] = {"sass_name": sel[0][name_c].split("(")[0], "launches": len(sel),
                    "dram_bytes_per_launch": sum(by) / len(by), "dram_bytes_total": sum(by),
                    "ms_per_launch_under_ncu": sum(ms) / len(ms)}
    json.dump({"source": args.csv, "command": args.command, "workload": args.workload,
               "shard_seconds": args.shard_seconds, "chunks_per_adaptation": args.chunks_per_adaptation, "kernels": out},
              sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
