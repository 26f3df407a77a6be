"""Markdown summary of the round's ncu evidence from the exported CSVs (tools/evidence_run.sh):

    python tools/ncu_summary.py profiles/r02_launches_c2.csv profiles/r02_ncu_raw_c2.csv [profiles/r02_ncu_raw_c4.csv] \
        [--bench profiles/r02_bench_c2_1gpu.json] > profiles/r02_ncu_summary.md

* launch list (`ncu --metrics gpu__time_duration.sum`): launches and total time per kernel, share of the profiled run, next to
  the live CUDA-event share of the plain bench run (the per-launch times under ncu are serialised and cold-cache: the SHARES
  must agree, not the absolutes);
* full-set captures (`ncu --set full`, `--page raw --csv`): one row per profiled launch with the counters DESIGN.md quotes.
"""
import argparse
import csv
import json
import re
import sys
from collections import OrderedDict

SCALE_T = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}
SCALE_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
HBM_PEAK = 6548.5e9     # MEASURED_PEAKS.json hbm_gbs
CLASS = [("gram", r"k_gram"), ("chol_update", r"k_chol_update"), ("chol_solve", r"k_chol_solve|k_chol_real"),
         ("basis", r"k_basis"), ("interp", r"k_interp"), ("a0 build", r"k_a0_build")]


def short(name):
    m = re.search(r"(k_[a-z0-9_]+(<[^>]*>)?)", name)
    return m.group(1) if m else name[:40]


def launch_list(path, bench):
    rows = [r for r in csv.reader(open(path)) if r and r[0].strip('"').isdigit()]
    hdr = None
    for r in csv.reader(open(path)):
        if r and r[0] == "ID":
            hdr = r
            break
    ci = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows:
        if r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = short(r[ci["Kernel Name"]])
        t = float(r[ci["Metric Value"]].replace(",", "")) * SCALE_T.get(r[ci["Metric Unit"]], 1e-6)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    live = None
    if bench:
        b = json.load(open(bench))
        kt = b["roofline"]["kernel_ms_per_step"]
        tot = sum(kt.values())
        live = {k: 100.0 * v / tot for k, v in kt.items()}
    print("## Launch list of the bench command (`%s`)\n" % path)
    print("| kernel | launches | total ms under ncu | share | live CUDA-event share of its class (plain run) |")
    print("|---|---:|---:|---:|---|")
    small = [0, 0.0]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / total < 0.004:
            small[0] += n
            small[1] += t
            continue
        cls = next((c for c, pat in CLASS if re.search(pat, k)), None)
        lv = ""
        if live and cls and cls.replace(" ", "_") in live:
            lv = "%s %.1f %%" % (cls, live[cls.replace(" ", "_")])
        elif live and cls == "a0 build":
            lv = "(in gram)"
        print("| %s | %d | %.3f | %.1f %% | %s |" % (k, n, t, 100 * t / total, lv))
    print("| everything else (torch copy / fill kernels, small stages) | %d | %.3f | %.1f %% | |" % (small[0], small[1], 100 * small[1] / total))
    print()


def full_set(path, title):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=None):
        if name not in ci:
            return float("nan")
        try:
            v = float(r[ci[name]].replace(",", ""))
        except ValueError:
            return float("nan")
        if scale:
            v *= scale.get(units[ci[name]], 1.0)
        return v
    dm = "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg"
    print("## %s (`%s`)\n" % (title, path))
    print("| kernel | ms | grid | regs | DMMA pipe active % | scalar FP64 % | SM throughput % | occupancy % | DRAM read GB | DRAM write GB | "
          "DRAM bytes/s as % of the measured 6,548 GB/s | L2 hit % | long-scoreboard stall / issue | math-pipe throttle / issue |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for r in data:
        cyc = val(r, "sm__cycles_elapsed.max")
        print("| %s | %.3f | %s | %s | %.1f | %.1f | %.1f | %.1f | %.2f | %.2f | %.1f | %.1f | %.2f | %.2f |" % (
            short(r[ci["Kernel Name"]]), val(r, "gpu__time_duration.sum", SCALE_T), r[ci["Grid Size"]].replace(",", ""),
            r[ci["launch__registers_per_thread"]], 100 * val(r, dm) / cyc if cyc == cyc and cyc else float("nan"),
            val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            val(r, "dram__bytes_read.sum", SCALE_B) / 1e9, val(r, "dram__bytes_write.sum", SCALE_B) / 1e9,
            100 * (val(r, "dram__bytes_read.sum", SCALE_B) + val(r, "dram__bytes_write.sum", SCALE_B))
            / (val(r, "gpu__time_duration.sum", SCALE_T) * 1e-3) / HBM_PEAK,
            val(r, "lts__t_sector_hit_rate.pct"),
            val(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
            val(r, "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio")))
    print()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("launches")
    ap.add_argument("raw", nargs="+")
    ap.add_argument("--bench", default=None)
    a = ap.parse_args()
    print("# Round 2 ncu evidence (B200, sm_100a; recipe: tools/evidence_run.sh, generated by tools/ncu_summary.py)\n")
    launch_list(a.launches, a.bench)
    titles = ["`ncu --set full --clock-control none`: hot kernels of adaptation 1 of the default c2 step (two chunks of frames per kernel)",
              "`ncu --set full`: config c4 (48 kHz, P = 1,186), panel Cholesky",
              "`ncu --set full`: config c4, adaptive basis and Gram"]
    for i, p in enumerate(a.raw):
        full_set(p, titles[min(i, len(titles) - 1)])
