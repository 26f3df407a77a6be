#!/usr/bin/env python
"""Benchmark of the eaQHM hot path (BASELINE.json metric: audio-seconds analysed+resynthesised per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shard-seconds S]

Workload (config c2 of BASELINE.json): the synthetic 16 kHz harmonic AM-FM signal
(f0(t) = 170 + 80 sin(2 pi 0.25 t + phi0), gender=(90,250)) is processed as
contiguous frame ranges; one STEP is the complete eaQHM analysis+synthesis (all adaptations
until the signal's SRER stops improving, functions.py:163-402) of the first N x S seconds of
that signal, S = `--shard-seconds` of audio per GPU (weak scaling).  With N GPUs every rank owns
one contiguous range; after every analysis pass the ranks exchange the per-frame parameters of
their boundary frames with their neighbours (NCCL send/recv) and all-gather the residual
statistics, so that all ranks take the reference's single convergence decision
(eaqhm_b200/longsignal.py).  The whole hour is 3600/S such ranges; tools/run_full_hour.py streams
it through one GPU with the same engine.

Timed regions:  value = steps with the shard's inputs already resident in HBM (device time,
CUDA events, max over ranks);  e2e = the same through the public API with host buffers
(host->device copies of samples and tables, device->host copy of waveform + parameters inside).
`--impl reference` times the oracle port of the reference's CPU path (oracle/eaqhm_oracle.py)
on a bounded excerpt of the same workload on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-seconds analysed+resynthesised per second (eaQHM, c2 1-hour 16 kHz AM-FM signal, frame-range shards)"
UNIT = "audio-s/s"
GRAM_TRAFFIC_BYTES = 23.1e9  # dram read+write per k_gram_warp launch: mean of the two launches of adaptation 1
                             # (38.7 and 7.5 GB; 12,501 and 8,769 frames), ncu --set full, profiles/r01_ncu_summary_final.md
FP64_PEAK_TFLOPS = 36.9     # measured on this pool's B200 with tools/fp64_peak.cu (profiles/r01_fp64_peak.jsonl);
                            # MEASURED_PEAKS.json holds no FP64 figure.  DFMA and DMMA share one pipe.


def shard_inputs(shard_index, seconds):
    """Samples, f0 grid and flags of shard `shard_index` of the c2 signal (host arrays)."""
    from eaqhm_b200 import synthetic
    cfg = synthetic.config_c2()
    base = cfg["f0_of_t"]
    off = shard_index * seconds
    s, f0s = synthetic.harmonic_signal(lambda t: base(t + off), seconds, fs=cfg["fs"], seed=1000 + shard_index)
    n5 = len(np.arange(1, len(s), 80))
    return dict(s=s, fs=cfg["fs"], f0s=f0s, voiced5=np.ones(n5, bool), frame_step=80, f0min=float(cfg["gender"][0]),
                Kmax=int(round(int(cfg["fs"] / 2 - 200) / 90.0) + 10))   # min f0 of the whole hour (functions.py:120)


def long_signal_inputs(seconds):
    """The first `seconds` of the c2 signal (host arrays): samples, 5 ms f0 grid, flags."""
    from eaqhm_b200 import synthetic
    cfg = synthetic.config_c2(duration=seconds)
    s, f0s = synthetic.make(cfg)
    n5 = len(np.arange(1, len(s), 80))
    return dict(s=s, fs=cfg["fs"], f0s=f0s, voiced5=np.ones(n5, bool), frame_step=80, f0min=float(cfg["gender"][0]),
                Kmax=int(round(int(cfg["fs"] / 2 - 200) / 90.0) + 10))   # min f0 of the whole hour (functions.py:120)


def corpus_inputs(first_index, count, seconds=3.0):
    """`count` utterances of the config-c3 corpus (10,000 x 3 s at 16 kHz, utterance-sharded) as host arrays."""
    from eaqhm_b200 import synthetic
    out = []
    for i in range(first_index, first_index + count):
        cfg = synthetic.config_c3(i, duration=seconds)
        s, f0s = synthetic.make(cfg)
        n5 = len(np.arange(1, len(s), 80))
        out.append(dict(s=s, fs=cfg["fs"], f0s=f0s, voiced5=np.ones(n5, bool), frame_step=80, f0min=90.0,
                        Kmax=int(round(int(cfg["fs"] / 2 - 200) / f0s[:, 1].min()) + 10)))
    return out


def algorithmic_flops(N, M):
    """Structure-aware minimum of one frame solve: three Hermitian half-Grams of the
    M-column basis under weights w^2, n w^2, n^2 w^2 (12 L M^2), rhs (16 L M), complex Cholesky of
    P = 2M ((4/3) P^3) and two triangular solves (8 P^2)."""
    L = 2.0 * N + 1
    P = 2.0 * M
    gram = 12.0 * L * M * M
    chol = (4.0 / 3.0) * P ** 3
    return gram, chol, 16.0 * L * M + 8.0 * P * P


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_oracle_sample(inp, excerpt_seconds):
    """The reference's CPU algorithm (oracle port) on the first `excerpt_seconds` of the shard."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import eaqhm_oracle as O
    n = int(excerpt_seconds * inp["fs"])
    n5 = len(np.arange(1, n, inp["frame_step"]))
    t0 = time.perf_counter()
    r = O.eaqhm_oracle(inp["s"][:n], inp["fs"], inp["f0s"], inp["voiced5"][:n5], inp["frame_step"], inp["f0min"],
                       inp["Kmax"])
    dt = time.perf_counter() - t0
    solves = int(r.isVoiced.sum()) * len(r.SRER)
    return dict(seconds=dt, audio_seconds=n / inp["fs"], adaptations=len(r.SRER), frame_solves=solves)


def emit(fd, obj):
    """The ONE JSON line of the contract, on the process's original stdout."""
    os.write(fd, (json.dumps(obj) + "\n").encode())


def main():
    # libraries chat on stdout (NCCL prints its version there when NCCL_DEBUG is set on the box): keep the original
    # stdout for the JSON line only and send everything else to stderr
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shard-seconds", type=float, default=20.0)
    ap.add_argument("--cpu-excerpt-seconds", type=float, default=0.25)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"],
                    help="c2 (default, the metric's configuration): frame-range shards of the 1-hour signal; "
                         "c3: a batch of --batch 3-second utterances per GPU per step (utterance sharding)")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--scratch-gb", type=float, default=0.0,
                    help="analysis scratch budget in GiB (frames per launch follow from it); 0 = engine default")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    config = {"workload": "c2: synthetic 1-hour 16 kHz harmonic AM-FM signal (f0 90-250 Hz), its first %d x %.0f s as one "
                          "contiguous frame range of %.0f s of audio per GPU per step, full eaQHM analysis+synthesis "
                          "until the signal's SRER converges" % (world, args.shard_seconds, args.shard_seconds),
              "fs": 16000, "step": 15, "maxAdpt": 10, "shard_seconds": args.shard_seconds,
              "l2_policy": "per-step working set (dense tracks + Gram scratch, > 1 GB) exceeds the 126 MB L2",
              "parallelism": "contiguous frame ranges, one per GPU; per adaptation one neighbour exchange of the "
                             "boundary frames' parameters (NCCL send/recv) and one all-gather of the SRER statistics"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        inp = long_signal_inputs(max(1.0, args.cpu_excerpt_seconds))
        times, last = [], None
        for i in range(args.warmup + args.steps):
            if i < args.warmup and i > 0:
                continue          # one warm-up pass is enough for a CPU code path; keep the run bounded
            last = run_oracle_sample(inp, args.cpu_excerpt_seconds)
            if i >= args.warmup:
                times.append(last["seconds"])
        dt = float(np.mean(times))
        val = last["audio_seconds"] / dt
        cfg = dict(config)
        cfg["workload"] += " (reference arm: oracle port on the first %.2f s excerpt per step)" % args.cpu_excerpt_seconds
        emit(out_fd, {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "first %.2f s of shard 0, %d adaptations, %d frame solves; numpy/OpenBLAS threads = all cores"
                                       % (args.cpu_excerpt_seconds, last["adaptations"], last["frame_solves"]),
                             "frame_solves_per_s": last["frame_solves"] / dt},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})
        return 0

    import torch
    from eaqhm_b200 import _lib, engine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    budget = int(args.scratch_gb * (1 << 30)) if args.scratch_gb > 0 else None
    kw = {} if budget is None else dict(scratch_budget_bytes=budget)
    if args.workload == "c3":
        inps = corpus_inputs(rank * args.batch, args.batch)
        utts = [engine.Utterance(**i) for i in inps]
        inp = inps[0]
        args.shard_seconds = 3.0 * args.batch
        config["workload"] = ("c3: batch of %d synthetic 3-second utterances at 16 kHz per GPU per step (utterance-sharded "
                              "corpus), full eaQHM analysis+synthesis until each utterance's SRER converges" % args.batch)
        config["parallelism"] = "utterance sharding, %d utterances per GPU per step, no data-path collective" % args.batch
        make_engine = lambda: engine.Engine(utts, **kw)
    else:
        from eaqhm_b200 import longsignal
        inp = long_signal_inputs(args.shard_seconds * world)
        make_engine = lambda: longsignal.LongSignalEngine(shard_seconds=args.shard_seconds, rank=rank, world=world,
                                                          dist=dist, **inp, **kw)
    L = _lib.lib()

    # ---------------- device-resident arm ----------------------------------------------------
    eng = make_engine()
    for _ in range(args.warmup):
        res = eng.run_device()
    barrier()
    L.eaqhm_profile_reset()
    L.eaqhm_profile_enable(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        res = eng.run_device()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    L.eaqhm_profile_enable(0)
    prof = _lib.profile_read()
    ms_dev = ev0.elapsed_time(ev1) / args.steps
    SRER, best_a, frames_solved = res
    if args.workload == "c3":
        n_adapt = int(max(len(x) for x in SRER))
        srer_list = [float(x) for x in SRER[0]]
        n_frames = int(sum(int(f.analysed.sum()) for f in eng.ft))
        halo_bytes = 0
    else:
        n_adapt, srer_list = len(SRER), [float(x) for x in SRER]
        n_frames, halo_bytes = eng.frames_owned, int(eng.exchanged_bytes)
    launches = sum(c for _, c in prof.values()) // args.steps

    useful_flops = eng.useful_flops
    del eng
    torch.cuda.empty_cache()

    # ---------------- end-to-end arm (host buffers in, host results out) ----------------------
    e2e_parts = [0.0, 0.0, 0.0]

    def e2e_step():
        # what a caller of the public batch API does: host arrays in, host results out; the engine (device buffers)
        # lives only inside the call
        ta = time.perf_counter()
        e = make_engine()
        tb = time.perf_counter()
        out_dev = e.run_device()
        tc = time.perf_counter()
        r = e.results(*out_dev)
        td = time.perf_counter()
        for i, v in enumerate((tb - ta, tc - tb, td - tc)):
            e2e_parts[i] += v
        return int(e.h2d_bytes), int(e.d2h_bytes), r
    for _ in range(max(2, min(args.warmup, 3))):
        e2e_step()
    barrier()
    e2e_parts[:] = [0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h2d, d2h, r = e2e_step()
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / args.steps

    if dist is not None:
        t = torch.tensor([ms_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, t_e2e = float(t[0]), float(t[1])
    audio = args.shard_seconds * world
    value = audio / (ms_dev * 1e-3)
    e2e_value = audio / t_e2e

    # ---------------- roofline of the dominant kernel ------------------------------------------
    # algorithmic flops: accumulated on the device by k_active from the actual active-harmonic counts of every
    # frame solve; adaptation 0 uses Toeplitz tables instead of the Gram and is not counted there
    useful = useful_flops
    kt = {k: v[0] / args.steps for k, v in prof.items()}
    dom = max(("gram", "chol_solve", "basis"), key=lambda k: kt[k])
    # the Gram kernel forms every complex product from THREE real DMMA products (3M: Re = P1 + P2, Im = P3 - P1 + P2)
    # instead of four, so it executes 9 L M^2 real flops where the conventional count (8 per complex multiply-add,
    # accumulated on the device by k_active) is 12 L M^2: `achieved` is what the FP64 pipe really ran (useful flops of
    # the algorithm used), `achieved_conventional` the same time against the conventional count
    executed = {"gram": 0.75 * useful["gram"], "chol_solve": useful["chol"], "basis": useful["basis"]}
    conventional = {"gram": useful["gram"], "chol_solve": useful["chol"], "basis": useful["basis"]}
    achieved = executed[dom] / (kt[dom] * 1e-3) / 1e12 if kt[dom] > 0 else 0.0
    conv = conventional[dom] / (kt[dom] * 1e-3) / 1e12 if kt[dom] > 0 else 0.0
    roofline = {"bound": "tensor", "pipe": "FP64 (DFMA and DMMA share one pipe on sm_100a)", "kernel": dom,
                "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                "achieved_conventional": conv, "frac_conventional": conv / FP64_PEAK_TFLOPS,
                # dram__bytes_read.sum + dram__bytes_write.sum of k_gram_warp, mean of the launches captured with
                # `ncu --set full` on this command's default workload (profiles/r01_ncu_summary_final.md); the algorithmic
                # bytes of the same launches (basis planes read once + normal matrix written once) are 10.4 GB: panels
                # are re-fetched ~3x from DRAM (L2 hit 59 %), at 15 % DRAM utilisation under a kernel at 86 % of the FP64 pipe
                "traffic": (GRAM_TRAFFIC_BYTES if (dom == "gram" and args.workload == "c2" and args.shard_seconds == 20.0) else None),
                "peak_source": "measured FP64 burst (tools/fp64_peak.cu, profiles/r01_fp64_peak.jsonl); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "whole_step_executed_tflops": (executed["gram"] + executed["chol_solve"] + executed["basis"]) / (ms_dev * 1e-3) / 1e12,
                "kernel_ms_per_step": kt, "algorithmic_flops_per_step": useful,
                "all_kernels_frac_of_step": sum(kt.values()) / ms_dev}

    metric = METRIC if args.workload == "c2" else METRIC.replace("c2 1-hour 16 kHz AM-FM signal, frame-range shards",
                                                                 "c3 corpus of 3-second utterances, utterance-sharded")
    out = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": dict(config, adaptations_run=n_adapt, frames_per_gpu=n_frames,
                                               frame_solves_per_gpu_per_step=int(frames_solved),
                                               halo_bytes_sent_per_gpu_per_step=halo_bytes, SRER_dB=srer_list),
           "frame_solves_per_s": frames_solved * world / (ms_dev * 1e-3),
           "ms_per_adaptation": ms_dev / max(1, n_adapt),
           "realtime_factor_per_gpu": value / world,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": t_e2e * 1e3,
                   "ms_setup_loop_readback": [1e3 * v / args.steps for v in e2e_parts]},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c = run_oracle_sample(inp, args.cpu_excerpt_seconds)
        out["cpu_baseline"] = {"value": c["audio_seconds"] / c["seconds"], "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "oracle port (numpy restatement of functions.py:163-402) on the first %.2f s of the "
                                         "shard: %d adaptations, %d frame solves in %.1f s"
                                         % (c["audio_seconds"], c["adaptations"], c["frame_solves"], c["seconds"]),
                               "frame_solves_per_s": c["frame_solves"] / c["seconds"]}
    if rank == 0:
        emit(out_fd, out)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
