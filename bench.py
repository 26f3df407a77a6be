#!/usr/bin/env python
"""Benchmark of the eaQHM hot path (BASELINE.json metric: audio-seconds analysed+resynthesised per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shard-seconds S] [--adaptations A]

Workload (config c2 of BASELINE.json): the synthetic 16 kHz harmonic AM-FM signal
(f0(t) = 170 + 80 sin(2 pi 0.25 t + phi0), gender=(90,250)) is processed as
contiguous frame ranges; one STEP is the complete eaQHM analysis+synthesis (functions.py:163-402) of the first
N x S seconds of that signal, S = `--shard-seconds` of audio per GPU (weak scaling).  With N GPUs every rank owns
one contiguous range; after every analysis pass the ranks exchange the per-frame parameters of their boundary frames
with their neighbours (NCCL send/recv) and all-gather the residual statistics, so that all ranks take the reference's
single convergence decision (eaqhm_b200/longsignal.py).

Fixed work: the reference stops when the SRER stops rising (functions.py:394-396), and increments of
1e-5 dB decide between 5 and 6 adaptations on these signals, so runs of different length would do different work.
The timed step therefore runs exactly `--adaptations` A adaptations (default 6 = the count the 20 s range takes by
itself) with the stop disabled, on every N and in the reference arm; the data-dependent run is reported next to it
(`data_dependent`).  `--adaptations 0` times the data-dependent loop itself.

Timed regions:  value = steps with the inputs already resident in HBM (device time, CUDA events, max over ranks);
e2e = the same through the plugin call `eaqhm_b200.api.analyze_arrays` with HOST buffers (host->device copies of
samples and tables, device->host copy of waveform + per-frame parameters inside), tensor-form result; the cost of
building the reference's list of Deterministic objects on top is reported separately (`e2e.with_det_components`).
`--impl reference` times the oracle port of the reference's CPU path (oracle/eaqhm_oracle.py) on a bounded excerpt of the
same workload on all host cores.
"""
import os
import sys

if "reference" in sys.argv:
    # the reference arm uses every host core: torchrun exports OMP_NUM_THREADS=1, and BLAS reads it at import
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import argparse  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-seconds analysed+resynthesised per second (eaQHM, c2 1-hour 16 kHz AM-FM signal, frame-range shards)"
UNIT = "audio-s/s"
FP64_PEAK_TFLOPS = 36.9     # measured on this pool's B200 with tools/fp64_peak.cu (profiles/r01_fp64_peak.jsonl);
                            # MEASURED_PEAKS.json holds no FP64 figure.  DFMA and DMMA share one pipe.
NCU_TRAFFIC = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # written by tools/ncu_traffic.py from an ncu CSV


def long_signal_inputs(seconds):
    """The first `seconds` of the c2 signal (host arrays): samples, 5 ms f0 grid, flags."""
    from eaqhm_b200 import synthetic
    cfg = synthetic.config_c2(duration=seconds)
    s, f0s = synthetic.make(cfg)
    n5 = len(np.arange(1, len(s), 80))
    return dict(s=s, fs=cfg["fs"], f0s=f0s, voiced5=np.ones(n5, bool), frame_step=80, f0min=float(cfg["gender"][0]),
                Kmax=int(round(int(cfg["fs"] / 2 - 200) / f0s[:, 1].min()) + 10))   # functions.py:120


def shard_inputs(shard_index, seconds):
    """Samples, f0 grid and flags of shard `shard_index` of the c2 signal (host arrays; tools/)."""
    from eaqhm_b200 import synthetic
    cfg = synthetic.config_c2()
    base = cfg["f0_of_t"]
    off = shard_index * seconds
    s, f0s = synthetic.harmonic_signal(lambda t: base(t + off), seconds, fs=cfg["fs"], seed=1000 + shard_index)
    n5 = len(np.arange(1, len(s), 80))
    return dict(s=s, fs=cfg["fs"], f0s=f0s, voiced5=np.ones(n5, bool), frame_step=80, f0min=float(cfg["gender"][0]),
                Kmax=int(round(int(cfg["fs"] / 2 - 200) / 90.0) + 10))


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


def blas_threads(n):
    """All host cores for the CPU leg, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=int(n))
        return int(n)
    except Exception:
        return int(os.environ.get("OPENBLAS_NUM_THREADS", os.environ.get("OMP_NUM_THREADS", "0")) or 0)


def run_oracle_sample(inp, excerpt_seconds, adaptations):
    """The reference's CPU algorithm (oracle port) on the first `excerpt_seconds` of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import eaqhm_oracle as O
    n = int(excerpt_seconds * inp["fs"])
    n5 = len(np.arange(1, n, inp["frame_step"]))
    t0 = time.perf_counter()
    r = O.eaqhm_oracle(inp["s"][:n], inp["fs"], inp["f0s"], inp["voiced5"][:n5], inp["frame_step"], inp["f0min"],
                       inp["Kmax"], fixed_adaptations=adaptations or None)
    dt = time.perf_counter() - t0
    solves = int(r.isVoiced.sum()) * len(r.SRER)
    return dict(seconds=dt, audio_seconds=n / inp["fs"], adaptations=len(r.SRER), frame_solves=solves)


def emit(fd, obj):
    """The ONE JSON line of the contract, on the process's original stdout."""
    os.write(fd, (json.dumps(obj) + "\n").encode())


def measured_traffic(kernel, workload, shard_seconds, chunks_now=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    of this command (profiles/ncu_traffic.json, extracted from the raw ncu CSV by tools/ncu_traffic.py); None when no
    capture of this workload is committed.  The analysed frames of an adaptation are processed in chunks sized by the
    scratch budget; when this run's chunk count (`chunks_now`) differs from the captured run's, the per-launch figure
    is rescaled by their ratio (the bytes per adaptation do not depend on the chunking)."""
    try:
        t = json.load(open(NCU_TRAFFIC))
    except Exception:
        return None, "no ncu capture committed (profiles/ncu_traffic.json)"
    e = t.get("kernels", {}).get(kernel)
    if e is None or t.get("workload") != workload or float(t.get("shard_seconds", -1)) != float(shard_seconds):
        return None, "profiles/ncu_traffic.json holds no capture of this kernel / workload"
    per_launch, note = float(e["dram_bytes_per_launch"]), ""
    cap = int(t.get("chunks_per_adaptation", 0))
    if chunks_now and cap and int(chunks_now) != cap:
        per_launch *= cap / float(chunks_now)
        note = "; captured with %d launches per adaptation, this run has %d: per-launch figure rescaled" % (cap, int(chunks_now))
    return per_launch, "%s (%s, %d launches averaged%s)" % (t.get("source"), e.get("sass_name", kernel),
                                                           int(e.get("launches", 0)), note)


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
    ap.add_argument("--adaptations", type=int, default=6,
                    help="fixed work: exactly this many adaptations per step, convergence stop disabled (0 = the "
                         "reference's data-dependent stop)")
    ap.add_argument("--cpu-excerpt-seconds", type=float, default=0.25)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true",
                    help="world > 1: skip the single-rank re-run of the whole signal on rank 0 after the timed region")
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
    fixed = args.adaptations if args.adaptations > 0 else None
    work = ("exactly %d adaptations, convergence stop disabled" % fixed) if fixed else \
        "until the signal's SRER stops rising (functions.py:394-396)"
    config = {"workload": "c2: synthetic 1-hour 16 kHz harmonic AM-FM signal (f0 90-250 Hz), its first %d x %.0f s as one "
                          "contiguous frame range of %.0f s of audio per GPU per step, full eaQHM analysis+synthesis, %s"
                          % (world, args.shard_seconds, args.shard_seconds, work),
              "fs": 16000, "step": 15, "maxAdpt": 10, "shard_seconds": args.shard_seconds, "fixed_adaptations": fixed,
              "l2_policy": "per-step working set (dense tracks + Gram scratch, > 1 GB) exceeds the 126 MB L2",
              "parallelism": "contiguous frame ranges, one per GPU; per adaptation one neighbour exchange of the "
                             "boundary frames' parameters (NCCL send/recv) and one all-gather of the SRER statistics"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = blas_threads(cores)
        inp = long_signal_inputs(max(1.0, args.cpu_excerpt_seconds))
        times, last = [], None
        for i in range(args.warmup + args.steps):
            if i < args.warmup and i > 0:
                continue          # one warm-up pass is enough for a CPU code path; keep the run bounded
            last = run_oracle_sample(inp, args.cpu_excerpt_seconds, fixed)
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
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "blas_threads": threads, "kind": "port",
                             "sample": "first %.2f s of the signal, %d adaptations, %d frame solves; numpy/OpenBLAS "
                                       "threads = %d" % (args.cpu_excerpt_seconds, last["adaptations"],
                                                         last["frame_solves"], threads),
                             "frame_solves_per_s": last["frame_solves"] / dt},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0})
        return 0

    import torch
    from eaqhm_b200 import _lib, api, engine
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
                              "corpus), full eaQHM analysis+synthesis of every utterance, %s"
                              % (args.batch, work.replace("the signal's", "each utterance's")))
        config["parallelism"] = "utterance sharding, %d utterances per GPU per step, no data-path collective" % args.batch

        class CorpusHandle:
            """Same surface as api.PathHandle for a batch of utterances (api.analyze_batch is make + run)."""
            def __init__(self):
                self.eng = engine.Engine(utts, **kw)

            def run_device(self, fixed_adaptations=None):
                return self.eng.run_device(fixed_adaptations=fixed_adaptations)

            def results(self, out):
                return self.eng.results(*out)
        make_handle = CorpusHandle
    else:
        inp = long_signal_inputs(args.shard_seconds * world)
        ones = np.ones(len(inp["voiced5"]), bool)
        make_handle = lambda: api.make_engine(inp["s"], inp["fs"], inp["f0s"], ones, ones, inp["frame_step"],
                                              gender=(90, 250), shard_seconds=args.shard_seconds, rank=rank, world=world,
                                              dist=dist, **kw)
    L = _lib.lib()

    # ---------------- device-resident arm ----------------------------------------------------
    h = make_handle()
    for _ in range(args.warmup):
        res = h.run_device(fixed_adaptations=fixed)
    barrier()
    L.eaqhm_profile_reset()
    L.eaqhm_profile_enable(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        res = h.run_device(fixed_adaptations=fixed)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    L.eaqhm_profile_enable(0)
    prof = _lib.profile_read()
    ms_dev = ev0.elapsed_time(ev1) / args.steps
    SRER, best_a, frames_solved = res
    eng = h.eng
    if args.workload == "c3":
        n_adapt = int(max(len(x) for x in SRER))
        srer_list = [float(x) for x in SRER[0]]
        n_frames = int(sum(int(f.analysed.sum()) for f in eng.ft))
        halo_bytes = 0
    else:
        sr = SRER[0] if getattr(h, "single", False) else SRER
        n_adapt, srer_list = len(sr), [float(x) for x in sr]
        n_frames = int(eng.ft[0].analysed.sum()) if getattr(h, "single", False) else eng.frames_owned
        halo_bytes = 0 if getattr(h, "single", False) else int(eng.exchanged_bytes)
    launches = sum(c for _, c in prof.values()) // args.steps
    useful_flops = eng.useful_flops
    engine_kind = type(eng).__name__

    # the data-dependent run (the reference's own stop rule) next to the fixed-work figure: one extra timed step
    dd = None
    if fixed is not None:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r2 = h.run_device()
        e1.record()
        barrier()
        ms_dd = e0.elapsed_time(e1)
        if args.workload == "c3":
            sr2 = max(r2[0], key=len)
        else:
            sr2 = r2[0][0] if getattr(h, "single", False) else r2[0]
        if dist is not None:
            t = torch.tensor([ms_dd], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_dd = float(t[0])
        dd = {"value": args.shard_seconds * world / (ms_dd * 1e-3), "unit": UNIT, "ms_per_step": ms_dd,
              "adaptations_run": len(sr2), "frame_solves_per_gpu": int(r2[2])}

    # multi-rank parity (world > 1): keep this rank's part of the last fixed-work result for the check below
    own_s = own_range = None
    if world > 1 and args.workload == "c2" and not args.no_parity_check:
        h.run_device(fixed_adaptations=fixed)
        own_s, own_range = eng.s_best.clone(), (eng.R_lo, eng.R_hi)
    del h, eng
    torch.cuda.empty_cache()

    # ---------------- end-to-end arm: the plugin call with host buffers -----------------------
    e2e_bytes = [0, 0]

    def e2e_step():
        if args.workload == "c3":
            hh = make_handle()
            r = hh.results(hh.run_device(fixed_adaptations=fixed))
            e2e_bytes[:] = [int(hh.eng.h2d_bytes), int(hh.eng.d2h_bytes)]
            return r
        return api.analyze_arrays(inp["s"], inp["fs"], inp["f0s"], ones, ones, inp["frame_step"], gender=(90, 250),
                                  shard_seconds=args.shard_seconds, rank=rank, world=world, dist=dist,
                                  fixed_adaptations=fixed, **kw)
    for _ in range(max(2, min(args.warmup, 3))):
        r = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = e2e_step()
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / args.steps
    if args.workload == "c2":
        # bytes the call moved, counted from the arrays it copies: samples (and SRER target) + per-frame tables in,
        # waveform + three parameter arrays + DC + per-adaptation statistics out
        nfr = len(r.ti)
        own_frames = nfr if world == 1 else (r.stats["owned_frames"][1] - r.stats["owned_frames"][0])
        own_samples = len(r.s_recon) if world == 1 else (r.stats["owned_samples"][1] - r.stats["owned_samples"][0])
        e2e_bytes[:] = [int(own_samples * 8 + own_frames * 40), int(own_samples * 8 + own_frames * (3 * r.am.shape[1] + 1) * 8)]
    pack_ms = None
    if world == 1 and args.workload == "c2":
        t1 = time.perf_counter()
        det = api.pack_det_components(r)
        pack_ms = (time.perf_counter() - t1) * 1e3
        del det

    if dist is not None:
        t = torch.tensor([ms_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, t_e2e = float(t[0]), float(t[1])
    audio = args.shard_seconds * world
    value = audio / (ms_dev * 1e-3)
    e2e_value = audio / t_e2e

    # ---------------- multi-rank parity: the same signal on rank 0 alone -----------------------
    parity = None
    if own_s is not None:
        from eaqhm_b200 import longsignal
        total = len(inp["s"])
        if rank == 0:
            whole = torch.zeros(total, dtype=torch.float64, device="cuda")
            whole[own_range[0]:own_range[1]] = own_s
            bounds = [None] * world
            dist.gather_object(own_range, bounds, dst=0)
            for src in range(1, world):
                lo, hi = bounds[src]
                buf = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
                dist.recv(buf, src=src)
                whole[lo:hi] = buf
            one = longsignal.LongSignalEngine(shard_seconds=args.shard_seconds, rank=0, world=1, dist=None, **inp, **kw)
            S1, _, _ = one.run_device(fixed_adaptations=fixed)
            n = min(len(S1), len(srer_list))
            parity = {"max_ds": float((whole - one.s_best).abs().max().item()),
                      "dSRER": float(np.abs(np.array(S1[:n]) - np.array(srer_list[:n])).max()) if len(S1) == len(srer_list) else float("inf"),
                      "against": "LongSignalEngine, world=1, the same %d shards of %.0f s on rank 0 alone" % (world, args.shard_seconds)}
            del one
        else:
            dist.gather_object(own_range, None, dst=0)
            dist.send(own_s, dst=0)
        dist.barrier()

    # ---------------- roofline of the dominant kernel ------------------------------------------
    # algorithmic flops: accumulated on the device by k_active from the actual active-harmonic counts of every
    # frame solve; adaptation 0 uses lag tables instead of the Gram and is not counted there.
    # The Gram kernel forms every complex product from THREE real DMMA products (3M), so it executes 9 L M^2 real flops
    # where the conventional count is 12 L M^2: `achieved` is what the FP64 pipe really ran, `achieved_conventional`
    # the same time against the conventional count.
    useful = useful_flops
    kt = {k: v[0] / args.steps for k, v in prof.items()}
    dom = max(("gram", "chol_solve", "basis"), key=lambda k: kt[k])
    executed = {"gram": 0.75 * useful["gram"], "chol_solve": useful["chol"], "basis": useful["basis"]}
    conventional = {"gram": useful["gram"], "chol_solve": useful["chol"], "basis": useful["basis"]}
    achieved = executed[dom] / (kt[dom] * 1e-3) / 1e12 if kt[dom] > 0 else 0.0
    conv = conventional[dom] / (kt[dom] * 1e-3) / 1e12 if kt[dom] > 0 else 0.0
    # launches of the dominant kernel per adaptation in this run (its class also counts one launch per chunk at adaptation 0)
    adapt_run = max(1, len(srer_list))
    chunks_now = max(1, int(round(prof[dom][1] / args.steps / adapt_run))) if dom in prof else None
    traffic, traffic_src = measured_traffic(dom, args.workload, args.shard_seconds, chunks_now)
    roofline = {"bound": "tensor", "pipe": "FP64 (DFMA and DMMA share one pipe on sm_100a)", "kernel": dom,
                "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                "achieved_conventional": conv, "frac_conventional": conv / FP64_PEAK_TFLOPS,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "measured FP64 burst (tools/fp64_peak.cu, profiles/r01_fp64_peak.jsonl); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "per_kernel_frac": {k: executed[k] / (kt[k] * 1e-3) / 1e12 / FP64_PEAK_TFLOPS
                                    for k in ("gram", "chol_solve") if kt[k] > 0},
                "whole_step_executed_tflops": (executed["gram"] + executed["chol_solve"] + executed["basis"]) / (ms_dev * 1e-3) / 1e12,
                "kernel_ms_per_step": kt, "algorithmic_flops_per_step": useful,
                "all_kernels_frac_of_step": sum(kt.values()) / ms_dev}

    metric = METRIC if args.workload == "c2" else METRIC.replace("c2 1-hour 16 kHz AM-FM signal, frame-range shards",
                                                                 "c3 corpus of 3-second utterances, utterance-sharded")
    out = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": dict(config, engine=engine_kind, adaptations_run=n_adapt, frames_per_gpu=n_frames,
                                               frame_solves_per_gpu_per_step=int(frames_solved),
                                               halo_bytes_sent_per_gpu_per_step=halo_bytes, SRER_dB=srer_list),
           "frame_solves_per_s": frames_solved * world / (ms_dev * 1e-3),
           "ms_per_adaptation": ms_dev / max(1, n_adapt),
           "realtime_factor_per_gpu": value / world,
           "data_dependent": dd,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_bytes[0], "d2h_bytes_per_step": e2e_bytes[1],
                   "ms_per_step": t_e2e * 1e3, "call": "eaqhm_b200.api.analyze_arrays (tensor-form result)"
                   if args.workload == "c2" else "engine.Engine(batch).run()",
                   "with_det_components": None if pack_ms is None else
                   {"value": audio / (t_e2e + pack_ms * 1e-3), "pack_ms": pack_ms,
                    "note": "api.pack_det_components: the reference's list of Deterministic objects (object-dtype arrays, Q8)"}},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if parity is not None:
        out["parity_vs_single_rank"] = parity

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = blas_threads(cores)
        c = run_oracle_sample(inp, args.cpu_excerpt_seconds, fixed)
        out["cpu_baseline"] = {"value": c["audio_seconds"] / c["seconds"], "unit": UNIT, "cores": cores,
                               "blas_threads": threads, "kind": "port",
                               "sample": "oracle port (numpy restatement of functions.py:163-402) on the first %.2f s of the "
                                         "signal: %d adaptations, %d frame solves in %.1f s"
                                         % (c["audio_seconds"], c["adaptations"], c["frame_solves"], c["seconds"]),
                               "frame_solves_per_s": c["frame_solves"] / c["seconds"]}
    if rank == 0:
        emit(out_fd, out)
    if dist is not None:
        dist.destroy_process_group()
    if parity is not None and (parity["max_ds"] > 0 or parity["dSRER"] > 0):
        print("multi-rank result differs from the single-rank run: %r" % parity, file=sys.stderr)
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
