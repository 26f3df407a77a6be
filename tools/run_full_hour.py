"""The WHOLE config-c2 signal (3600 s, 16 kHz, 3.84 M frames) on ONE GPU through the streamed long-signal engine
(eaqhm_b200/longsignal.py): per-frame parameter store resident, dense tracks rebuilt shard by shard, ONE global SRER
decision per adaptation exactly like the reference's loop (functions.py:388,394-396).  Prints one JSON line.
The synthetic signal is generated with torch on the device (same formula and random draws as
eaqhm_b200/synthetic.py:harmonic_signal, which would need minutes of host time for 57.6 M samples) and handed to the
engine as a host array; generation is outside the timed region.
Usage: python tools/run_full_hour.py [shard_seconds] [total_seconds] [maxAdpt]
       torchrun --nproc-per-node N tools/run_full_hour.py ...   every rank streams its contiguous block of the same shards
       (per-adaptation halo exchange + statistics all-gather over NCCL); rank 0 prints the line, GPU time = max over ranks."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
from eaqhm_b200 import _lib, longsignal, synthetic

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
shard_s = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
total_s = float(sys.argv[2]) if len(sys.argv) > 2 else 3600.0
maxAdpt = int(sys.argv[3]) if len(sys.argv) > 3 else 10


def signal_on_device(cfg, duration, noise_db=-50.0, peak=0.5):
    fs, rng = cfg["fs"], np.random.default_rng(cfg["seed"])
    length = int(round(duration * fs))
    dev = torch.device("cuda")
    t = torch.arange(length, dtype=torch.float64, device=dev) / fs
    f0_h = np.asarray(cfg["f0_of_t"](np.arange(length) / fs), float)
    f0 = torch.from_numpy(f0_h).to(dev)
    # running sum on the host: a device scan is not bitwise reproducible from run to run, and every rank of a multi-GPU run
    # (and the single-GPU run it is compared with) must analyse the SAME samples
    phi = torch.from_numpy(2 * np.pi * np.cumsum(f0_h) / fs).to(dev)
    Kmax = int(np.floor((fs / 2 - 200) / float(f0.min())))
    s = torch.zeros(length, dtype=torch.float64, device=dev)
    for k in range(1, Kmax + 1):
        theta, r, d, p = rng.uniform(-np.pi, np.pi), rng.uniform(2.0, 6.0), rng.uniform(0.0, 0.3), rng.uniform(-np.pi, np.pi)
        amp = (1.0 / k) * (1.0 + d * torch.sin(2 * np.pi * r * t + p))
        amp = torch.where(k * f0 <= fs / 2 - 200, amp, torch.zeros_like(amp))
        s += amp * torch.cos(k * phi + theta)
    s *= peak / float(s.abs().max())
    s += torch.from_numpy(rng.standard_normal(length)).to(dev) * peak * 10 ** (noise_db / 20)
    s = torch.clamp(torch.round(s * 32768), -32768, 32767) / 32768
    tg = synthetic.f0_grid_times(length, fs)
    f0s = np.stack([tg, np.asarray(cfg["f0_of_t"](tg), float), np.ones_like(tg)], axis=1)
    return s.cpu().numpy(), f0s


t0 = time.perf_counter()
cfg = synthetic.config_c2(duration=total_s)
s, f0s = signal_on_device(cfg, total_s)
torch.cuda.empty_cache()
n5 = len(np.arange(1, len(s), 80))
t_gen = time.perf_counter() - t0
t0 = time.perf_counter()
eng = longsignal.LongSignalEngine(s, 16000, f0s, np.ones(n5, bool), 80, 90.0, int(round(7800 / 90.0) + 10),
                                  shard_seconds=shard_s, maxAdpt=maxAdpt, rank=rank, world=world, dist=dist)
eng.prepare_readback()          # page-lock the 9.5 GB of result staging before the loop
torch.cuda.synchronize()
t_setup = time.perf_counter() - t0
L = _lib.lib()
L.eaqhm_profile_reset(); L.eaqhm_profile_enable(1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def progress(a, SRER):
    if rank == 0:
        print("adaptation %d  SRER %.6f dB" % (a, SRER[-1]), file=sys.stderr, flush=True)


if dist is not None:
    dist.barrier()
torch.cuda.synchronize()


ev0.record()
SRER, best_a, solves = eng.run_device(progress, prefetch_results=True)    # read-back overlapped with the next adaptation
ev1.record()
torch.cuda.synchronize()
t_gpu = ev0.elapsed_time(ev1) * 1e-3
if dist is not None:
    tt = torch.tensor([t_gpu], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_gpu = float(tt[0])
L.eaqhm_profile_enable(0)
prof = {k: round(v[0] * 1e-3, 3) for k, v in _lib.profile_read().items()}
t0 = time.perf_counter()
res = eng.results(SRER, best_a, solves)
t_out = time.perf_counter() - t0
if world > 1:
    # the owned ranges of all ranks -> rank 0 (outside every timed region) for the check of the returned waveform
    parts = [None] * world
    dist.gather_object((eng.R_lo, eng.R_hi, res.s_recon[eng.R_lo:eng.R_hi]), parts if rank == 0 else None, dst=0)
    if rank == 0:
        full = np.zeros(len(s))
        for lo, hi, x in parts:
            full[lo:hi] = x
        srer_check = 20 * np.log10(np.std(s) / np.std(s - full))
    solves_t = torch.tensor([float(solves)], dtype=torch.float64, device="cuda")
    dist.all_reduce(solves_t)
    solves = int(solves_t[0])
else:
    srer_check = 20 * np.log10(np.std(s) / np.std(s - res.s_recon))
if rank == 0:
  print(json.dumps({"workload": "c2 whole signal, %d shards of %.0f s per rank streamed through %d GPU(s), global SRER decision"
                              % (len(eng.shards), shard_s, world), "n_gpus": world,
                  "audio_seconds": total_s, "gpu_seconds": t_gpu, "audio_s_per_s": total_s / t_gpu,
                  "frames": int(eng.full.No_ti_total), "frame_solves": int(solves), "frame_solves_per_s": solves / t_gpu,
                  "adaptations_run": len(SRER), "SRER_dB": [float(x) for x in SRER], "best_adaptation": int(best_a),
                  "SRER_of_returned_waveform_dB": float(srer_check),
                  "device_memory_peak_GiB": torch.cuda.max_memory_allocated() / 2 ** 30,
                  "kernel_seconds": prof, "host_signal_generation_seconds": t_gen, "engine_setup_seconds": t_setup,
                  "result_readback_seconds": t_out, "not_pd_frames": res.stats["not_pd_frames"],
                  "halo_bytes_sent_by_rank0": res.stats["exchanged_bytes"]}))
if dist is not None:
    dist.destroy_process_group()
