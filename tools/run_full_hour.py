"""One-off: the WHOLE config-c2 signal (3600 s, 16 kHz) on one GPU as consecutive frame-range shards, each analysed
completely (shard-local stopping).  Prints one JSON line with the total device time and audio-s/s.  Synthetic shard
generation on the host is excluded from the timed region (it is preprocessing) but reported."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from eaqhm_b200 import engine

shard_s = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
total_s = float(sys.argv[2]) if len(sys.argv) > 2 else 3600.0
n = int(round(total_s / shard_s))
t_gen = t_gpu = 0.0
solves = 0
adapt_hist = {}
srer_w = 0.0
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n):
    t0 = time.perf_counter()
    inp = bench.shard_inputs(i, shard_s)
    t_gen += time.perf_counter() - t0
    eng = engine.Engine([engine.Utterance(**inp)])
    torch.cuda.synchronize()
    ev0.record()
    SRER, best_a, fs_ = eng.run_device()
    ev1.record()
    torch.cuda.synchronize()
    t_gpu += ev0.elapsed_time(ev1) * 1e-3
    solves += fs_
    adapt_hist[len(SRER[0])] = adapt_hist.get(len(SRER[0]), 0) + 1
    srer_w += max(SRER[0])
    del eng
    if i % 20 == 0:
        print("shard %d/%d  gpu %.1f s  gen %.1f s" % (i, n, t_gpu, t_gen), file=sys.stderr, flush=True)
print(json.dumps({"workload": "c2 full signal as %d shards of %.0f s" % (n, shard_s), "audio_seconds": n * shard_s,
                  "gpu_seconds": t_gpu, "audio_s_per_s": n * shard_s / t_gpu, "frame_solves": solves,
                  "frame_solves_per_s": solves / t_gpu, "adaptations_histogram": adapt_hist,
                  "mean_best_SRER_dB": srer_w / n, "host_signal_generation_seconds": t_gen}))
