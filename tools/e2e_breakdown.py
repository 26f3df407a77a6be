"""Where the end-to-end overhead of one bench step goes: engine construction from host arrays, device loop, result read-back."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, bench
from eaqhm_b200 import longsignal
inp = bench.long_signal_inputs(20.0)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e = longsignal.LongSignalEngine(shard_seconds=20.0, **inp)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out = e.run_device()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    r = e.results(*out)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    del e
    print("rep %d: construct %.1f ms, device loop %.1f ms, results %.1f ms" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
e = longsignal.LongSignalEngine(shard_seconds=20.0, **inp); torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
