"""Where the end-to-end overhead of one bench step goes (host arrays in, host results out through api.analyze_arrays):
engine construction, device loop, result read-back; for the monolithic engine (what N = 1 uses for 20 s) and for the
frame-range engine (what every rank of N > 1 uses; forced here with stream_threshold_bytes=0).  Prints a cProfile of each."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, bench
from eaqhm_b200 import api

sec = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
inp = bench.long_signal_inputs(sec)
ones = np.ones(len(inp["voiced5"]), bool)
for name, kw in (("Engine", {}), ("LongSignalEngine", dict(stream_threshold_bytes=0, shard_seconds=sec))):
    def make():
        return api.make_engine(inp["s"], inp["fs"], inp["f0s"], ones, ones, inp["frame_step"], gender=(90, 250), **kw)
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h = make()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        out = h.run_device(fixed_adaptations=6, prefetch_results=True)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        r = h.results(out)
        t3 = time.perf_counter()
        del h
        print("%s rep %d: construct %.1f ms, device loop %.1f ms, results %.1f ms" % (name, rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    pr = cProfile.Profile(); pr.enable()
    h = make(); torch.cuda.synchronize()
    pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
    out = h.run_device(fixed_adaptations=6, prefetch_results=True); torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    r = h.results(out)
    pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(10)
    del h
