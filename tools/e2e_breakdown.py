import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."); sys.path.insert(0, ROOT)
import numpy as np, torch, bench
from eaqhm_b200 import engine
inp = bench.shard_inputs(0, 20.0); utts = [engine.Utterance(**inp)]
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e = engine.Engine(utts); torch.cuda.synchronize(); t1 = time.perf_counter()
    r = e.run_device(); torch.cuda.synchronize(); t2 = time.perf_counter()
    out = e.results(*r); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("init %.1f ms  run %.1f ms  results %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); e = engine.Engine(utts); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
