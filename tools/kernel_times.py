"""Per-kernel-class device time of one step (complete adaptation loop) on a c2 range, for kernel experiments:
EAQHM_LIB=build/lib_x.so python tools/kernel_times.py [seconds=5] [steps=2]   -> one JSON line."""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import bench
from eaqhm_b200 import _lib, longsignal

sec = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
inp = bench.long_signal_inputs(sec)
eng = longsignal.LongSignalEngine(shard_seconds=sec, **inp)
L = _lib.lib()
for _ in range(2):
    eng.run_device()
torch.cuda.synchronize()
L.eaqhm_profile_reset(); L.eaqhm_profile_enable(1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(steps):
    SRER, best, solves = eng.run_device()
ev1.record(); torch.cuda.synchronize()
L.eaqhm_profile_enable(0)
prof = {k: round(v[0] / steps, 3) for k, v in _lib.profile_read().items() if v[0] / steps >= 0.05}
print(json.dumps({"lib": os.environ.get("EAQHM_LIB", "in-tree"), "seconds": sec, "ms_per_step": round(ev0.elapsed_time(ev1) / steps, 2),
                  "adaptations": len(SRER), "SRER_last": float(SRER[-1]), "kernels_ms": prof}))
