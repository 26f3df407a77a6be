"""Per-phase cycle breakdown of k_chol_solve (needs the -DEAQHM_CHOL_TIMING build passed as argv[1])."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from eaqhm_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
import bench
from eaqhm_b200 import engine
L = _lib.lib()
L.eaqhm_debug_chol_cycles.argtypes = [C.POINTER(C.c_double), C.c_int]
inp = bench.shard_inputs(0, 5.0)
eng = engine.Engine([engine.Utterance(**inp)])
eng.run_device()
out = np.zeros(8)
L.eaqhm_debug_chol_cycles(out.ctypes.data_as(C.POINTER(C.c_double)), 1)
eng.run_device()
L.eaqhm_debug_chol_cycles(out.ctypes.data_as(C.POINTER(C.c_double)), 1)
names = ["k loop (DMMA)", "staging (R load + smem)", "factor diag", "wait D barrier", "solve rows", "wait end barrier", "back substitution", "-"]
tot = out.sum()
for n, v in zip(names, out):
    print("%-26s %6.1f%%   %.0f cycles/warp-frame" % (n, 100 * v / tot, v / (31620 * 2)))
