"""Per-kernel device time of the adaptation loop on the small-system configurations (c1 SA19: P 106-194; c3 batch: P <= 346,
mostly ~200; c5 batch: P 62-126), for kernel experiments on that regime:
    [EAQHM_CHOL_WARPS=1] [EAQHM_LIB=build/lib_x.so] python tools/small_p_timing.py [utterances=32]  -> one JSON line per configuration"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from common import load_golden, inputs_from_golden
from eaqhm_b200 import _lib, engine, host, synthetic

nutt = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = _lib.lib()


def timed(name, utts, **kw):
    eng = engine.Engine(utts, **kw)
    eng.run_device()
    L.eaqhm_profile_reset(); L.eaqhm_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); ev0.record()
    SRER, best, solves = eng.run_device()
    ev1.record(); torch.cuda.synchronize()
    L.eaqhm_profile_enable(0)
    prof = {k: round(v[0], 2) for k, v in _lib.profile_read().items() if v[0] >= 0.3}
    fl = eng.useful_flops
    ms = ev0.elapsed_time(ev1)
    print(json.dumps({"config": name, "ms": round(ms, 2), "frame_solves": int(solves), "kernel_ms": prof,
                      "chol_tflops": round(fl["chol"] / (prof.get("chol_solve", 1e9) * 1e-3) / 1e12, 2),
                      "gram_tflops_conventional": round(fl["gram"] / (prof.get("gram", 1e9) * 1e-3) / 1e12, 2),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("EAQHM_")}}), flush=True)
    del eng


g, kw = load_golden("sa19")
inp = inputs_from_golden(g, kw)
timed("c1", [engine.Utterance(s=inp["s"], fs=inp["fs"], f0s=inp["f0s"], voiced5=inp["voiced5"], frame_step=inp["frame_step"],
                              f0min=inp["f0min"], Kmax=inp["Kmax"])])
timed("c3 x%d" % nutt, [engine.Utterance(**d) for d in bench.corpus_inputs(0, nutt)])
utts = []
for i in range(nutt):
    cfg = synthetic.config_c5(i, duration=3.0)
    s, f0s = synthetic.make(cfg)
    t5 = (np.arange(1, len(s), 80) - 1) / 16000
    v5 = np.asarray(cfg["voiced_mask_of_t"](t5), bool)
    utts.append(engine.Utterance(s=s, fs=16000, f0s=f0s, voiced5=v5, frame_step=80, f0min=250.0,
                                 Kmax=host.harmonic_capacity(16000, f0s)[1], det_part=host.voiced_only_target(s, v5, v5, 80)))
timed("c5 x%d" % nutt, utts)
