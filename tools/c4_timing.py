"""Config c4 (48 kHz, f0 ~ 80 Hz, P = 1,190, L = 1,801): device time and per-kernel split of the adaptation loop.
    python tools/c4_timing.py [seconds=0.25] [maxAdpt=2] [fs=48000]  -> one JSON line  (EAQHM_CHOL_PANEL=0/1 selects the
factorisation; other sample rates give intermediate system sizes: 24 kHz P = 590, 32 kHz P = 790)"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
from eaqhm_b200 import _lib, engine, host, synthetic

sec = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
maxAdpt = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fs = int(sys.argv[3]) if len(sys.argv) > 3 else 48000
cfg = dict(synthetic.config_c4(duration=sec), fs=fs)
s, f0s = synthetic.make(cfg)
fstep = int(round(0.005 * fs))
n5 = len(np.arange(1, len(s), fstep))
u = engine.Utterance(s=s, fs=fs, f0s=f0s, voiced5=np.ones(n5, bool), frame_step=fstep, f0min=70.0,
                     Kmax=host.harmonic_capacity(fs, f0s)[1])
eng = engine.Engine([u], analysisWindow=64, maxAdpt=maxAdpt, scratch_budget_bytes=80 << 30)
L = _lib.lib()
eng.run_device(fixed_adaptations=maxAdpt + 1)
torch.cuda.synchronize()
L.eaqhm_profile_reset(); L.eaqhm_profile_enable(1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
SRER, best, solves = eng.run_device(fixed_adaptations=maxAdpt + 1)
ev1.record(); torch.cuda.synchronize()
L.eaqhm_profile_enable(0)
ms = ev0.elapsed_time(ev1)
prof = {k: round(v[0], 2) for k, v in _lib.profile_read().items() if v[0] >= 0.05}
fl = eng.useful_flops
print(json.dumps({"fs": fs, "P": 2 * (2 * int(eng.nnz_max) + 1), "panel": os.environ.get("EAQHM_CHOL_PANEL", "default"), "seconds": sec, "adaptations": len(SRER[0]), "ms": ms,
                  "audio_s_per_s": sec / (ms * 1e-3), "frame_solves": int(solves), "SRER": [float(x) for x in SRER[0]],
                  "kernel_ms": prof, "algorithmic_flops": fl,
                  "gram_tflops_executed": 0.75 * fl["gram"] / (prof.get("gram", 1) * 1e-3) / 1e12,
                  "chol_tflops": fl["chol"] / ((prof.get("chol_solve", 0) + prof.get("chol_update", 0)) * 1e-3) / 1e12}))
