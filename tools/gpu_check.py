"""First-contact GPU check: helper seams and the full loop against the oracle / golden fixtures."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import ctypes as C
from common import load_golden, inputs_from_golden, compare_to_golden
import eaqhm_oracle as O
from eaqhm_b200 import _lib, engine

L = _lib.lib()
print("devices", L.eaqhm_device_count())
h = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "helpers.npz"))

def P(a): return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(_lib.c_f64p)

# iqhmLS seam
s, f0r, win = h["iq_s"], h["iq_f0range"], h["iq_win"]
M = len(f0r); amp = np.zeros(2*M); slope = np.zeros(2*M)
_lib.check(L.eaqhm_iqhmLS_complexamps(P(s), len(s), P(f0r), M, P(win), 16000.0, amp.ctypes.data_as(_lib.c_f64p), slope.ctypes.data_as(_lib.c_f64p)))
a = amp.view(complex); b = slope.view(complex)
print("iqhmLS amp rel", np.abs(a-h["iq_amp"]).max()/np.abs(h["iq_amp"]).max(), "slope", np.abs(b-h["iq_slope"]).max()/np.abs(h["iq_slope"]).max())
# eaqhmLS seam
s, am, fm, win = h["ea_s"], h["ea_am"], h["ea_fm"], h["ea_win"]
M = am.shape[1]; amp = np.zeros(2*M); slope = np.zeros(2*M)
_lib.check(L.eaqhm_eaqhmLS_complexamps(P(s), len(s), P(am), P(fm), M, P(win), 16000.0, amp.ctypes.data_as(_lib.c_f64p), slope.ctypes.data_as(_lib.c_f64p)))
a = amp.view(complex); b = slope.view(complex)
print("eaqhmLS amp rel", np.abs(a-h["ea_amp"]).max()/np.abs(h["ea_amp"]).max(), "slope", np.abs(b-h["ea_slope"]).max()/np.abs(h["ea_slope"]).max())
# phase integration seam
idx = np.ascontiguousarray(h["pi_idx"], dtype=np.int64); out = np.zeros(idx[-1]-idx[0]+1)
_lib.check(L.eaqhm_phase_integr_interpolation(P(h["pi_fm"]), P(h["pi_ph"]), len(h["pi_fm"]), idx.ctypes.data_as(_lib.c_i64p), len(idx), out.ctypes.data_as(_lib.c_f64p)))
print("phase_integr abs", np.abs(out-h["pi_out"]).max())

for name in sys.argv[1:] or ["synth_a_adpt0", "synth_a_adpt3"]:
    g, kw = load_golden(name)
    inp = inputs_from_golden(g, kw)
    u = engine.Utterance(s=inp["s"], fs=inp["fs"], f0s=inp["f0s"], voiced5=inp["voiced5"], frame_step=inp["frame_step"],
                         f0min=inp["f0min"], Kmax=inp["Kmax"], det_part=inp["det_part"])
    t0 = time.time()
    eng = engine.Engine([u], step=inp["step"], maxAdpt=inp["maxAdpt"], analysisWindow=inp["analysisWindow"])
    t1 = time.time()
    res = eng.run()[0]
    import torch; torch.cuda.synchronize()
    t2 = time.time()
    print(name, "setup %.3fs run %.3fs" % (t1-t0, t2-t1), "SRER", res.SRER, "golden", list(g["SRER"]), res.stats)
    print("   ", compare_to_golden(res, g))
