"""Reference-vs-reference noise floor of the parity metrics (container only, like make_golden.py).

    OPENBLAS_NUM_THREADS=1 python tests/golden/make_floor.py sa19 synth_a_adpt3 ...

Re-runs the UNMODIFIED reference on a committed fixture's inputs with a different BLAS thread count than
the run that made the fixture (make_golden.py ran with the container's default, 8 OpenBLAS threads) and
records, per fixture, the six error measures of tests/common.compare_to_golden between the two reference
runs in tests/golden/floors.json.  The explicit inverse of the reference (functions.py:465,530) is computed by
LAPACK with a thread-count dependent blocking, so the two runs differ in the last bits of every solve; the
adaptation dynamics amplify that.  The GPU parity tests accept max(1e-9, FLOOR_FACTOR x floor) for the per-harmonic
relative measures (tests/test_gpu_parity.py).
"""
import json
import os
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
from scipy.io import wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import make_golden as MG  # noqa: E402
from common import compare_to_golden, load_golden  # noqa: E402

FLOORS = os.path.join(HERE, "floors.json")
KEYS = ("dSRER", "ds", "am_rel_framemax", "am_rel_strong", "fm_rel_strong", "ph_abs_strong", "a0_abs")


def rerun(name):
    g, kw = load_golden(name)
    gender = str(g["gender"])
    if gender == "tuple":
        gender = (float(g["f0min"]), float(g["f0max"]))
        gender = tuple(int(x) if float(x).is_integer() else x for x in gender)
    kwargs = {}
    for k, v in kw.items():
        kwargs[k] = bool(v) if k == "fullWaveform" else int(v)
    # SA19 uses the reference's own SWIPE' track; the synthetic cases patch swipep with the committed f0 grid
    f0s = None if name == "sa19" else g["f0s"]
    with tempfile.TemporaryDirectory() as td:
        wav = os.path.join(td, name + ".wav")
        wavfile.write(wav, int(g["fs"]), g["pcm"])
        s_recon, SRER, det, cap = MG.run_reference(wav, gender, f0s, **kwargs)
    packed = MG.pack_det(det, int(g["Kmax"]))
    res = SimpleNamespace(s_recon=np.asarray(s_recon, float).reshape(-1), SRER=[float(x) for x in SRER],
                          ti=packed["ti"], isSpeech=packed["isSpeech"], isVoiced=packed["isVoiced"], a0=packed["a0"],
                          am=packed["am"], fm=packed["fm"], ph=packed["ph"])
    c = compare_to_golden(res, g)
    assert c["n_adapt"][0] == c["n_adapt"][1] and c["flags_ok"] and c["support_equal"], c
    return {k: float(c[k]) for k in KEYS}


if __name__ == "__main__":
    floors = json.load(open(FLOORS)) if os.path.exists(FLOORS) else {}
    threads = os.environ.get("OPENBLAS_NUM_THREADS", "default")
    for name in sys.argv[1:]:
        out = rerun(name)
        out["blas_threads_of_second_run"] = threads
        floors[name] = out
        print(name, out, flush=True)
        json.dump(floors, open(FLOORS, "w"), indent=1, sort_keys=True)
