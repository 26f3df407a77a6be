"""Shared helpers of the test-suite: golden loading and array-level input assembly."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from eaqhm_b200 import host  # noqa: E402


KEYS = ("dSRER", "ds", "am_rel_framemax", "am_rel_strong", "fm_rel_strong", "ph_abs_strong", "a0_abs")
# north_star tolerances: SRER 0.01 dB, samples 1e-7 absolute, amplitudes / frequencies 1e-9 relative (phases: 1e-9 rad)
NORTH_STAR = dict(dSRER=0.01, ds=1e-7, am_rel_framemax=1e-9, am_rel_strong=1e-9, fm_rel_strong=1e-9, ph_abs_strong=1e-9,
                  a0_abs=1e-9)
# A metric may exceed its north-star tolerance only up to FLOOR_FACTOR x the reference's OWN run-to-run noise on the
# same fixture (tests/golden/floors.json: the unmodified reference re-run with 1 instead of 8 OpenBLAS threads,
# tests/golden/make_floor.py).  Two runs of the reference differ by `floor`; a third implementation that is as
# accurate as the reference differs from either by a comparable amount, and FLOOR_FACTOR = 10 leaves room for the
# sample-to-sample spread of a maximum over ~10^5 harmonics.
FLOOR_FACTOR = 10.0


def load_floors():
    import json
    p = os.path.join(GOLDEN, "floors.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def generic_floor(key):
    """Largest measured reference-vs-reference floor of `key` over the SHORT fixtures (<= 1.2 s, <= 5 adaptations):
    the floor used for oracle comparisons on fresh seeded inputs of that size, which have no fixture of their own."""
    fl = load_floors()
    vals = [v[key] for n, v in fl.items() if n != "sa19" and key in v]
    return max(vals) if vals else 0.0


def tolerance(name, key):
    """Test tolerance of metric `key` on fixture `name` (None: a seeded input without fixture)."""
    fl = load_floors()
    floor = fl[name][key] if name in fl and key in fl[name] else generic_floor(key)
    return max(NORTH_STAR[key], FLOOR_FACTOR * floor)


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    kw = dict(zip([str(k) for k in g["kw_keys"]], [float(v) for v in g["kw_vals"]]))
    return g, kw


def inputs_from_golden(g, kw):
    """Array-level inputs of the hot path (what functions.py:83-146 hands to the loop)."""
    s = g["pcm"].astype(float) / host.NORMALIZE
    fs = int(g["fs"])
    step = int(kw.get("step", 15))
    aw = int(kw.get("analysisWindow", 32))
    frame_step = int(g["frame_step"])
    full = bool(kw.get("fullWaveform", 1))
    if full:
        speech5, voiced5 = host.full_waveform_flags(g["speech5"], g["voiced5"], len(s), frame_step, aw * step)
        det = None
    else:
        speech5, voiced5 = g["speech5"], g["voiced5"]
        det = host.voiced_only_target(s, speech5, voiced5, frame_step)
    return dict(s=s, fs=fs, f0s=g["f0s"], voiced5=voiced5, speech5=speech5, frame_step=frame_step,
                f0min=float(g["f0min"]), Kmax=int(g["Kmax"]), det_part=det, step=step,
                maxAdpt=int(kw.get("maxAdpt", 10)), analysisWindow=aw)


def compare_to_golden(res, g, label=""):
    """Returns a dict of error measures of a result (oracle or CUDA) against a golden fixture."""
    sel = g["frame_sel"]
    v = g["det_isVoiced"]
    out = {}
    n = min(len(res.SRER), len(g["SRER"]))
    out["n_adapt"] = (len(res.SRER), len(g["SRER"]))
    out["dSRER"] = float(np.abs(np.array(res.SRER[:n]) - g["SRER"][:n]).max())
    out["ds"] = float(np.abs(res.s_recon - g["s_recon"]).max())
    out["flags_ok"] = bool((res.ti[sel] == g["det_ti"]).all() and (res.isVoiced[sel] == v).all()
                           and (res.isSpeech[sel] == g["det_isSpeech"]).all())
    am_r, am_g = res.am[sel][v], g["det_am"][v]
    fmax = np.abs(am_g).max(axis=1, keepdims=True)
    out["am_rel_framemax"] = float((np.abs(am_r - am_g) / fmax).max())
    strong = am_g > fmax * 1e-3                       # harmonics within 60 dB of the frame maximum
    fm_r, fm_g = res.fm[sel][v], g["det_fm"][v]
    fsel = strong & (fm_g != 0)                      # fm is 0 where only the fallback amplitude was written (Q10)
    out["fm_rel_strong"] = float((np.abs(fm_r - fm_g)[fsel] / np.abs(fm_g[fsel])).max())
    out["fm_zero_equal"] = bool(((fm_r == 0) == (fm_g == 0)).all())
    out["am_rel_strong"] = float((np.abs(am_r - am_g)[strong] / np.abs(am_g[strong])).max())
    out["ph_abs_strong"] = float(np.abs(res.ph[sel][v] - g["det_ph"][v])[strong].max())
    out["a0_abs"] = float(np.abs(res.a0[sel] - g["det_a0"])[v].max())
    out["support_equal"] = bool(((am_r != 0) == (am_g != 0)).all())
    return out
