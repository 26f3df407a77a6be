"""Generates the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference, imported through ref_shim.py) in the build container.

    python tests/golden/make_golden.py [case ...]

Each fixture <case>.npz holds the inputs the hot path receives (samples, fs, the
5 ms f0 grid, the 5 ms speech/voiced flags as estimated by the reference's own
voicedUnvoicedFrames, call arguments) and the reference's outputs (s_recon, SRER,
per-frame a0/amplitudes/frange/pk packed dense).  For the synthetic cases the f0
track is supplied by patching ``functions.swipep``; for SA19 the
reference's own SWIPE' output is captured.  ``helpers.npz`` holds input/output
pairs of the three helper seams.  The GPU box never runs this script.
"""
import os
import sys
import tempfile
import time

import numpy as np
from scipy.io import wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
from eaqhm_b200 import synthetic  # noqa: E402


def pack_det(det, Kmax):
    n = len(det)
    out = dict(ti=np.zeros(n, np.int64), isSpeech=np.zeros(n, bool), isVoiced=np.zeros(n, bool),
               a0=np.zeros(n), am=np.zeros((n, Kmax)), fm=np.zeros((n, Kmax)), ph=np.zeros((n, Kmax)),
               nk=np.zeros(n, np.int64))

    def dense(v):
        return np.array([float(np.asarray(x).reshape(-1)[0]) for x in v])

    for i, d in enumerate(det):
        out["ti"][i] = d.ti
        out["isSpeech"][i] = d.isSpeech
        out["isVoiced"][i] = d.isVoiced
        if d.isVoiced:
            out["a0"][i] = float(np.asarray(d.a0).reshape(-1)[0])
            a = dense(d.amplitudes)
            out["nk"][i] = len(a)
            out["am"][i, :len(a)] = a
            out["fm"][i, :len(a)] = dense(d.frange)
            out["ph"][i, :len(a)] = dense(d.pk)
    return out


def run_reference(wav_path, gender, f0s_5ms=None, **kw):
    ref = ref_shim.load_reference()
    cap = {}
    orig_gl, orig_vuv, orig_swipe = ref.getLinear, ref.voicedUnvoicedFrames, ref.swipep

    def gl(v, t):
        out = orig_gl(v, t)
        if isinstance(t, np.ndarray):
            cap["f0s"] = np.array(out, float)
        return out

    def vuv(s, fs, g):
        frames, fstep = orig_vuv(s, fs, g)
        cap["speech5"] = np.array([bool(f.isSpeech) for f in frames])
        cap["voiced5"] = np.array([bool(f.isVoiced) for f in frames])
        cap["frame_step"] = int(fstep)
        cap["frames"] = frames
        return frames, fstep

    ref.getLinear, ref.voicedUnvoicedFrames = gl, vuv
    if f0s_5ms is not None:
        ref.swipep = lambda x, fs, speechFile, plim: np.array(f0s_5ms, float)
    try:
        t0 = time.time()
        s_recon, SRER, det, _ = ref.eaQHMAnalysisAndSynthesis(wav_path, gender, printPrompts=False,
                                                              loadingScreen=False, **kw)
        cap["ref_seconds"] = time.time() - t0
    finally:
        ref.getLinear, ref.voicedUnvoicedFrames, ref.swipep = orig_gl, orig_vuv, orig_swipe
    cap["voiced5_final"] = np.array([bool(f.isVoiced) for f in cap["frames"]])
    cap["speech5_final"] = np.array([bool(f.isSpeech) for f in cap["frames"]])
    del cap["frames"]
    return s_recon, SRER, det, cap


def gender_limits(g):
    if isinstance(g, tuple):
        return g
    return {"male": (70, 180), "female": (160, 300), "child": (300, 600)}.get(g, (70, 500))


def make_case(name, pcm, fs, gender, f0s_5ms=None, subsample=1, **kw):
    with tempfile.TemporaryDirectory() as td:
        wav = os.path.join(td, name + ".wav")
        wavfile.write(wav, fs, pcm)
        s_recon, SRER, det, cap = run_reference(wav, gender, f0s_5ms, **kw)
    f0s = cap["f0s"]
    Fmax = int(fs / 2 - 200)
    Kmax = kw.get("partials", 0) or int(round(Fmax / min(f0s[:, 1])) + 10)
    packed = pack_det(det, Kmax)
    sel = np.arange(0, len(det), subsample)
    out = dict(pcm=pcm, fs=fs, f0s=f0s, speech5=cap["speech5"], voiced5=cap["voiced5"],
               voiced5_final=cap["voiced5_final"], speech5_final=cap["speech5_final"],
               frame_step=cap["frame_step"], f0min=gender_limits(gender)[0], f0max=gender_limits(gender)[1],
               gender=np.array(gender if isinstance(gender, str) else "tuple"),
               Kmax=Kmax, kw_keys=np.array(list(kw.keys())), kw_vals=np.array([float(v) for v in kw.values()]),
               s_recon=s_recon, SRER=np.array(SRER, float), frame_sel=sel, ref_seconds=cap["ref_seconds"])
    for k, v in packed.items():
        out["det_" + k] = v[sel]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "SRER", SRER, "ref_seconds %.1f" % cap["ref_seconds"], flush=True)


def synth_pcm(cfg):
    s, f0s = synthetic.make(cfg)
    return np.round(s * 32768).astype(np.int16), f0s


def case_synth_a(maxAdpt):
    cfg = dict(fs=16000, duration=0.4, seed=11, f0_of_t=lambda t: 200 + 25 * np.sin(2 * np.pi * 3 * t + 0.3))
    pcm, f0s = synth_pcm(cfg)
    make_case("synth_a_adpt%d" % maxAdpt, pcm, 16000, (90, 250), f0s, maxAdpt=maxAdpt)


def case_synth_low():
    cfg = dict(fs=16000, duration=0.3, seed=12, f0_of_t=lambda t: 100 + 8 * np.sin(2 * np.pi * 2 * t))
    pcm, f0s = synth_pcm(cfg)
    make_case("synth_low", pcm, 16000, (70, 180), f0s, maxAdpt=2)


def case_synth_child_voiced_only():
    cfg = synthetic.config_c5(0, duration=1.2)
    pcm, f0s = synth_pcm(cfg)
    make_case("synth_child_vo", pcm, 16000, (250, 500), f0s, maxAdpt=4, fullWaveform=False)


def case_synth_48k():
    cfg = synthetic.config_c4(duration=0.075)
    pcm, f0s = synth_pcm(cfg)
    make_case("synth_48k", pcm, 48000, (70, 100), f0s, maxAdpt=1, analysisWindow=64)


def case_c2_2s():
    """First 2 s of the c2 signal (f0 sweeping 90-250 Hz): the streamed long-signal engine is compared with this
    reference run in two or more frame-range shards (tests/test_gpu_parity.py)."""
    cfg = synthetic.config_c2(duration=2.0)
    pcm, f0s = synth_pcm(cfg)
    make_case("c2_2s", pcm, 16000, (90, 250), f0s, subsample=3, maxAdpt=2)


def case_c4_015():
    """0.15 s of the c4 configuration (48 kHz, f0 ~ 80 Hz, P = 1,190), two adaptive passes."""
    cfg = synthetic.config_c4(duration=0.15)
    pcm, f0s = synth_pcm(cfg)
    make_case("c4_015", pcm, 48000, (70, 100), f0s, subsample=2, maxAdpt=2, analysisWindow=64)


def silence_pcm():
    """Harmonic signal with two stretches of digital silence: frames whose window is all zeros accept nothing at
    adaptation 0 and take the 140 Hz fallback afterwards (functions.py:204-242, quirks Q6/Q10)."""
    cfg = dict(fs=16000, duration=0.6, seed=21, f0_of_t=lambda t: 210 + 15 * np.sin(2 * np.pi * 2.5 * t))
    pcm, f0s = synth_pcm(cfg)
    pcm = pcm.copy()
    pcm[2600:4100] = 0
    pcm[6300:7000] = 0
    return pcm, f0s


def case_synth_silence():
    pcm, f0s = silence_pcm()
    make_case("synth_silence", pcm, 16000, (90, 250), f0s, maxAdpt=3)


def case_sa19():
    fs, pcm = wavfile.read(os.path.join(ref_shim.REFERENCE_DIR, "SA19.WAV"))
    make_case("sa19", pcm, fs, "female", None, subsample=4)


def case_helpers():
    """Input/output pairs of the three helper seams (functions.py:420,472,537)."""
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(5)
    fs = 16000
    out = {}
    # iqhmLS
    N, f0, K = 130, 187.3, 30
    n = np.arange(-N, N + 1)
    s = sum((1.0 / k) * np.cos(2 * np.pi * k * f0 * n / fs + rng.uniform(-3, 3)) for k in range(1, K + 1))
    s = s + 1e-3 * rng.standard_normal(len(n))
    f0range = np.arange(-K, K + 1) * f0
    win = np.blackman(2 * N + 1)
    a, b = ref.iqhmLS_complexamps(s[:, None], f0range, win, fs)
    out.update(iq_s=s, iq_f0range=f0range, iq_win=win, iq_amp=a[:, 0], iq_slope=b[:, 0])
    # eaqhmLS: chirped tracks with AM, mirrored the way functions.py:284-285 does
    nn = len(n)
    Kp = 20
    fm = np.stack([(k + 1) * (f0 + 0.05 * n + 3e-4 * n ** 2) for k in range(Kp)], axis=1)
    am = np.stack([(1.0 / (k + 1)) * (1 + 0.002 * n * np.cos(k)) for k in range(Kp)], axis=1)
    ph = 2 * np.pi * np.cumsum(fm, axis=0) / fs
    s2 = (2 * am * np.cos(ph - ph[N] + rng.uniform(-3, 3, Kp)[None, :])).sum(axis=1) + 1e-3 * rng.standard_normal(nn)
    z = np.zeros((nn, 1))
    fm_all = np.concatenate((-np.flipud(fm), z, fm), axis=1)
    am_all = np.concatenate((np.flipud(am), z, am), axis=1)
    win2 = np.hamming(nn)
    a, b = ref.eaqhmLS_complexamps(s2[:, None], am_all, fm_all, win2, fs)
    out.update(ea_s=s2, ea_am=am_all, ea_fm=fm_all, ea_win=win2, ea_amp=a[:, 0], ea_slope=b[:, 0])
    # phase integration
    length = 400
    idx = np.arange(45, 300, 15)
    fmr = 2 * np.pi / fs * (1500 + 40 * np.sin(np.arange(length) / 30.0))
    phr = np.zeros(length)
    phr[idx] = rng.uniform(-np.pi, np.pi, len(idx))
    pm = ref.phase_integr_interpolation(fmr, phr, idx)
    out.update(pi_fm=fmr, pi_ph=phr, pi_idx=idx, pi_out=pm)
    np.savez_compressed(os.path.join(HERE, "helpers.npz"), **out)
    print("helpers done", flush=True)


CASES = {
    "helpers": case_helpers,
    "synth_a0": lambda: case_synth_a(0), "synth_a1": lambda: case_synth_a(1),
    "synth_a2": lambda: case_synth_a(2), "synth_a3": lambda: case_synth_a(3),
    "synth_low": case_synth_low, "synth_child_vo": case_synth_child_voiced_only,
    "synth_48k": case_synth_48k, "c2_2s": case_c2_2s, "c4_015": case_c4_015, "synth_silence": case_synth_silence, "sa19": case_sa19,
}

if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        CASES[c]()
