"""Tensor form of the eaQHM result, its compact on-disk format and resynthesis from parameters.

The reference returns a Python list of ``Deterministic`` objects whose fields are object-dtype arrays of one-element
arrays (structs.py:7-22, functions.py:404-411, quirk Q8), and its only way back to a waveform is the analysis itself
(main.py:72 writes ``s_recon``).  Here the per-frame parameters stay dense ``(No_ti, Kmax)`` arrays
(``engine.Result``; ``api.pack_det_components`` converts to the reference's records on demand),

* ``save_result`` / ``load_result`` store them as one compressed ``.npz``: flags bit-packed, the accepted harmonics
  of the voiced frames in CSR form (frame pointer, harmonic index, amplitude, frequency, phase), the DC values, the
  SRER list and the call parameters - about 3 x 8 bytes per accepted (frame, harmonic) before compression;
* ``resynthesize`` turns parameters back into the waveform with the SAME device kernels the analysis loop uses for
  its interpolation and synthesis stages (eaqhm_interpolate, eaqhm_synthesize: functions.py:337-388), so the result
  equals the analysis' own ``s_recon`` bit for bit (except on signals that used the 140 Hz fallback, whose returned
  amplitudes carry the reference's aliasing artefact Q10 that its own ``s_recon`` does not contain).
"""
import numpy as np
import torch

from .engine import Engine, Result, Utterance
from .host import FrameTables

FORMAT_VERSION = 1


def save_result(path, res, include_waveform=False):
    """Write an engine.Result as a compact .npz (see the module docstring)."""
    v = np.asarray(res.isVoiced, bool)
    am = np.asarray(res.am)[v]
    nz = am != 0
    ptr = np.concatenate(([0], np.cumsum(nz.sum(axis=1)))).astype(np.int64)
    out = dict(version=FORMAT_VERSION, fs=int(res.fs), step=int(res.step), length=int(len(res.s_recon)),
               Kmax=int(res.am.shape[1]), n_frames=int(len(res.ti)), SRER=np.asarray(res.SRER, float),
               n_adapt_best=int(res.n_adapt_best),
               speech_bits=np.packbits(np.asarray(res.isSpeech, bool)), voiced_bits=np.packbits(v),
               a0=np.asarray(res.a0)[v], ptr=ptr, k=np.nonzero(nz)[1].astype(np.uint16),
               am=am[nz], fm=np.asarray(res.fm)[v][nz], ph=np.asarray(res.ph)[v][nz])
    if include_waveform:
        out["s_recon"] = np.asarray(res.s_recon)
    np.savez_compressed(path, **out)


def load_result(path):
    """Read a file written by save_result back into an engine.Result (s_recon zeros unless it was stored)."""
    z = np.load(path, allow_pickle=False)
    if int(z["version"]) != FORMAT_VERSION:
        raise ValueError("unknown eaQHM parameter file version %d" % int(z["version"]))
    n, K, length, step = int(z["n_frames"]), int(z["Kmax"]), int(z["length"]), int(z["step"])
    speech = np.unpackbits(z["speech_bits"])[:n].astype(bool)
    voiced = np.unpackbits(z["voiced_bits"])[:n].astype(bool)
    am, fm, ph = (np.zeros((n, K)) for _ in range(3))
    rows = np.repeat(np.flatnonzero(voiced), np.diff(z["ptr"]))
    cols = z["k"].astype(np.int64)
    am[rows, cols], fm[rows, cols], ph[rows, cols] = z["am"], z["fm"], z["ph"]
    a0 = np.zeros(n)
    a0[voiced] = z["a0"]
    s = z["s_recon"] if "s_recon" in z.files else np.zeros(length)
    return Result(s_recon=s, SRER=[float(x) for x in z["SRER"]], ti=np.arange(0, length - 1, step)[:n],
                  isSpeech=speech, isVoiced=voiced, a0=a0, am=am, fm=fm, ph=ph, n_adapt_best=int(z["n_adapt_best"]),
                  fs=int(z["fs"]), step=step)


def _bare_tables(length, step):
    """Frame grid of a signal without any analysis quantities (no frame is analysed)."""
    ft = object.__new__(FrameTables)
    ft.length, ft.step = int(length), int(step)
    ti = np.arange(1, length, step)
    n = len(ti)
    ft.No_ti = ft.No_ti_total = n
    ft.g_base = 0
    ft.isSpeech, ft.analysed = np.zeros(n, bool), np.zeros(n, bool)
    ft.f0, ft.K0, ft.N = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
    ft.ti0 = ti - 1
    ft.f0stale = 0.0
    ft.own = np.full(n, step, np.int32)
    ft.own[-1] = length - ft.ti0[-1]
    return ft


def resynthesize(res, device="cuda"):
    """Waveform from per-frame parameters: interpolation between the frame centres (run detection, linear amplitude,
    cubic frequency, integrated phase: functions.py:337-383) and s = a0 + 2 sum_k am cos(ph) (functions.py:385) on
    the GPU.  `res`: engine.Result (or anything with ti, a0, am, fm, ph, fs, step and the signal length in
    len(s_recon)).  Returns a float64 array of that length."""
    length, step, fs = int(len(res.s_recon)), int(res.step), int(res.fs)
    if fs <= 0:
        raise ValueError("the result carries no sample rate")
    n, K = res.am.shape
    ft = _bare_tables(length, step)
    if ft.No_ti != n:
        raise ValueError("parameters have %d frames, a signal of %d samples at step %d has %d" % (n, length, step, ft.No_ti))
    u = Utterance(s=np.zeros(length), fs=fs, f0s=None, voiced5=None, frame_step=1, f0min=0.0, Kmax=int(K))
    eng = Engine([u], step=step, maxAdpt=0, device=device, frame_tables=[ft], chunk=1)
    dev = eng.device
    for i, x in enumerate((res.am, res.ph, res.fm)):
        eng.kn[i].view(K, n).copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(x, float).T)).to(dev))
    eng.kn[3].copy_(torch.from_numpy(np.ascontiguousarray(res.a0, float)).to(dev))
    eng.set_active([True])
    eng.interpolate()
    eng.synthesize()
    return eng.s_hat.cpu().numpy().copy()
