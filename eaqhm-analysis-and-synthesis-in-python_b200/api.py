"""Drop-in entry points with the reference's call signature (functions.py:35-37,418).

``eaQHMAnalysisAndSynthesis`` keeps the reference's arguments and its four return values
(s_recon, SRER, DetComponents, endTime).  Host preprocessing (wav read, optional high-pass,
SWIPE' f0, 5 ms resampling, speech/voiced flags) stays on the CPU (``preprocess.py``); the
adaptation loop (functions.py:163-402) runs on the GPU through ``engine.Engine``.
"""
from time import time, strftime, gmtime

import numpy as np

from . import host
from .engine import Engine, Utterance
from .structs import Deterministic


def pack_det_components(res):
    """functions.py:151,325-329,404-411: list of Deterministic records, one per frame.

    The one-element arrays of all frames (quirk Q8: every accepted harmonic's amplitude, frequency and phase is its own
    array object) are created in three passes over one (n, 1) array each instead of frame by frame; a frame's object
    array is a slice of them (or a scatter, when a harmonic below the last accepted one was dropped)."""
    ti, sp, vo = res.ti.tolist(), res.isSpeech.tolist(), res.isVoiced.tolist()
    vmask = np.asarray(res.isVoiced, bool)
    rows, cols = np.nonzero(res.am[vmask])
    counts = np.bincount(rows, minlength=int(vmask.sum()))
    if (counts == 0).any():
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")   # misc.py:61-62 via functions.py:409
    offs = np.concatenate(([0], np.cumsum(counts)))
    last = cols[offs[1:] - 1].tolist() if len(counts) else []
    offs = offs.tolist()

    def entries(x):
        v = np.ascontiguousarray(x[vmask][rows, cols], dtype=float).reshape(-1, 1)
        return np.fromiter(v, dtype=object, count=len(v))       # the rows of v as one-element float arrays
    va, vf, vp = entries(res.am), entries(res.fm), entries(res.ph)
    out, j = [], 0
    for i in range(len(ti)):
        d = Deterministic(ti=ti[i], isSpeech=sp[i], isVoiced=vo[i])
        if vo[i]:
            o0, o1, size = offs[j], offs[j + 1], last[j] + 1
            d.a0 = res.a0[i]
            if size == o1 - o0:                                   # dense up to the last accepted harmonic
                d.amplitudes, d.frange, d.pk = va[o0:o1].copy(), vf[o0:o1].copy(), vp[o0:o1].copy()
            else:                                                 # int 0 at the dropped harmonics (misc.py:89-93)
                c = cols[o0:o1]
                for name, v in (("amplitudes", va), ("frange", vf), ("pk", vp)):
                    r = np.zeros(size, dtype=object)
                    r[c] = v[o0:o1]
                    setattr(d, name, r)
            j += 1
        out.append(d)
    return out


def make_utterance(s, fs, f0s, speech5, voiced5, frame_step, gender="other", step=15, analysisWindow=32,
                   fullWaveform=True, partials=0):
    """Mode handling of functions.py:95-146 on array inputs -> engine.Utterance."""
    s = np.asarray(s, float).reshape(-1)
    f0min, _ = host.gender_limits(gender)
    _, Kmax = host.harmonic_capacity(fs, np.asarray(f0s, float), partials)
    if fullWaveform:
        speech5, voiced5 = host.full_waveform_flags(speech5, voiced5, len(s), frame_step, analysisWindow * step)
        det = None
    else:
        det = host.voiced_only_target(s, speech5, voiced5, frame_step)
    return Utterance(s=s, fs=int(fs), f0s=np.asarray(f0s, float), voiced5=np.asarray(voiced5, bool),
                     frame_step=int(frame_step), f0min=float(f0min), Kmax=int(Kmax), det_part=det)


STREAM_THRESHOLD_BYTES = 8 << 30     # dense tracks (5 x Kmax x length doubles) above this are streamed shard by shard
MONOLITHIC_LIMIT_BYTES = 100 << 30   # ... unless the signal has wrapping windows and still fits one GPU


class PathHandle:
    """The resident form of one analyze_arrays call: inputs on the device, engine built.  ``run_device`` repeats the
    adaptation loop on the resident inputs (what bench.py times as `value`), ``results`` copies one run's outputs to
    host arrays; ``analyze_arrays`` is make_engine(...).run()."""

    def __init__(self, eng, single):
        self.eng, self.single = eng, single

    def run_device(self, on_adaptation=None, fixed_adaptations=None, prefetch_results=False):
        cb = on_adaptation
        if cb is not None and not self.single:
            cb = lambda a, SRER: on_adaptation(a, [SRER])           # the long-signal engine keeps one SRER list
        return self.eng.run_device(cb, fixed_adaptations=fixed_adaptations, prefetch_results=prefetch_results)

    def results(self, out):
        r = self.eng.results(*out)
        return r[0] if self.single else r

    def run(self, on_adaptation=None, fixed_adaptations=None):
        self.eng.prepare_readback()
        return self.results(self.run_device(on_adaptation, fixed_adaptations, prefetch_results=True))


def make_engine(s, fs, f0s, speech5, voiced5, frame_step, gender="other", step=15, maxAdpt=10, pitchPeriods=3,
                analysisWindow=32, fullWaveform=True, partials=0, stream_threshold_bytes=None, shard_seconds=20.0,
                rank=0, world=1, dist=None, **engine_kw):
    """Host arrays -> PathHandle.  One GPU and dense tracks that fit: the monolithic engine.  Otherwise (tracks above
    the streaming threshold, or world > 1: this rank analyses its contiguous frame range of the signal and exchanges
    halos with its neighbours through `dist`) the frame-range engine."""
    u = make_utterance(s, fs, f0s, speech5, voiced5, frame_step, gender, step, analysisWindow, fullWaveform, partials)
    limit = STREAM_THRESHOLD_BYTES if stream_threshold_bytes is None else stream_threshold_bytes
    stream = world > 1 or 5 * 8 * u.Kmax * len(u.s) > limit
    ft = None
    if stream and world == 1:
        # a window that wraps to the end of the signal (functions.py:196 with a negative start index) is only
        # reproduced by the monolithic engine: keep such signals there as long as their tracks fit at all
        ft = host.FrameTables(len(u.s), u.fs, u.f0s, u.voiced5, u.frame_step, u.Kmax, step, pitchPeriods, analysisWindow)
        if len(ft.wrapping_frames()) and 5 * 8 * u.Kmax * len(u.s) <= MONOLITHIC_LIMIT_BYTES:
            stream = False
    if stream:
        from .longsignal import LongSignalEngine
        eng = LongSignalEngine(u.s, u.fs, u.f0s, u.voiced5, u.frame_step, u.f0min, u.Kmax, shard_seconds=shard_seconds,
                               det_part=u.det_part, step=step, maxAdpt=maxAdpt, pitchPeriods=pitchPeriods,
                               analysisWindow=analysisWindow, rank=rank, world=world, dist=dist, full_tables=ft, **engine_kw)
        return PathHandle(eng, single=False)
    eng = Engine([u], step=step, maxAdpt=maxAdpt, pitchPeriods=pitchPeriods, analysisWindow=analysisWindow, **engine_kw)
    return PathHandle(eng, single=True)


def analyze_arrays(s, fs, f0s, speech5, voiced5, frame_step, gender="other", step=15, maxAdpt=10,
                   pitchPeriods=3, analysisWindow=32, fullWaveform=True, partials=0, on_adaptation=None,
                   stream_threshold_bytes=None, shard_seconds=20.0, rank=0, world=1, dist=None,
                   fixed_adaptations=None, **engine_kw):
    """The hot path on array-level inputs; returns an engine.Result.  Signals whose dense
    per-sample tracks would not fit comfortably (the reference's layout needs length x Kmax x 7 doubles,
    functions.py:159-171) go through the streamed long-signal engine, which gives the same result.  With
    world > 1 (one process per GPU, `dist` = torch.distributed) the call is collective: every rank passes the whole
    signal and gets its own contiguous range of the result inside full-length arrays."""
    h = make_engine(s, fs, f0s, speech5, voiced5, frame_step, gender, step, maxAdpt, pitchPeriods, analysisWindow,
                    fullWaveform, partials, stream_threshold_bytes, shard_seconds, rank, world, dist, **engine_kw)
    return h.run(on_adaptation, fixed_adaptations)


def analyze_batch(utterances, step=15, maxAdpt=10, pitchPeriods=3, analysisWindow=32, **engine_kw):
    """A batch of engine.Utterance on one GPU; returns a list of engine.Result."""
    eng = Engine(utterances, step=step, maxAdpt=maxAdpt, pitchPeriods=pitchPeriods, analysisWindow=analysisWindow,
                 **engine_kw)
    return eng.run()


def eaQHMAnalysisAndSynthesis(speechFile: str, gender: str or tuple = 'other', step: int = 15,
                              maxAdpt: int = 10, pitchPeriods: int = 3, analysisWindow: int = 32,
                              fullWaveform: bool = True, fc: int = 0, partials: int = 0,
                              printPrompts: bool = True, loadingScreen: bool = True, *, f0s=None, frames=None,
                              return_tensors: bool = False, host_preprocessing: bool = False):
    """Same call signature and return values as the reference (functions.py:35-81,418).

    Keyword-only extensions: ``f0s`` ((n,3) [t,f0,strength] track, replaces SWIPE') and ``frames``
    ((speech5, voiced5, frame_step), replaces voicedUnvoicedFrames) let a caller supply the
    host preprocessing results (by default both are computed on the CPU by preprocess.py like the
    reference does); ``loadingScreen`` is accepted and ignored (there is no per-frame Python loop);
    ``return_tensors=True`` returns the per-frame parameters as the dense ``engine.Result`` (a0 (No_ti,), am / fm / ph
    (No_ti, Kmax), flags) in place of the list of ``Deterministic`` objects, whose construction costs more than the
    whole GPU path for one utterance (``pack_det_components(result)`` converts later; ``params.save_result`` /
    ``params.resynthesize`` work on the tensor form); SWIPE' and the energy loop of the speech / voiced flags run on the
    GPU (preprocess.swipep_gpu, voicedUnvoicedFrames_gpu) unless ``host_preprocessing=True`` selects their numpy
    restatements (same pitch track and flags).
    """
    from . import preprocess
    startTime = time()
    fs, s = preprocess.read_wav(speechFile)                                   # functions.py:86-87
    if fc > 0:
        s = preprocess.ellipFilter(s, fs, fc)                                 # functions.py:90-91
    f0min, f0max = host.gender_limits(gender)
    grid = np.arange(0, len(s) - 1, round(fs * 5 / 1000)) / fs                # functions.py:113
    if f0s is None:
        swipe = preprocess.swipep if host_preprocessing else preprocess.swipep_gpu
        f0s = swipe(s, fs, [f0min, f0max])                                    # functions.py:111
    f0s = np.asarray(f0s, float)
    if not (len(f0s) == len(grid) and np.array_equal(f0s[:, 0], grid)):       # already on the 5 ms grid?
        f0s = preprocess.getLinear(f0s, grid)
    if frames is None:
        vuv = preprocess.voicedUnvoicedFrames if host_preprocessing else preprocess.voicedUnvoicedFrames_gpu
        speech5, voiced5, frame_step = vuv(s, fs, gender)                     # functions.py:125
    else:
        speech5, voiced5, frame_step = frames

    def prompt(a, SRER):
        if printPrompts:
            print('---- Adaptation No. {} ----\n'.format(a))
            print('\nSRER: {} dB in Adaptation No: {}'.format(SRER[0][a], a))

    res = analyze_arrays(s, fs, f0s, speech5, voiced5, frame_step, gender, step, maxAdpt, pitchPeriods,
                         analysisWindow, fullWaveform, partials, on_adaptation=prompt)
    det = res if return_tensors else pack_det_components(res)
    endTime = time() - startTime
    if printPrompts:
        print('Signal adapted to {} dB SRER'.format(round(max(res.SRER), 6)))
        print('Total Time: {}\n\n'.format(strftime("%H:%M:%S", gmtime(endTime))))
    return res.s_recon, res.SRER, det, endTime
