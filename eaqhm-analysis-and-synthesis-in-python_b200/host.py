"""Host-side setup of the hot path: the O(No_ti) integer/flag logic of functions.py:95-161
and :176-194 (gender limits, Kmax, full-waveform override, voiced-only SRER target, frame
grid, gating, per-frame f0 / K / window length).  Plain numpy; no device work.
"""
import numpy as np

NORMALIZE = 32768      # misc.py:13


def gender_limits(gender):
    """functions.py:95-109."""
    if isinstance(gender, tuple):
        return gender[0], gender[1]
    if gender == "male":
        return 70, 180
    if gender == "female":
        return 160, 300
    if gender == "child":
        return 300, 600
    return 70, 500


def harmonic_capacity(fs, f0s, partials=0):
    """Fmax and Kmax (functions.py:115-120)."""
    Fmax = int(fs / 2 - 200)
    if partials > 0:
        return Fmax, int(partials)
    return Fmax, int(round(Fmax / min(f0s[:, 1])) + 10)


def full_waveform_flags(speech5, voiced5, length, frame_step, analysisWindowSamples):
    """functions.py:140-145: inside the analysable span every 5 ms frame becomes speech+voiced."""
    speech5 = np.array(speech5, bool)
    voiced5 = np.array(voiced5, bool)
    ti5 = np.arange(1, length, frame_step)[:len(speech5)]
    inside = (ti5 > analysisWindowSamples / 2) & (ti5 < length - analysisWindowSamples / 2)
    first = inside & speech5 & ~voiced5
    voiced5 = voiced5 | first
    second = inside & ~speech5 & ~voiced5
    speech5 = speech5 | second
    voiced5 = voiced5 | second
    return speech5, voiced5


def voiced_only_target(s, speech5, voiced5, frame_step):
    """functions.py:127-138: SRER target that keeps only the voiced regions (quirk Q9: a voiced
    region that reaches the last 5 ms frame is never copied)."""
    s = np.asarray(s, float).reshape(-1)
    ss = np.zeros(len(s))
    ti5 = np.arange(1, len(s), frame_step)
    sp = []
    for j in range(len(speech5)):
        if speech5[j] and voiced5[j]:
            sp.append(ti5[j])
        elif sp:
            lo, hi = sp[0] - frame_step, sp[-1] + frame_step + 1
            idx = np.arange(lo, hi)          # negative indices wrap like the reference's fancy index
            ss[idx] = s[idx]
            sp = []
    return ss


class FrameTables:
    """Per-frame quantities of one utterance (functions.py:148-157,176-194).

    frame_range=(g_lo, g_hi) limits the per-frame arrays to that range of the signal's frames (a rank of a multi-GPU
    long-signal run only builds the tables of the range it touches); `g_base` is the signal frame index of array
    element 0, `No_ti_total` the frame count of the whole signal, `No_ti` the number of frames the arrays cover."""

    @staticmethod
    def _values(ti, length, fs, f0s, voiced5, frame_step, Kmax, step, pitchPeriods, analysisWindow):
        Fmax = int(fs / 2 - 200)
        framei = ti / frame_step                                          # functions.py:155
        fi = framei.astype(int)
        dec = framei - fi
        aw = analysisWindow * step
        isSpeech = (ti > aw) & (ti < length - aw)                         # functions.py:180
        nv = len(voiced5)
        analysed = isSpeech & voiced5[(fi - 1) % nv] & voiced5[np.minimum(fi, nv - 1)]   # :181
        idx = np.flatnonzero(analysed)
        f0v = np.zeros(len(ti))
        K0 = np.zeros(len(ti), np.int32)
        N = np.zeros(len(ti), np.int32)
        if len(idx):
            f0 = (1 - dec[idx]) * f0s[fi[idx] - 1, 1] + dec[idx] * f0s[fi[idx], 1]      # functions.py:185
            if not np.all(np.isfinite(f0)) or np.any(f0 <= 0):
                raise ValueError("f0 track has non-finite or non-positive values on analysed frames")
            f0v[idx] = f0
            K0[idx] = np.minimum(Kmax, (Fmax / f0).astype(int))                    # functions.py:187
            N[idx] = np.maximum(120, np.rint((pitchPeriods / 2) * (fs / f0))).astype(np.int32)   # :191
            c = ti[idx] - 1
            if np.any(c + N[idx] >= length):
                raise IndexError("analysis window runs past the end of the signal "
                                 "(the reference raises IndexError at functions.py:196)")
            if np.any(K0[idx] < 1):
                raise ValueError("f0 above Fmax: no harmonic fits")
        return isSpeech, analysed, f0v, K0, N

    def __init__(self, length, fs, f0s, voiced5, frame_step, Kmax, step=15, pitchPeriods=3,
                 analysisWindow=32, frame_range=None):
        if analysisWindow * step < 2 * step:
            raise ValueError("analysisWindow*step must be at least 2*step (the reference's run detection "
                             "and short-run spline assume no frame is analysed at samples 0 and step)")
        f0s = np.asarray(f0s, float)
        voiced5 = np.asarray(voiced5, bool)
        self.length, self.step = int(length), int(step)
        total = len(range(1, self.length, self.step))                     # functions.py:148: ti = arange(1, length, step)
        if total < 4:
            raise ValueError("signal too short: the DC cubic spline needs at least 4 frames (functions.py:340)")
        g_lo, g_hi = (0, total) if frame_range is None else (max(0, int(frame_range[0])), min(total, int(frame_range[1])))
        self.No_ti_total, self.g_base, self.No_ti = total, g_lo, g_hi - g_lo
        args = (self.length, fs, f0s, voiced5, frame_step, Kmax, step, pitchPeriods, analysisWindow)
        ti = 1 + self.step * np.arange(g_lo, g_hi, dtype=np.int64)        # 1-based centres
        self.isSpeech, self.analysed, self.f0, self.K0, self.N = self._values(ti, *args)
        self.ti0 = ti - 1                                                 # 0-based centres (functions.py:325)
        # quirk Q1: f0 of the LAST analysed frame of the whole signal; a range that stops short of the signal's end
        # looks for it from the end in blocks
        self.f0stale = 0.0
        idx = np.flatnonzero(self.analysed)
        if g_hi == total and len(idx):
            self.f0stale = float(self.f0[idx[-1]])
        else:
            hi = total
            while hi > 0:
                lo = max(0, hi - 4096)
                if lo >= g_lo and hi <= g_hi:
                    an, f0b = self.analysed[lo - g_lo:hi - g_lo], self.f0[lo - g_lo:hi - g_lo]
                else:
                    _, an, f0b, _, _ = self._values(1 + self.step * np.arange(lo, hi, dtype=np.int64), *args)
                j = np.flatnonzero(an)
                if len(j):
                    self.f0stale = float(f0b[j[-1]])
                    break
                hi = lo
        self.own = np.full(self.No_ti, step, np.int32)
        if g_hi == total:
            self.own[-1] = length - self.ti0[-1]

    def wrapping_frames(self):
        """Analysed frames whose window starts before sample 0: the reference's fancy index wraps those samples to
        the END of the signal (functions.py:196,244-245; happens when N > analysisWindow*step, e.g. 44.1/48 kHz
        with f0 below ~150 Hz at the default analysisWindow).  The monolithic engine reproduces that; a frame-range
        excerpt cannot (the tail belongs to another shard), so the streamed engine refuses such signals."""
        return np.flatnonzero(self.analysed & (self.ti0 < self.N))

    def excerpt(self, lo, hi):
        """Tables of the sample range [lo, hi) of this signal (lo a multiple of step), with the SAME per-frame
        values as the whole signal's (f0, K, N, flags are sliced, not recomputed, so a frame-range shard analyses
        bit-identical frames); centres become excerpt-local.  Frames whose window leaves the excerpt keep their
        flags: the caller restricts the analysed list to the frames it owns (longsignal.py)."""
        if lo % self.step or not (0 <= lo < hi <= self.length):
            raise ValueError("excerpt must start on the frame grid and lie inside the signal")
        e = object.__new__(FrameTables)
        e.length, e.step = int(hi - lo), self.step
        g0 = lo // self.step
        n = len(np.arange(1, e.length, self.step))
        if n < 4 or g0 + n > self.No_ti_total:
            raise ValueError("excerpt too short or off the frame grid")
        if g0 < self.g_base or g0 + n > self.g_base + self.No_ti:
            raise ValueError("excerpt outside the frame range these tables were built for")
        e.No_ti, e.g0 = n, int(g0)
        e.No_ti_total, e.g_base = n, 0
        sl = slice(g0 - self.g_base, g0 - self.g_base + n)
        e.isSpeech, e.analysed = self.isSpeech[sl].copy(), self.analysed[sl].copy()
        e.f0, e.K0, e.N = self.f0[sl].copy(), self.K0[sl].copy(), self.N[sl].copy()
        e.ti0 = self.ti0[sl] - lo
        e.f0stale = self.f0stale
        e.own = np.full(n, self.step, np.int32)
        e.own[-1] = e.length - e.ti0[-1]
        return e
