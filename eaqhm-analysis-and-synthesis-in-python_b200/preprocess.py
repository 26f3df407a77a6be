"""Host-side preprocessing that feeds the hot path (functions.py:86-125): wav read, optional elliptic high-pass,
5 ms resampling of the f0 track, energy-based speech/voiced flags.  numpy/scipy on the CPU — outside the hot path
by design (BASELINE.json north_star); outputs are the array-level inputs of engine.Utterance.
"""
import numpy as np

from .host import NORMALIZE


def read_wav(path):
    """functions.py:86-87: int16 samples / 32768 -> float64 in [-1, 1)."""
    from scipy.io import wavfile
    fs, s = wavfile.read(path)
    return int(fs), np.asarray(s, float).reshape(-1) / NORMALIZE


def ellipFilter(s, fs, fc, ftype="highpass"):
    """misc.py:167-182: zero-phase 6th-order elliptic filter (0.5 dB ripple, 60 dB stop-band)."""
    from scipy.signal import ellip, filtfilt
    b, a = ellip(6, .5, 60, 2 * fc / fs, ftype)
    return filtfilt(b, a, s)


def getLinear(v, t):
    """functions.py:644-680 for an array of times: rows [t, linear interpolation of v[:,1:] at t]; times before the
    first row take the first row, times at/after the last row the last one."""
    v = np.asarray(v, float)
    t = np.asarray(t, float)
    times = v[:, 0]
    out = np.ones((len(t), v.shape[1]))
    out[:, 0] = t
    prev = np.searchsorted(times, t, side="right") - 1          # last index with times <= t
    for j in range(1, v.shape[1]):
        col = v[:, j]
        res = np.empty(len(t))
        before = prev < 0
        res[before] = col[0]
        last = prev >= len(v) - 1
        res[last & ~before] = col[-1]
        mid = ~before & ~last
        p = prev[mid]
        exact = times[p] == t[mid]
        g = (t[mid] - times[p]) / (times[p + 1] - times[p])
        if np.any((g < 0) | (g > 1)):
            raise ValueError('linearity factor unbound, g not in [0, 1]')      # functions.py:677
        val = col[p] * (1 - g) + col[p + 1] * g
        val[exact] = col[p][exact]
        res[mid] = val
        out[:, j] = res
    return out


def medfilt5(x):
    """misc.py:184-206 with p = 5 as the reference calls it.  The reference builds its sliding windows with a
    Toeplitz construction that yields FOUR taps per output (x[i-1], x[i], x[i+1], x[i+2] after edge padding by 2),
    so the statistics.median of an even count can be 0.5; both are reproduced."""
    x = np.asarray(x, float)
    L = len(x)
    xp = np.concatenate((np.full(2, x[0]), x, np.full(2, x[L - 1])))
    # A = fliplr(toeplitz(flipud(xp[0:L]), xp[L:L+4])): row i = [xp[L+3], xp[L+2], xp[L+1], c_i] reversed ...
    from scipy.linalg import toeplitz
    A = np.fliplr(toeplitz(np.flipud(xp[0:L]), xp[L:L + 4]))
    return np.array([np.median(a) for a in A])


def voicedUnvoicedFrames(s, fs, gender):
    """functions.py:577-642: speech / voiced flags every 5 ms from the 30 ms frame energy of the signal and of its
    low-passed version.  Returns (isSpeech, isVoiced, frame_step) as arrays (the reference returns Frame objects)."""
    s = ellipFilter(np.asarray(s, float).reshape(-1), fs, 30)
    length = len(s)
    s_smooth = ellipFilter(s, fs, 1000 if gender == "male" else 1500, 'lowpass')
    windowLen = int(round(0.03 * fs))
    if windowLen % 2 == 0:
        windowLen += 1
    step = int(round(0.005 * fs))
    midlen = (windowLen - 1) / 2
    wr = np.arange(-midlen - 1, midlen, dtype=int)
    ti = np.arange(1, length, step)
    isSpeech = np.zeros(len(ti), bool)
    isVoiced = np.zeros(len(ti), bool)
    with np.errstate(all="ignore"):
        for i, tith in enumerate(ti):
            if tith > midlen and tith < length - midlen:
                spEn = 20 * np.log10(np.std(s[tith + wr]))
                spEn_smooth = 20 * np.log10(np.std(s_smooth[tith + wr]))
                isSpeech[i] = spEn > -60
                if isSpeech[i]:
                    isVoiced[i] = (spEn - spEn_smooth < 10) and (spEn_smooth > -50)
    return medfilt5(isSpeech), medfilt5(isVoiced), step


def swipep(x, fs, plim):
    """SWIPE' pitch track (SWIPE.py:14-132 of the reference, a third-party port).  Host preprocessing outside the hot
    path; not restated in round 1 — supply the track through ``f0s=``."""
    raise NotImplementedError("SWIPE' f0 estimation is host preprocessing outside the hot path and is not part of this "
                              "round; pass f0s=(n,3) [t, f0, strength] to eaQHMAnalysisAndSynthesis")
