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


def _psd_spectrogram(x, nfft, fs, window, noverlap):
    """One-sided PSD spectrogram without detrending (what matplotlib.pyplot.specgram(mode='psd') returns and
    SWIPE.py:66 consumes): |X|^2 scaled by 1/(fs*sum(w^2)), doubled except at DC and Nyquist; frame times are the
    window centres."""
    hop = nfft - noverlap
    nseg = (len(x) - noverlap) // hop
    idx = np.arange(nfft)[:, None] + hop * np.arange(nseg)[None, :]
    X = np.fft.rfft(x[idx] * window[:, None], n=nfft, axis=0)
    P = (X.conj() * X).real
    if nfft % 2 == 0:
        P[1:-1] *= 2.0
    else:
        P[1:] *= 2.0
    P /= fs
    P /= (np.abs(window) ** 2).sum()
    return P, np.fft.rfftfreq(nfft, 1.0 / fs), np.arange(nfft / 2, len(x) - nfft / 2 + 1, hop) / fs


def _hz2erbs(hz):
    return 21.4 * np.log10(1 + hz / 229)


def _erbs2hz(erbs):
    return (np.power(10, erbs / 21.4) - 1) * 229


def _is_prime(n):
    return all(n % i for i in range(2, int(np.sqrt(n)) + 1))          # SWIPE.py:151-158: 1 counts as prime


def _candidate_kernel(f, pc):
    """SWIPE.py:166-184: normalised cosine kernel of one pitch candidate over the ERB frequency grid."""
    n = np.fix(f[-1] / pc - 0.75)
    k = np.zeros(len(f))
    q = f / pc
    for i in (i for i in range(1, int(n) + 2) if _is_prime(i)):
        a = np.abs(q - i)
        peak = a < .25
        k[peak] = np.cos(2 * np.pi * q[peak])
        valley = (.25 < a) & (a < .75)
        k[valley] = np.cos(2 * np.pi * q[valley]) / 2
    k = k * np.sqrt(1.0 / f)
    return k / np.linalg.norm(k[k > 0.0])


def swipep(x, fs, plim, dt=0.001, dlog2p=1.0 / 96.0, dERBs=0.1):
    """SWIPE' pitch track (Camacho's sawtooth-waveform-inspired estimator) with the behaviour of the port the
    reference ships (SWIPE.py:14-132): returns (n,3) rows [t, pitch, strength] every `dt` seconds.  Restated
    quirks: the last pitch candidate never receives a strength (SWIPE.py:146), a maximum at either end of the
    candidate range reports the LOWEST candidate (SWIPE.py:115-118), 1 is treated as a prime harmonic."""
    x = np.asarray(x, float).reshape(-1)
    t = np.arange(0, len(x) / float(fs), dt)
    dc, K = 4, 2
    log2pc = np.arange(np.log2(plim[0]), np.log2(plim[-1]), dlog2p)
    pc = np.power(2, log2pc)
    S = np.zeros((len(pc), len(t)))
    logWs = np.round(np.log2(4 * K * (float(fs) / np.asarray(plim, float))))
    ws = np.power(2, np.arange(logWs[0], logWs[1] - 1, -1))
    pO = 4 * K * fs / ws
    d = 1 + log2pc - np.log2(4 * K * (fs / ws[0]))
    fERBs = _erbs2hz(np.arange(_hz2erbs(pc[0] / 4), _hz2erbs(fs / 2), dERBs))
    for i in range(len(ws)):
        w_i = int(ws[i])
        dn = round(dc * fs / pO[i])
        xzp = np.concatenate((np.zeros(int(w_i / 2)), x, np.zeros(int(dn + w_i / 2))))
        o = max(0, round(w_i - dn))
        X, f, ti = _psd_spectrogram(xzp, w_i, fs, np.hanning(w_i), int(o))
        # linear interpolation of every frame's spectrum onto the ERB grid, loudness = sqrt(magnitude)
        lo = np.clip(np.searchsorted(f, fERBs, side="right") - 1, 0, len(f) - 2)
        slope = (X[lo + 1, :] - X[lo, :]) / (f[lo + 1] - f[lo])[:, None]
        Lmat = np.sqrt(np.maximum(0, slope * (fERBs - f[lo])[:, None] + X[lo, :]))
        if i == len(ws) - 1:
            j = np.flatnonzero(d - (i + 1) > -1)
            k = np.flatnonzero(d[j] - (i + 1) < 0)
        elif i == 0:
            j = np.flatnonzero(d - (i + 1) < 1)
            k = np.flatnonzero(d[j] - (i + 1) > 0)
        else:
            j = np.flatnonzero(np.abs(d - (i + 1)) < 1)
            k = np.arange(len(j))
        nrm = np.sqrt((Lmat * Lmat).sum(axis=0))
        nrm[nrm == 0] = np.inf
        Ln = Lmat / nrm[None, :]
        Si = np.zeros((len(j), Ln.shape[1]))
        for r, cand in enumerate(pc[j][:-1]):                        # SWIPE.py:146: the last candidate is skipped
            Si[r, :] = _candidate_kernel(fERBs, cand) @ Ln
        if Si.shape[1] > 1:
            tim = np.concatenate(([0.0], ti[:-1]))                   # SWIPE.py:92-97
            out = np.full((len(j), len(t)), np.nan)
            inside = (t >= tim[0]) & (t <= tim[-1])
            hi = np.clip(np.searchsorted(tim, t[inside], side="left"), 1, len(tim) - 1)
            l0 = hi - 1
            sl = (Si[:, hi] - Si[:, l0]) / (tim[hi] - tim[l0])[None, :]
            out[:, inside] = sl * (t[inside] - tim[l0])[None, :] + Si[:, l0]
            Si = out
        else:
            Si = np.full((len(j), len(t)), np.nan)
        mu = np.ones(len(j))
        mu[k] = 1 - np.abs(d[j[k]] - (i + 1))
        S[j, :] = S[j, :] + mu[:, None] * Si
    # parabolic refinement around the strongest candidate of every frame (SWIPE.py:106-131), all frames that share a
    # winning candidate at once: the parabola through the three strengths (what polyfit(.., 2) returns for three points,
    # here in Newton form) evaluated by Horner on the fine pitch grid of that candidate
    p = np.full(len(t), np.nan)
    s = np.full(len(t), np.nan)
    best = np.argmax(S, axis=0)
    s[:] = S[best, np.arange(len(t))]
    edge = (best == 0) | (best == len(pc) - 1)
    p[edge] = pc[0]
    for i in np.unique(best[~edge]):
        cols = np.flatnonzero(best == i)
        I = np.arange(i - 1, i + 2)
        tc = 1 / pc[I]
        x = ((tc / tc[1]) - 1) * 2 * np.pi
        y = S[I][:, cols]                                            # (3, frames)
        d01 = (y[1] - y[0]) / (x[1] - x[0])
        d12 = (y[2] - y[1]) / (x[2] - x[1])
        c2 = (d12 - d01) / (x[2] - x[0])
        c1 = d01 - c2 * (x[0] + x[1])
        c0 = y[0] - x[0] * (d01 - c2 * x[1])
        ftc = 1 / np.power(2, np.arange(np.log2(pc[I[0]]), np.log2(pc[I[2]]), 0.0013021))
        nftc = ((ftc / tc[1]) - 1) * 2 * np.pi
        vals = (c2[None, :] * nftc[:, None] + c1[None, :]) * nftc[:, None] + c0[None, :]      # Horner, like polyval
        s[cols] = vals.max(axis=0)
        km = np.argmax(vals, axis=0)
        for kk in np.unique(km):                                      # scalar power like SWIPE.py:131 (same last bit)
            p[cols[km == kk]] = 2 ** (np.log2(pc[I[0]]) + (int(kk) - 1) / 768)
    return np.stack([t, p, s], axis=1)
