"""Preprocessing that feeds the hot path (functions.py:86-125): wav read, optional elliptic high-pass, SWIPE' pitch
track, 5 ms resampling of the f0 track, energy-based speech/voiced flags; outputs are the array-level inputs of
engine.Utterance.  Every stage has a numpy/scipy restatement of the reference (outside the hot path by design,
BASELINE.json north_star); the two that dominate a single call, SWIPE' and the per-frame energy loop of the flags, also
run on the GPU (``swipep_gpu``, ``voicedUnvoicedFrames_gpu``: csrc/preproc.cu), which is what the
drop-in call uses.
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


def ellipFilter_batch_gpu(signals, fs, fc, ftype="highpass", device="cuda"):
    """ellipFilter for a batch of signals on the GPU (csrc/preproc.cu:k_filtfilt_batch: one thread per signal, scipy's
    filtfilt recursion in scipy's operation order): a list of arrays with the same bits as [ellipFilter(s, fs, fc, ftype)
    for s in signals].  Meant for corpora (thousands of utterances per launch); one signal is faster on the host."""
    import ctypes as C
    import torch
    from scipy.signal import ellip, lfilter_zi
    from . import _lib
    L = _lib.lib()
    if not torch.cuda.is_available():
        raise _lib.EaqhmError("no CUDA device: ellipFilter_batch_gpu has no CPU fallback (use ellipFilter)")
    dev = torch.device(device)
    b, a = ellip(6, .5, 60, 2 * fc / fs, ftype)
    if a[0] != 1.0:
        b, a = b / a[0], a / a[0]
    zi = lfilter_zi(b, a)
    ntaps = max(len(a), len(b))
    padlen = 3 * ntaps
    lens = np.array([len(s) for s in signals], np.int64)
    if np.any(lens <= padlen):
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % padlen)
    start = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.int64)
    wstart = np.concatenate(([0], np.cumsum(lens + 2 * padlen)[:-1])).astype(np.int64)
    with torch.cuda.device(dev):
        up = lambda v, dt: torch.from_numpy(np.ascontiguousarray(v, dtype=dt)).to(dev)
        x = up(np.concatenate([np.asarray(s, float).reshape(-1) for s in signals]), np.float64)
        y = torch.empty_like(x)
        work = torch.empty(int((lens + 2 * padlen).sum()), dtype=torch.float64, device=dev)
        args = [up(start, np.int64), up(lens, np.int64), up(b, np.float64), up(a, np.float64), up(zi, np.float64),
                up(wstart, np.int64)]
        _lib.check(L.eaqhm_filtfilt_batch(x.data_ptr(), args[0].data_ptr(), args[1].data_ptr(), len(signals),
                                          args[2].data_ptr(), args[3].data_ptr(), args[4].data_ptr(), ntaps, padlen,
                                          y.data_ptr(), work.data_ptr(), args[5].data_ptr(),
                                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        out = y.cpu().numpy()
    return [out[o:o + n].copy() for o, n in zip(start, lens)]


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
    # one vectorised median over the L x 4 tap matrix (an hour of signal has 720,000 rows)
    from scipy.linalg import toeplitz
    A = np.fliplr(toeplitz(np.flipud(xp[0:L]), xp[L:L + 4]))
    return np.median(A, axis=1)


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


def voicedUnvoicedFrames_gpu(s, fs, gender, device="cuda"):
    """voicedUnvoicedFrames with the per-frame energy loop (functions.py:618-634: two numpy.std of a 30 ms window per 5 ms
    frame, a Python loop in the reference) on the GPU (csrc/preproc.cu:k_vuv_energy); the two zero-phase elliptic filters
    (misc.py:167-182, recursive) and the 4-tap median of the flags stay on the host.  Same flags as the CPU path (the
    energies differ by the rounding of the sums, the thresholds are whole decibels)."""
    import ctypes as C
    import torch
    from . import _lib
    L = _lib.lib()
    if not torch.cuda.is_available():
        raise _lib.EaqhmError("no CUDA device: voicedUnvoicedFrames_gpu has no CPU fallback (use voicedUnvoicedFrames)")
    dev = torch.device(device)
    s = ellipFilter(np.asarray(s, float).reshape(-1), fs, 30)
    s_smooth = ellipFilter(s, fs, 1000 if gender == "male" else 1500, 'lowpass')
    windowLen = int(round(0.03 * fs))
    if windowLen % 2 == 0:
        windowLen += 1
    step = int(round(0.005 * fs))
    n = len(range(1, len(s), step))
    with torch.cuda.device(dev):
        sd = torch.from_numpy(np.ascontiguousarray(s)).to(dev)
        smd = torch.from_numpy(np.ascontiguousarray(s_smooth)).to(dev)
        en = torch.empty((2, n), dtype=torch.float64, device=dev)
        _lib.check(L.eaqhm_vuv_energy(sd.data_ptr(), smd.data_ptr(), len(s), windowLen, step, n, en[0].data_ptr(),
                                      en[1].data_ptr(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        spEn, spEn_smooth = en.cpu().numpy()
    with np.errstate(all="ignore"):
        isSpeech = spEn > -60                                   # NaN (window outside the signal) compares False
        isVoiced = isSpeech & (spEn - spEn_smooth < 10) & (spEn_smooth > -50)
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


_SWIPE_GRIDS = {}


def _swipe_grids(fs, plim, dlog2p, dERBs):
    """Everything of SWIPE' that depends on the sample rate and the pitch range only (SWIPE.py:33-53,74-90,166-184):
    candidates, window sizes, ERB grid and, per window size, the candidate rows it serves, their weights and their
    kernels.  Cached: building the kernels is most of a CPU call's time."""
    key = (float(fs), float(plim[0]), float(plim[-1]), float(dlog2p), float(dERBs))
    g = _SWIPE_GRIDS.get(key)
    if g is not None:
        return g
    dc, K = 4, 2
    log2pc = np.arange(np.log2(plim[0]), np.log2(plim[-1]), dlog2p)
    pc = np.power(2, log2pc)
    logWs = np.round(np.log2(4 * K * (float(fs) / np.asarray(plim, float))))
    ws = np.power(2, np.arange(logWs[0], logWs[1] - 1, -1))
    pO = 4 * K * fs / ws
    d = 1 + log2pc - np.log2(4 * K * (fs / ws[0]))
    fERBs = _erbs2hz(np.arange(_hz2erbs(pc[0] / 4), _hz2erbs(fs / 2), dERBs))
    wins = []
    for i in range(len(ws)):
        w_i = int(ws[i])
        dn = round(dc * fs / pO[i])
        o = int(max(0, round(w_i - dn)))
        if i == len(ws) - 1:
            j = np.flatnonzero(d - (i + 1) > -1)
            k = np.flatnonzero(d[j] - (i + 1) < 0)
        elif i == 0:
            j = np.flatnonzero(d - (i + 1) < 1)
            k = np.flatnonzero(d[j] - (i + 1) > 0)
        else:
            j = np.flatnonzero(np.abs(d - (i + 1)) < 1)
            k = np.arange(len(j))
        mu = np.ones(len(j))
        mu[k] = 1 - np.abs(d[j[k]] - (i + 1))
        kern = np.zeros((len(j), len(fERBs)))
        for r, cand in enumerate(pc[j][:-1]):                        # SWIPE.py:146: the last candidate is skipped
            kern[r] = _candidate_kernel(fERBs, cand)
        f = np.fft.rfftfreq(w_i, 1.0 / fs)
        lo = np.clip(np.searchsorted(f, fERBs, side="right") - 1, 0, len(f) - 2)
        hann = np.hanning(w_i)
        wins.append(dict(ws=w_i, dn=int(dn), noverlap=o, j=j, mu=mu, kern=kern, f=f, lo=lo, hann=hann,
                         sumw2=float((np.abs(hann) ** 2).sum())))
    g = dict(pc=pc, log2pc=log2pc, fERBs=fERBs, wins=wins)
    _SWIPE_GRIDS[key] = g
    return g


def _swipe_time_tables(ti, t):
    """SWIPE.py:92-100: linear interpolation of the per-frame strengths onto the output times; frame times shifted as the
    port does (a leading 0, the last one dropped).  Returns (inside mask, hi index, t - tim[hi-1], tim[hi] - tim[hi-1])."""
    tim = np.concatenate(([0.0], ti[:-1]))
    inside = (t >= tim[0]) & (t <= tim[-1])
    hi = np.clip(np.searchsorted(tim, t[inside], side="left"), 1, len(tim) - 1)
    l0 = hi - 1
    return inside, hi, t[inside] - tim[l0], tim[hi] - tim[l0]


def _swipe_refine(t, pc, best, y, s_best):
    """Parabolic refinement around the strongest candidate of every frame (SWIPE.py:106-131), all frames that share a
    winning candidate at once: the parabola through the three strengths y = S[best-1 .. best+1] (what polyfit(.., 2)
    returns for three points, here in Newton form) evaluated by Horner on the fine pitch grid of that candidate."""
    p = np.full(len(t), np.nan)
    s = np.array(s_best, float)
    edge = (best == 0) | (best == len(pc) - 1)
    p[edge] = pc[0]
    for i in np.unique(best[~edge]):
        cols = np.flatnonzero(best == i)
        I = np.arange(i - 1, i + 2)
        tc = 1 / pc[I]
        x = ((tc / tc[1]) - 1) * 2 * np.pi
        yy = y[:, cols]                                              # (3, frames)
        d01 = (yy[1] - yy[0]) / (x[1] - x[0])
        d12 = (yy[2] - yy[1]) / (x[2] - x[1])
        c2 = (d12 - d01) / (x[2] - x[0])
        c1 = d01 - c2 * (x[0] + x[1])
        c0 = yy[0] - x[0] * (d01 - c2 * x[1])
        ftc = 1 / np.power(2, np.arange(np.log2(pc[I[0]]), np.log2(pc[I[2]]), 0.0013021))
        nftc = ((ftc / tc[1]) - 1) * 2 * np.pi
        vals = (c2[None, :] * nftc[:, None] + c1[None, :]) * nftc[:, None] + c0[None, :]      # Horner, like polyval
        s[cols] = vals.max(axis=0)
        km = np.argmax(vals, axis=0)
        for kk in np.unique(km):                                      # scalar power like SWIPE.py:131 (same last bit)
            p[cols[km == kk]] = 2 ** (np.log2(pc[I[0]]) + (int(kk) - 1) / 768)
    return p, s


def swipep(x, fs, plim, dt=0.001, dlog2p=1.0 / 96.0, dERBs=0.1):
    """SWIPE' pitch track (Camacho's sawtooth-waveform-inspired estimator) with the behaviour of the port the
    reference ships (SWIPE.py:14-132): returns (n,3) rows [t, pitch, strength] every `dt` seconds.  Restated
    quirks: the last pitch candidate never receives a strength (SWIPE.py:146), a maximum at either end of the
    candidate range reports the LOWEST candidate (SWIPE.py:115-118), 1 is treated as a prime harmonic."""
    x = np.asarray(x, float).reshape(-1)
    t = np.arange(0, len(x) / float(fs), dt)
    g = _swipe_grids(fs, plim, dlog2p, dERBs)
    pc, fERBs = g["pc"], g["fERBs"]
    S = np.zeros((len(pc), len(t)))
    for i, w in enumerate(g["wins"]):
        w_i, dn, j = w["ws"], w["dn"], w["j"]
        xzp = np.concatenate((np.zeros(int(w_i / 2)), x, np.zeros(int(dn + w_i / 2))))
        X, f, ti = _psd_spectrogram(xzp, w_i, fs, w["hann"], w["noverlap"])
        # linear interpolation of every frame's spectrum onto the ERB grid, loudness = sqrt(magnitude)
        lo = w["lo"]
        slope = (X[lo + 1, :] - X[lo, :]) / (f[lo + 1] - f[lo])[:, None]
        Lmat = np.sqrt(np.maximum(0, slope * (fERBs - f[lo])[:, None] + X[lo, :]))
        nrm = np.sqrt((Lmat * Lmat).sum(axis=0))
        nrm[nrm == 0] = np.inf
        Ln = Lmat / nrm[None, :]
        Si = np.zeros((len(j), Ln.shape[1]))
        for r in range(len(j) - 1):                                  # SWIPE.py:146: the last candidate is skipped
            Si[r, :] = w["kern"][r] @ Ln
        if Si.shape[1] > 1:
            inside, hi, w1, w2 = _swipe_time_tables(ti, t)
            l0 = hi - 1
            out = np.full((len(j), len(t)), np.nan)
            sl = (Si[:, hi] - Si[:, l0]) / w2[None, :]
            out[:, inside] = sl * w1[None, :] + Si[:, l0]
            Si = out
        else:
            Si = np.full((len(j), len(t)), np.nan)
        S[j, :] = S[j, :] + w["mu"][:, None] * Si
    best = np.argmax(S, axis=0)
    cols = np.arange(len(t))
    y = np.stack([S[np.maximum(best - 1, 0), cols], S[best, cols], S[np.minimum(best + 1, len(pc) - 1), cols]])
    p, s = _swipe_refine(t, pc, best, y, y[1])
    return np.stack([t, p, s], axis=1)


def swipep_gpu(x, fs, plim, dt=0.001, dlog2p=1.0 / 96.0, dERBs=0.1, device="cuda"):
    """SWIPE' with the pitch-strength matrix built on the GPU (csrc/swipe.cu: hand-written batched FFT -> PSD -> ERB
    loudness, candidate-kernel products, time interpolation and accumulation, numpy-rule argmax); the host keeps what
    depends on (fs, pitch range) only (cached grids and kernels) and the parabolic refinement of the winning candidate,
    shared with ``swipep``.  Same rows [t, pitch, strength]; the pitch equals the CPU path's bit for bit unless two
    candidates tie to the last bits of their strength (the strengths differ by the rounding of the FFT and of the sums)."""
    import ctypes as C
    import torch
    from . import _lib
    L = _lib.lib()
    if not torch.cuda.is_available():
        raise _lib.EaqhmError("no CUDA device: swipep_gpu has no CPU fallback (use swipep)")
    dev = torch.device(device)
    x = np.ascontiguousarray(np.asarray(x, float).reshape(-1))
    t = np.arange(0, len(x) / float(fs), dt)
    g = _swipe_grids(fs, plim, dlog2p, dERBs)
    pc, fERBs = g["pc"], g["fERBs"]
    nt, npc, nerb = len(t), len(pc), len(fERBs)

    def up(a, dtype):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        xd, ferb_d = up(x, np.float64), up(fERBs, np.float64)
        S = torch.zeros((npc, nt), dtype=torch.float64, device=dev)
        for w in g["wins"]:
            w_i, dn, j = w["ws"], w["dn"], w["j"]
            hop = w_i - w["noverlap"]
            nx_zp = int(w_i / 2) + len(x) + int(dn + w_i / 2)
            nseg = (nx_zp - w["noverlap"]) // hop
            if "dev" not in w or w["dev"][0] != str(dev):
                w["dev"] = (str(dev), up(w["hann"], np.float64), up(w["f"], np.float64), up(w["lo"], np.int32),
                            up(w["kern"], np.float64), up(j, np.int32), up(w["mu"], np.float64))
            _, hann_d, f_d, lo_d, kern_d, j_d, mu_d = w["dev"]
            if nseg < 1:
                S[j_d.long()] = float("nan")
                continue
            Ln = torch.empty((nerb, nseg), dtype=torch.float64, device=dev)
            _lib.check(L.eaqhm_swipe_loudness(xd.data_ptr(), len(x), w_i, hop, int(w_i / 2), hann_d.data_ptr(), float(fs),
                                              w["sumw2"], f_d.data_ptr(), lo_d.data_ptr(), ferb_d.data_ptr(), nerb, nseg,
                                              Ln.data_ptr(), st))
            Si = torch.empty((len(j), nseg), dtype=torch.float64, device=dev)
            _lib.check(L.eaqhm_swipe_strength(kern_d.data_ptr(), len(j), nerb, Ln.data_ptr(), nseg, Si.data_ptr(), st))
            ti = np.arange(w_i / 2, nx_zp - w_i / 2 + 1, hop) / fs
            hi = np.full(nt, -1, np.int32)
            w1, w2 = np.zeros(nt), np.ones(nt)
            if nseg > 1:
                inside, h, a1, a2 = _swipe_time_tables(ti, t)
                hi[inside], w1[inside], w2[inside] = h, a1, a2
            hi_d, w1_d, w2_d = up(hi, np.int32), up(w1, np.float64), up(w2, np.float64)
            _lib.check(L.eaqhm_swipe_accumulate(Si.data_ptr(), len(j), nseg, j_d.data_ptr(), mu_d.data_ptr(), hi_d.data_ptr(),
                                                w1_d.data_ptr(), w2_d.data_ptr(), nt, S.data_ptr(), st))
        best_d = torch.empty(nt, dtype=torch.int32, device=dev)
        y_d = torch.empty((3, nt), dtype=torch.float64, device=dev)
        _lib.check(L.eaqhm_swipe_argmax(S.data_ptr(), npc, nt, best_d.data_ptr(), y_d.data_ptr(), st))
        best, y = best_d.cpu().numpy().astype(np.int64), y_d.cpu().numpy()
    p, s = _swipe_refine(t, pc, best, y, y[1])
    return np.stack([t, p, s], axis=1)
