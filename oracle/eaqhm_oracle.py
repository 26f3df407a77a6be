"""CPU oracle for the eaQHM analysis+synthesis hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain numpy float64, the algorithm of the reference's hot
path (``/root/reference/functions.py:148-411`` plus the three helpers at
``functions.py:420-575``).  It exists so that the CUDA path can be checked
against it; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
never does (and fails loudly when its CUDA library is missing).

Parity pin: the oracle is pinned against
  * the reference's only published vector, the six SA19 SRER values of
    ``img/SA19out.JPG`` (``README.md:46-48``), and
  * outputs of the unmodified reference imported in the build container through
    ``tests/golden/ref_shim.py`` (script ``tests/golden/make_golden.py``; fixtures
    ``tests/golden/*.npz``): per-frame amplitudes/frequencies/phases, waveform
    and SRER per adaptation for SA19 and for short synthetic signals, and
    input/output pairs of the three helper seams.
``tests/test_oracle_golden.py`` runs those checks on CPU.

Third-party arithmetic the reference reaches through numpy/scipy (not under
/root/reference; versions: requirements.txt:125 numpy==1.18.5, :197 scipy 1.5)
is restated from its published definition:
  * ``interp1d(kind='linear')`` on 1-D float data == ``numpy.interp``
    (exact at knots, ``slope*(x-x_lo)+y_lo`` between),
  * ``interp1d(kind=3)`` == ``make_interp_spline(k=3)`` == the unique
    not-a-knot cubic interpolating spline (restated in moment form below),
  * ``lfilter([1],[1,-1],x)`` == sequential ``cumsum``,
  * ``numpy.unwrap`` / ``numpy.std`` (population, ddof 0).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

TWO_PI = 2 * np.pi
NORMALIZE = 32768        # misc.py:13
MIN_INTERP_SIZE = 4      # misc.py:14
AM_EPS = 10e-5           # functions.py:517
FALLBACK_FM = 140.0      # functions.py:209
FALLBACK_AM = 10e-4      # functions.py:210


# ----------------------------------------------------------------------------
# helper seams (same names / argument meaning as the reference)
# ----------------------------------------------------------------------------
def _ls_normal_equations(E2, n, window, s):
    """functions.py:455-468 and :519-533 — joint LS for amplitudes and slopes.

    E2 (L,M) complex, n (L,) centred sample index, window (L,), s (L,) real.
    The window multiplies both the basis and the signal (weight = w^2, Q3) and
    the solve is the explicit inverse times the right-hand side, as in the
    reference.
    """
    E = np.concatenate((E2, n[:, None] * E2), axis=1)
    Ew = E * window[:, None]
    EwH = Ew.conj().T
    R = EwH @ Ew
    b = EwH @ (window * s)
    x = np.linalg.inv(R) @ b
    M = E2.shape[1]
    return x[:M], x[M:]


def iqhmLS_complexamps(s, f0range, window, fs):
    """functions.py:420-470.  s (L,) or (L,1); returns (M,) complex amplitudes, slopes."""
    s = np.asarray(s, float).reshape(-1)
    f0range = np.asarray(f0range, float)
    L = len(s)
    midlen = (L - 1) / 2
    n = np.arange(-midlen, midlen + 1)
    t = (n[:, None] * 2 * np.pi * f0range[None, :]) / fs          # :453 (this operation order)
    E2 = np.cos(t) + 1j * np.sin(t)                                 # :454
    return _ls_normal_equations(E2, n, np.asarray(window, float), s)


def eaqhmLS_complexamps(s, am, fm, window, fs):
    """functions.py:472-535.  am, fm (L,M') per-sample tracks; returns (M',) amplitudes, slopes."""
    s = np.asarray(s, float).reshape(-1)
    L, _ = fm.shape
    midlen = int((L - 1) / 2)
    n = np.arange(-midlen, midlen + 1).astype(float)
    f_an = np.cumsum(fm, axis=0)                                    # :510 lfilter([1],[1,-1]) == cumsum
    f_an = f_an - f_an[midlen][None, :]                             # :511
    t = (2 * np.pi * f_an) / fs                                     # :513
    E1 = np.cos(t) + 1j * np.sin(t)                                 # :516
    E2 = ((AM_EPS + am) / (am[midlen][None, :] + AM_EPS)) * E1      # :518
    return _ls_normal_equations(E2, n, np.asarray(window, float), s)


def phase_integr_interpolation(fm_recon, ph_recon, indices):
    """functions.py:537-575.  fm_recon in rad/sample (full column), ph_recon full column."""
    indices = np.asarray(indices).astype(int)
    out = np.zeros(indices[-1] - indices[0] + 1)
    base = indices[0]
    for i in range(len(indices) - 1):
        i0, i1 = indices[i], indices[i + 1]
        D = i1 - i0
        p = np.cumsum(fm_recon[i0:i1 + 1])                          # :561
        p = p + (ph_recon[i0] - p[0])                               # :562
        Mr = round((p[-1] - ph_recon[i1]) / (2 * np.pi))            # :564 (half-to-even)
        er = np.pi * (p[-1] - ph_recon[i1] - 2 * np.pi * Mr) / (2 * D)   # :565
        ft = np.sin(np.pi * np.arange(0, D + 1) / D)                # :566-567
        p = p - np.cumsum(ft * er)                                  # :569
        out[i0 - base:i1 - base + 1] = p                            # :571 (next segment overwrites the shared end)
    return out


# ----------------------------------------------------------------------------
# third-party routines restated
# ----------------------------------------------------------------------------
def notaknot_cubic(xk, yk, xq):
    """Not-a-knot cubic interpolating spline through (xk, yk) evaluated at xq
    (with extrapolation by the end pieces) — what ``interp1d(kind=3)`` /
    ``make_interp_spline(k=3)`` computes (functions.py:340,367,371).

    Moment form: M_j = S''(x_j); interior equations
    h_{j-1} M_{j-1} + 2(h_{j-1}+h_j) M_j + h_j M_{j+1} = 6 (d_j - d_{j-1}),
    end conditions S''' continuous at x_1 and x_{n-2}.  Needs >= 4 knots.
    """
    xk = np.asarray(xk, float)
    yk = np.asarray(yk, float)
    xq = np.asarray(xq, float)
    n = len(xk)
    if n < 4:
        raise ValueError("cubic not-a-knot spline needs at least 4 knots")
    h = np.diff(xk)
    d = np.diff(yk) / h
    # banded storage (2 sub-, 2 super-diagonals): ab[2 + i - j, j] = A[i, j]
    ab = np.zeros((5, n))
    r = np.zeros(n)
    ab[3, 0:n - 2] = h[0:n - 2]                     # A[j, j-1] = h[j-1], j = 1..n-2
    ab[2, 1:n - 1] = 2 * (h[0:n - 2] + h[1:n - 1])  # A[j, j]
    ab[1, 2:n] = h[1:n - 1]                         # A[j, j+1] = h[j]
    r[1:n - 1] = 6 * (d[1:] - d[:-1])
    ab[2, 0], ab[1, 1], ab[0, 2] = h[1], -(h[0] + h[1]), h[0]                          # row 0
    ab[4, n - 3], ab[3, n - 2], ab[2, n - 1] = h[n - 2], -(h[n - 3] + h[n - 2]), h[n - 3]   # row n-1
    from scipy.linalg import solve_banded            # banded LU of the same linear system
    Mom = solve_banded((2, 2), ab, r)
    j = np.clip(np.searchsorted(xk, xq, side="right") - 1, 0, n - 2)
    hj = h[j]
    t1 = xq - xk[j]
    t0 = xk[j + 1] - xq
    return (Mom[j] * t0 ** 3 + Mom[j + 1] * t1 ** 3) / (6 * hj) \
        + (yk[j] / hj - Mom[j] * hj / 6) * t0 + (yk[j + 1] / hj - Mom[j + 1] * hj / 6) * t1


def fill_track_gaps(fm_col, am_col):
    """functions.py:250-278 (and :215-242) for one harmonic column of a window.

    Where fm == 0: hold the first/last non-zero value out to the window edges and
    interpolate interior gaps linearly; ``am`` is filled at fm's zero positions.
    Every sample goes through numpy.interp like the reference's interp1d.
    """
    L = len(fm_col)
    nz = np.flatnonzero(fm_col)
    fm_col = fm_col.copy()
    am_col = am_col.copy()
    if len(nz) != L:
        if fm_col[0] == 0:
            fm_col[0] = fm_col[nz[0]]
            am_col[0] = am_col[nz[0]]
            nz = np.insert(nz, 0, 0)
        if fm_col[L - 1] == 0:
            fm_col[L - 1] = fm_col[nz[-1]]
            am_col[L - 1] = am_col[nz[-1]]
            nz = np.append(nz, L - 1)
    x_new = np.arange(0, L)
    return np.interp(x_new, nz, fm_col[nz]), np.interp(x_new, nz, am_col[nz])


# ----------------------------------------------------------------------------
# the adaptation loop (functions.py:148-411) on array-level inputs
# ----------------------------------------------------------------------------
@dataclass
class OracleResult:
    s_recon: np.ndarray                 # (length,)
    SRER: list                          # one value per adaptation run
    ti: np.ndarray                      # (No_ti,) 0-based centre sample of every frame (functions.py:325)
    isSpeech: np.ndarray                # (No_ti,) bool
    isVoiced: np.ndarray                # (No_ti,) bool
    a0: np.ndarray                      # (No_ti,) a0_fin at the centres
    am: np.ndarray                      # (No_ti,Kmax) am_fin at the centres
    fm: np.ndarray                      # (No_ti,Kmax) fm_fin at the centres (Hz)
    ph: np.ndarray                      # (No_ti,Kmax) pm_fin at the centres (rad)
    window_lengths: np.ndarray          # (No_ti,) half window N per frame
    n_adapt_best: int = 0               # index of the adaptation whose outputs are returned
    per_adapt: list = field(default_factory=list)   # optional per-adaptation centre params
    n_fallback: int = 0                 # frames that took the 140 Hz fallback (functions.py:204-242)


def frame_gate(length, step, analysisWindow, voiced5, frame_step):
    """functions.py:148-157,176-181.  Returns ti (1-based), fi, dec, analysed mask, speech mask."""
    ti = np.arange(1, length, step)
    framei = ti / frame_step
    fi = framei.astype(int)
    dec = framei - fi
    aw = analysisWindow * step
    speech = (ti > aw) & (ti < length - aw)
    v = np.asarray(voiced5, bool)
    analysed = speech & v[fi - 1] & v[np.minimum(fi, len(v) - 1)]
    return ti, fi, dec, analysed, speech


def eaqhm_oracle(s, fs, f0s, voiced5, frame_step, f0min, Kmax, *, step=15, maxAdpt=10,
                 pitchPeriods=3, analysisWindow=32, det_part=None, keep_per_adapt=False,
                 progress=None, fixed_adaptations=None):
    """Restatement of functions.py:148-411.

    s (length,) float64 in [-1,1); f0s (n5,3) rows [t,f0,strength] on the 5 ms grid
    (functions.py:113); voiced5 (n5',) bool isVoiced flag per 5 ms frame AFTER the
    full-waveform override (functions.py:140-145); det_part = SRER target
    (functions.py:138,146; default s).

    fixed_adaptations (benchmark mode, not in the reference): run exactly that many adaptations without the
    convergence stop of functions.py:394-396, so that two sides do the same work; the returned result is still the
    adaptation the reference would have returned (the last one before the first non-improvement).
    """
    s = np.asarray(s, float).reshape(-1)
    length = len(s)
    det = s if det_part is None else np.asarray(det_part, float).reshape(-1)
    Fmax = int(fs / 2 - 200)                                        # :115
    ti, fi, dec, analysed, speech = frame_gate(length, step, analysisWindow, voiced5, frame_step)
    No_ti = len(ti)
    window_lengths = np.zeros(No_ti, int)
    fm_current = np.zeros((length, Kmax))
    am_current = np.zeros((length, Kmax))
    std_det = np.std(det)                                           # :161
    SRER = []
    f0 = None
    best = None
    per_adapt = []
    frames_idx = np.flatnonzero(analysed)
    n_fallback = 0

    stopped = False
    for a in range(maxAdpt + 1 if fixed_adaptations is None else int(fixed_adaptations)):
        a0_recon = np.zeros(length)
        am_recon = np.zeros((length, Kmax))
        fm_recon = np.zeros((length, Kmax))
        ph_recon = np.zeros((length, Kmax))
        for i in frames_idx:
            tith = ti[i]
            c = tith - 1
            if a == 0:
                f0 = (1 - dec[i]) * f0s[fi[i] - 1][1] + dec[i] * f0s[fi[i]][1]     # :185
                K = int(min(Kmax, int(Fmax / f0)))                                  # :187
                f0range = np.arange(-K, K + 1) * f0                                 # :189
                N = max(120, round((pitchPeriods / 2) * (fs / f0)))                 # :191
                window_lengths[i] = N
                window = np.blackman(2 * N + 1)                                     # :194
                win_idx = np.arange(c - N, c + N + 1)       # fancy index like :192-196: a negative start wraps
                amps, slopes = iqhmLS_complexamps(s[win_idx], f0range, window, fs)
                fmis = np.zeros(2 * K + 1)                                          # :197
            else:
                N = window_lengths[i]
                window = np.hamming(2 * N + 1)                                      # :200
                act = np.flatnonzero(fm_current[c])                                 # :202
                if len(act) == 0:                                                   # :204-242 fallback (Q6)
                    n_fallback += 1
                    fm_current[c, 0] = FALLBACK_FM
                    am_current[c, 0] = FALLBACK_AM
                    act = np.array([0])
                    K = 1
                else:
                    K = act[-1] + 1                                                 # :248
                win_idx = np.arange(c - N, c + N + 1)       # :199,244-245 (fancy index: wraps like the reference)
                fm = fm_current[win_idx][:, act]
                am = am_current[win_idx][:, act]
                for j in range(len(act)):
                    fm[:, j], am[:, j] = fill_track_gaps(fm[:, j], am[:, j])
                zc = np.zeros((2 * N + 1, 1))
                fm_all = np.concatenate((-fm[::-1], zc, fm), axis=1)                # :284 flipud reverses TIME (Q2)
                am_all = np.concatenate((am[::-1], zc, am), axis=1)                 # :285
                amps_t, slopes_t = eaqhmLS_complexamps(s[win_idx], am_all, fm_all, window, fs)
                with np.errstate(all="ignore"):
                    fmis_t = fs / (2 * np.pi) * (amps_t.real * slopes_t.imag - amps_t.imag * slopes_t.real) \
                        / np.abs(amps_t) ** 2                                       # :297
                idxmap = np.concatenate((-(act[::-1] + 1), [0], act + 1)) + K        # :281
                amps = np.zeros(2 * K + 1, complex)
                fmis = np.zeros(2 * K + 1)
                amps[idxmap] = amps_t                                               # :299-301 arrayByIndex
                fmis[idxmap] = fmis_t
            a0_recon[c] = amps[K].real                                              # :303
            amps_p = amps[K + 1:2 * K + 1]
            fmis_p = fmis[K + 1:2 * K + 1]
            with np.errstate(all="ignore"):
                logs = 20 * np.log10(np.abs(amps_p))
                log_max = 20 * np.log10(np.max(np.abs(amps_p))) - 150               # :309
            h = f0 / (a + 1)                                                        # :310 (stale f0 for a>=1, Q1)
            acc = (logs > log_max) & (np.abs(fmis_p) < h)                            # :315
            ks = np.flatnonzero(acc)
            am_recon[c, ks] = np.abs(amps_p[ks])
            ph_recon[c, ks] = np.angle(amps_p[ks])
            if a == 0:
                fm_recon[c, ks] = (ks + 1) * f0                                     # :320
            elif f0 > f0min:
                fm_recon[c, ks] = fm_current[c, ks] + fmis_p[ks]                    # :322
            else:
                fm_recon[c, ks] = fm_current[c, ks]                                 # :324
            if progress is not None:
                progress(a, i)

        fm_current = np.zeros((length, Kmax))                                       # :337
        a0_dense = notaknot_cubic(ti - 1, a0_recon[ti - 1], np.arange(0, length))   # :340

        for k in range(Kmax):                                                       # :346
            nzs = np.flatnonzero(am_recon[:, k])                                    # :350
            if len(nzs) == 0:
                continue
            q = np.concatenate(([0], nzs, [length - 1]))
            flag = (np.diff(q) <= step).astype(int)                                 # :352-354
            dd = np.diff(flag)
            starts = np.flatnonzero(dd == 1)
            ends = np.flatnonzero(dd == -1)
            for r_i, st in enumerate(starts):                                       # :360
                knots = nzs[st:ends[r_i] + 1]
                rng = np.arange(knots[0], knots[-1] + 1)
                am_recon[rng, k] = np.interp(rng, knots, am_recon[knots, k])        # :364
                if len(knots) >= MIN_INTERP_SIZE:
                    fm_recon[rng, k] = notaknot_cubic(knots, fm_recon[knots, k], rng)          # :367
                else:
                    extra = np.arange(0, (MIN_INTERP_SIZE - len(knots)) * step, step)          # :369 (Q4)
                    kn2 = np.concatenate((extra, knots))
                    fm_recon[rng, k] = notaknot_cubic(kn2, fm_recon[kn2, k], rng)              # :371
                ph_recon[rng, k] = phase_integr_interpolation(2 * np.pi / fs * fm_recon[:, k],
                                                              ph_recon[:, k], knots)           # :373
                fm_current[rng, k] = np.concatenate(([fm_recon[rng[0], k]],
                                                     fs / (2 * np.pi) * np.diff(np.unwrap(ph_recon[rng, k]))))  # :375
        am_current = am_recon                                                       # :383 (alias)
        s_tmp = a0_dense + 2 * (am_recon * np.cos(ph_recon)).sum(axis=1)            # :385
        with np.errstate(all="ignore"):
            SRER.append(20 * np.log10(std_det / np.std(det - s_tmp)))               # :388
        if keep_per_adapt:
            c_all = ti - 1
            per_adapt.append(dict(a0=a0_dense[c_all].copy(), am=am_recon[c_all].copy(),
                                  fm=fm_recon[c_all].copy(), ph=ph_recon[c_all].copy(),
                                  s=s_tmp.copy()))
        if a != 0 and SRER[a] <= SRER[a - 1]:                                       # :394-396
            if fixed_adaptations is None:
                break
            stopped = True
        if stopped:
            continue
        best = (s_tmp.copy(), a0_dense, am_recon, fm_recon, ph_recon, a)            # :397-402 (am_fin aliases am_recon)

    s_recon, a0_fin, am_fin, fm_fin, pm_fin, a_best = best
    c_all = ti - 1
    return OracleResult(s_recon=s_recon, SRER=SRER, ti=c_all, isSpeech=speech.copy(),
                        isVoiced=analysed.copy(), a0=a0_fin[c_all].copy(), am=am_fin[c_all].copy(),
                        fm=fm_fin[c_all].copy(), ph=pm_fin[c_all].copy(),
                        window_lengths=window_lengths, n_adapt_best=a_best, per_adapt=per_adapt,
                        n_fallback=n_fallback)
