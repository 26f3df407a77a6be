"""Seeded synthetic harmonic AM-FM speech-like signals.

Pure numpy, host side.  Used by bench.py, the tests and the golden-vector
generator so that every side analyses exactly the same samples and f0 track.
"""
import numpy as np


def f0_grid_times(length, fs):
    """Times of the 5 ms f0 grid the reference resamples SWIPE's track to (functions.py:113)."""
    return np.arange(0, length - 1, round(fs * 5 / 1000)) / fs


def harmonic_signal(f0_of_t, duration, fs=16000, seed=0, noise_db=-50.0, peak=0.5,
                    voiced_mask_of_t=None, unvoiced_noise_db=-30.0, quantise=True):
    """Returns (s, f0s): s (length,) float64 in [-1,1) (int16-quantised like a wav read
    through functions.py:86-87 when ``quantise``) and f0s (n5,3) rows [t, f0, 1.0].

    s[n] = sum_k A_k(n) cos(k*phi(n) + theta_k), phi = 2*pi*cumsum(f0)/fs,
    K(n) = floor((fs/2-200)/f0(n)), A_k = (1/k) * (1 + d_k sin(2*pi*r_k*t + p_k)),
    r_k in [2,6] Hz, d_k <= 0.3, plus white noise.
    """
    rng = np.random.default_rng(seed)
    length = int(round(duration * fs))
    t = np.arange(length) / fs
    f0 = np.asarray(f0_of_t(t), float)
    phi = 2 * np.pi * np.cumsum(f0) / fs
    Kmax = int(np.floor((fs / 2 - 200) / f0.min()))
    s = np.zeros(length)
    for k in range(1, Kmax + 1):
        theta = rng.uniform(-np.pi, np.pi)
        r = rng.uniform(2.0, 6.0)
        d = rng.uniform(0.0, 0.3)
        p = rng.uniform(-np.pi, np.pi)
        amp = (1.0 / k) * (1.0 + d * np.sin(2 * np.pi * r * t + p))
        amp = np.where(k * f0 <= fs / 2 - 200, amp, 0.0)
        s += amp * np.cos(k * phi + theta)
    s *= peak / np.abs(s).max()
    if voiced_mask_of_t is not None:
        vm = np.asarray(voiced_mask_of_t(t), bool)
        uv = rng.standard_normal(length) * peak * 10 ** (unvoiced_noise_db / 20)
        s = np.where(vm, s, uv)
    s = s + rng.standard_normal(length) * peak * 10 ** (noise_db / 20)
    if quantise:
        s = np.clip(np.round(s * 32768), -32768, 32767) / 32768
    tg = f0_grid_times(length, fs)
    f0g = np.asarray(f0_of_t(tg), float)
    f0s = np.stack([tg, f0g, np.ones_like(tg)], axis=1)
    return s, f0s


def config_c2(duration=3600.0, seed=2):
    """BASELINE config 2: 16 kHz, f0(t) = 170 + 80 sin(2*pi*0.25*t + phi0), gender=(90,250)."""
    phi0 = np.random.default_rng(seed).uniform(-np.pi, np.pi)
    return dict(fs=16000, duration=duration, seed=seed, gender=(90, 250),
                f0_of_t=lambda t: 170.0 + 80.0 * np.sin(2 * np.pi * 0.25 * t + phi0))


def config_c3(index, duration=3.0):
    """BASELINE config 3: utterance ``index`` of the 10,000 x 3 s corpus, gender=(90,250)."""
    rng = np.random.default_rng(10_000_019 + index)
    c, d, r, p = rng.uniform(120, 220), rng.uniform(10, 30), rng.uniform(1, 4), rng.uniform(-np.pi, np.pi)
    return dict(fs=16000, duration=duration, seed=index, gender=(90, 250),
                f0_of_t=lambda t: np.clip(c + d * np.sin(2 * np.pi * r * t + p), 90.0, 250.0))


def config_c4(duration=1.0, seed=4):
    """BASELINE config 4: 48 kHz, f0 = 80 + 3 sin(2*pi*0.5*t), gender=(70,100)."""
    return dict(fs=48000, duration=duration, seed=seed, gender=(70, 100),
                f0_of_t=lambda t: 80.0 + 3.0 * np.sin(2 * np.pi * 0.5 * t))


def config_c5(index, duration=3.0):
    """BASELINE config 5: child-range f0, alternating 300 ms voiced / 200 ms noise, gender=(250,500)."""
    rng = np.random.default_rng(50_000_017 + index)
    c, d, r, p = rng.uniform(300, 450), rng.uniform(10, 40), rng.uniform(1, 4), rng.uniform(-np.pi, np.pi)
    return dict(fs=16000, duration=duration, seed=index, gender=(250, 500),
                f0_of_t=lambda t: np.clip(c + d * np.sin(2 * np.pi * r * t + p), 250.0, 500.0),
                voiced_mask_of_t=lambda t: np.mod(t, 0.5) < 0.3)


def make(cfg):
    """Build (s, f0s) from one of the config dicts above."""
    return harmonic_signal(cfg["f0_of_t"], cfg["duration"], fs=cfg["fs"], seed=cfg["seed"],
                           voiced_mask_of_t=cfg.get("voiced_mask_of_t"))
