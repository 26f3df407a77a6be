"""Import shim for the UNMODIFIED reference at /root/reference (container only).

The reference imports pylab / matplotlib / tkinter and two names removed from
numpy 2.  This shim registers stub modules so that
``import functions`` works; the hot-path functions themselves run unmodified.
It is used ONLY by the golden-vector generator (make_golden.py) and by the
optional container-only cross-check tests; nothing on the GPU box imports it.
"""
import sys
import types

import numpy as np

REFERENCE_DIR = "/root/reference"


def _specgram(x, NFFT, Fs, window, noverlap, **_kw):
    """matplotlib.pyplot.specgram(mode='psd') restated: one-sided PSD, no detrend."""
    x = np.asarray(x, float)
    step = NFFT - noverlap
    nseg = (len(x) - noverlap) // step
    idx = np.arange(NFFT)[:, None] + step * np.arange(nseg)[None, :]
    seg = x[idx] * np.asarray(window, float)[:, None]
    X = np.fft.rfft(seg, n=NFFT, axis=0)
    P = (X.conj() * X).real
    if NFFT % 2 == 0:
        P[1:-1] *= 2.0
    else:
        P[1:] *= 2.0
    P /= Fs
    P /= (np.abs(window) ** 2).sum()
    freqs = np.fft.rfftfreq(NFFT, 1.0 / Fs)
    t = np.arange(NFFT / 2, len(x) - NFFT / 2 + 1, step) / Fs
    return P, freqs, t, None


def load_reference():
    """Returns the reference's ``functions`` module (imported from /root/reference)."""
    if "functions" in sys.modules and getattr(sys.modules["functions"], "__file__", "").startswith(REFERENCE_DIR):
        return sys.modules["functions"]
    if not hasattr(np, "round_"):
        np.round_ = np.round
    if not hasattr(np, "NAN"):
        np.NAN = np.nan
    pylab = types.ModuleType("pylab")
    pylab.norm = np.linalg.norm
    pylab.nan = np.nan
    pylab.size = np.size
    pylab.fix = np.fix
    pylab.polyval = np.polyval
    pylab.polyfit = np.polyfit
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.specgram = _specgram
    for name in ("plot", "show", "xlabel", "ylabel", "title", "colorbar"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules.setdefault("pylab", pylab)
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    if "numpy.matlib" not in sys.modules:
        try:
            import numpy.matlib  # noqa: F401
        except Exception:
            ml = types.ModuleType("numpy.matlib")
            ml.repmat = lambda a, m, n: np.tile(np.atleast_2d(a), (m, n))
            sys.modules["numpy.matlib"] = ml
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import functions  # noqa: E402  (the reference module)
    return functions
