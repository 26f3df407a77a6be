"""CPU: host preprocessing restated in eaqhm_b200/preprocess.py (SWIPE' f0, 5 ms resampling, speech/voiced flags;
SWIPE.py:14-184, functions.py:577-680, misc.py:167-206) against what the reference produced for the golden fixtures."""
import numpy as np

from common import load_golden
from eaqhm_b200 import host, preprocess


def test_swipe_and_resampling_reproduce_reference_f0_grid_on_sa19():
    g, _ = load_golden("sa19")
    s = g["pcm"] / host.NORMALIZE
    f0 = preprocess.swipep(s, int(g["fs"]), [int(g["f0min"]), int(g["f0max"])])
    assert f0.shape == (3968, 3) and not np.isnan(f0).any()
    grid = np.arange(0, len(s) - 1, 80) / int(g["fs"])
    f0s = preprocess.getLinear(f0, grid)
    assert f0s.shape == g["f0s"].shape
    assert np.abs(f0s - g["f0s"]).max() <= 1e-9
    assert host.harmonic_capacity(int(g["fs"]), f0s)[1] == int(g["Kmax"])


def test_speech_voiced_flags_reproduce_reference_including_half_valued_medians():
    for name, gender in (("sa19", "female"), ("synth_child_vo", (250, 500))):
        g, _ = load_golden(name)
        sp, vo, step = preprocess.voicedUnvoicedFrames(g["pcm"] / host.NORMALIZE, int(g["fs"]), gender)
        assert step == int(g["frame_step"])
        assert (sp.astype(bool) == g["speech5"]).all() and (vo.astype(bool) == g["voiced5"]).all()
    assert set(np.unique(sp)) <= {0.0, 0.5, 1.0}      # 4-tap "median of 5" quirk (misc.py:202-206)


def test_getlinear_edges_and_interior():
    v = np.array([[0.0, 10.0, 1.0], [1.0, 20.0, 2.0], [2.0, 40.0, 4.0]])
    out = preprocess.getLinear(v, np.array([-1.0, 0.0, 0.25, 1.0, 1.5, 2.0, 3.0]))
    assert np.allclose(out[:, 1], [10, 10, 12.5, 20, 30, 40, 40]) and np.allclose(out[:, 2], [1, 1, 1.25, 2, 3, 4, 4])
    assert np.allclose(out[:, 0], [-1.0, 0.0, 0.25, 1.0, 1.5, 2.0, 3.0])


def test_ellipfilter_is_zero_phase_highpass():
    fs = 16000
    t = np.arange(fs) / fs
    x = np.sin(2 * np.pi * 5 * t) + 0.1 * np.sin(2 * np.pi * 1000 * t)
    y = preprocess.ellipFilter(x, fs, 30)
    mid = slice(2000, -2000)
    assert np.abs(y[mid] - 0.1 * np.sin(2 * np.pi * 1000 * t)[mid]).max() < 2e-2
