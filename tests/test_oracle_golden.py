"""CPU: pins the oracle (oracle/eaqhm_oracle.py) against the reference.

Golden fixtures were produced by the unmodified reference in the build container
(tests/golden/make_golden.py); the SA19 SRER values are also the ones the reference
publishes (img/SA19out.JPG via README.md:46-48).
"""
import numpy as np
import pytest

import eaqhm_oracle as O
from common import GOLDEN, compare_to_golden, inputs_from_golden, load_golden

PUBLISHED_SA19_SRER = [17.86520945273994, 24.431728752204954, 24.67698055430504, 25.291120491477024,
                       25.497403658214047, 25.446628776435006]


def run_oracle(inp, **over):
    kw = dict(step=inp["step"], maxAdpt=inp["maxAdpt"], analysisWindow=inp["analysisWindow"], det_part=inp["det_part"])
    kw.update(over)
    return O.eaqhm_oracle(inp["s"], inp["fs"], inp["f0s"], inp["voiced5"], inp["frame_step"], inp["f0min"],
                          inp["Kmax"], **kw)


def test_helper_seams_match_reference():
    h = np.load(GOLDEN + "/helpers.npz")
    a, b = O.iqhmLS_complexamps(h["iq_s"], h["iq_f0range"], h["iq_win"], 16000)
    assert np.abs(a - h["iq_amp"]).max() <= 1e-13 * np.abs(h["iq_amp"]).max()
    assert np.abs(b - h["iq_slope"]).max() <= 1e-10 * np.abs(h["iq_slope"]).max()
    a, b = O.eaqhmLS_complexamps(h["ea_s"], h["ea_am"], h["ea_fm"], h["ea_win"], 16000)
    assert np.abs(a - h["ea_amp"]).max() <= 1e-13 * np.abs(h["ea_amp"]).max()
    assert np.abs(b - h["ea_slope"]).max() <= 1e-12 * np.abs(h["ea_slope"]).max()
    p = O.phase_integr_interpolation(h["pi_fm"], h["pi_ph"], h["pi_idx"])
    assert np.abs(p - h["pi_out"]).max() <= 1e-13


def test_notaknot_spline_matches_scipy():
    from scipy.interpolate import interp1d
    rng = np.random.default_rng(0)
    for n in (4, 5, 6, 9, 200):
        x = np.arange(n) * 15 + 480
        y = rng.normal(size=n) * 1000 + 3000
        xq = np.arange(x[0], x[-1] + 1)
        assert np.abs(interp1d(x, y, kind=3)(xq) - O.notaknot_cubic(x, y, xq)).max() <= 1e-12 * np.abs(y).max()
    x = np.array([0, 15, 3000, 3015])          # short-run quirk Q4
    y = np.array([0, 0, 2500., 2510.])
    xq = np.arange(3000, 3016)
    assert np.abs(interp1d(x, y, kind=3)(xq) - O.notaknot_cubic(x, y, xq)).max() <= 1e-11


@pytest.mark.parametrize("name", ["synth_a_adpt0", "synth_a_adpt3", "synth_child_vo", "synth_silence"])
def test_oracle_matches_reference_golden(name):
    g, kw = load_golden(name)
    res = run_oracle(inputs_from_golden(g, kw))
    c = compare_to_golden(res, g)
    assert c["n_adapt"][0] == c["n_adapt"][1]
    assert c["flags_ok"] and c["support_equal"]
    assert c["dSRER"] <= 1e-9
    assert c["ds"] <= 1e-10
    assert c["am_rel_framemax"] <= 1e-10
    assert c["fm_rel_strong"] <= 1e-10


def test_oracle_sa19_first_adaptations_match_published_srer():
    """SA19.WAV / female / defaults: adaptations 0 and 1 against the published screenshot values."""
    g, kw = load_golden("sa19")
    res = run_oracle(inputs_from_golden(g, kw), maxAdpt=1)
    assert abs(res.SRER[0] - PUBLISHED_SA19_SRER[0]) <= 1e-12
    assert abs(res.SRER[1] - PUBLISHED_SA19_SRER[1]) <= 1e-9
    assert np.abs(np.array(res.SRER) - g["SRER"][:2]).max() <= 1e-9


def test_golden_sa19_is_the_published_run():
    g, _ = load_golden("sa19")
    assert len(g["SRER"]) == 6
    assert np.abs(g["SRER"] - np.array(PUBLISHED_SA19_SRER)).max() <= 1e-10


@pytest.mark.slow
@pytest.mark.parametrize("name", ["sa19", "synth_low", "synth_48k"])
def test_oracle_matches_reference_golden_slow(name):
    g, kw = load_golden(name)
    res = run_oracle(inputs_from_golden(g, kw))
    c = compare_to_golden(res, g)
    assert c["n_adapt"][0] == c["n_adapt"][1] and c["flags_ok"] and c["support_equal"]
    assert c["dSRER"] <= 1e-9 and c["ds"] <= 1e-10 and c["am_rel_framemax"] <= 5e-9
