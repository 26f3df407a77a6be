"""CPU: host-side logic (functions.py:95-161,176-194 restated in eaqhm_b200/host.py) and the C ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from common import ROOT, inputs_from_golden, load_golden
from eaqhm_b200 import _lib, api, host


def test_gender_limits():
    assert host.gender_limits("male") == (70, 180)
    assert host.gender_limits("female") == (160, 300)
    assert host.gender_limits("child") == (300, 600)
    assert host.gender_limits("other") == (70, 500)
    assert host.gender_limits((90, 250)) == (90, 250)


@pytest.mark.parametrize("name", ["sa19", "synth_a_adpt0", "synth_child_vo"])
def test_mode_handling_matches_reference_flags(name):
    g, kw = load_golden(name)
    inp = inputs_from_golden(g, kw)
    # the reference mutates its Frame list in place; the fixture holds the flags after the run
    assert (np.asarray(inp["voiced5"]) == g["voiced5_final"]).all()
    assert (np.asarray(inp["speech5"]) == g["speech5_final"]).all()
    assert host.harmonic_capacity(inp["fs"], inp["f0s"])[1] == int(g["Kmax"])


@pytest.mark.parametrize("name", ["sa19", "synth_child_vo"])
def test_frame_tables_match_reference_flags(name):
    g, kw = load_golden(name)
    inp = inputs_from_golden(g, kw)
    ft = host.FrameTables(len(inp["s"]), inp["fs"], inp["f0s"], inp["voiced5"], inp["frame_step"], inp["Kmax"],
                          inp["step"], 3, inp["analysisWindow"])
    sel = g["frame_sel"]
    assert (ft.ti0[sel] == g["det_ti"]).all()
    assert (ft.isSpeech[sel] == g["det_isSpeech"]).all()
    assert (ft.analysed[sel] == g["det_isVoiced"]).all()
    a = ft.analysed
    assert (ft.N[a] >= 120).all() and (ft.K0[a] >= 1).all() and (ft.K0[a] <= inp["Kmax"]).all()
    assert (ft.N[~a] == 0).all()
    assert ft.own.sum() == len(inp["s"])
    assert ft.f0stale == ft.f0[np.flatnonzero(a)[-1]]


def test_window_half_length_rounds_half_to_even():
    # functions.py:191 uses Python round(): (5/2)*(16000/128) = 312.5 -> 312 (half-up would give 313)
    fs = 16000
    f0s = np.array([[0.0, 128.0, 1.0]] * 400)
    ft = host.FrameTables(20000, fs, f0s, np.ones(400, bool), 80, 60, pitchPeriods=5)
    assert set(int(x) for x in ft.N[ft.analysed]) == {312}
    ft = host.FrameTables(20000, fs, f0s, np.ones(400, bool), 80, 60, pitchPeriods=3)
    assert set(int(x) for x in ft.N[ft.analysed]) == {188}      # 187.5 -> 188


def test_window_past_signal_end_raises_like_reference():
    fs = 48000
    f0s = np.array([[0.0, 80.0, 1.0]] * 100)
    with pytest.raises(IndexError):
        host.FrameTables(6000, fs, f0s, np.ones(100, bool), 240, 300)   # N = 900 > analysisWindow*step


def test_voiced_only_target_keeps_voiced_regions_only():
    s = np.arange(1000, dtype=float)
    speech = np.zeros(13, bool)
    voiced = np.zeros(13, bool)
    speech[3:6] = voiced[3:6] = True
    ss = host.voiced_only_target(s, speech, voiced, 80)
    ti = np.arange(1, 1000, 80)
    lo, hi = ti[3] - 80, ti[5] + 80 + 1
    assert (ss[lo:hi] == s[lo:hi]).all() and ss[:lo].sum() == 0 and ss[hi:].sum() == 0
    speech[10:] = voiced[10:] = True        # region reaching the last frame is never copied (Q9)
    assert (host.voiced_only_target(s, speech, voiced, 80) == ss).all()


def test_pack_det_components_shapes_follow_reference_quirk_q8():
    from eaqhm_b200.engine import Result
    am = np.array([[0.0, 0.5, 0.0, 0.25, 0.0], [0, 0, 0, 0, 0.0]])
    res = Result(s_recon=np.zeros(4), SRER=[1.0], ti=np.array([0, 15]), isSpeech=np.array([True, True]),
                 isVoiced=np.array([True, False]), a0=np.array([0.1, 0.0]), am=am, fm=am * 1000, ph=am)
    det = api.pack_det_components(res)
    assert det[0].ak == [] and det[0].amplitudes.dtype == object and len(det[0].amplitudes) == 4
    assert isinstance(det[0].amplitudes[0], int) and det[0].amplitudes[1].shape == (1,)
    assert float(det[0].frange[3][0]) == 250.0 and det[1].a0 == [] and det[1].isVoiced is False
    res.isVoiced[1] = True
    with pytest.raises(IndexError):          # voiced frame without an accepted harmonic (misc.py:61-62)
        api.pack_det_components(res)


def test_pack_det_components_on_dense_and_gappy_frames():
    """Batched packing against the per-frame definition (misc.py:65-93 through functions.py:409-411): an object array dense
    up to the last accepted harmonic, one-element float arrays at the accepted harmonics, int 0 at the dropped ones."""
    from eaqhm_b200.engine import Result
    rng = np.random.default_rng(5)
    G, K = 60, 12
    am = rng.random((G, K)) + 0.1
    for i in range(G):
        am[i, rng.integers(3, K + 1):] = 0.0            # frames end at different harmonics
    am[rng.random((G, K)) < 0.1] = 0.0                  # dropped harmonics in between
    am[:, 0] = np.where(am.sum(1) == 0, 0.3, am[:, 0])  # no voiced frame without a harmonic
    voiced = rng.random(G) > 0.2
    res = Result(s_recon=np.zeros(4), SRER=[1.0], ti=np.arange(G) * 15, isSpeech=np.ones(G, bool), isVoiced=voiced,
                 a0=rng.random(G), am=am, fm=am * 1000, ph=am - 0.5)
    det = api.pack_det_components(res)
    assert len(det) == G
    for i, d in enumerate(det):
        assert d.ti == 15 * i and d.isVoiced is bool(voiced[i]) and d.ak == []
        if not voiced[i]:
            assert d.a0 == [] and d.frange == [] and d.pk == [] and not hasattr(d, "amplitudes")
            continue
        nz = np.flatnonzero(am[i])
        assert d.a0 == res.a0[i]
        for arr, src in ((d.amplitudes, res.am), (d.frange, res.fm), (d.pk, res.ph)):
            assert arr.dtype == object and arr.shape == (nz[-1] + 1,)
            for k in range(nz[-1] + 1):
                if am[i, k] != 0:
                    assert isinstance(arr[k], np.ndarray) and arr[k].shape == (1,) and arr[k][0] == src[i, k]
                else:
                    assert isinstance(arr[k], int) and arr[k] == 0


def test_c_abi_exports_every_declared_symbol():
    _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "eaqhm_b200.h")).read()
    declared = set(re.findall(r"\b(eaqhm_[A-Za-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.exported_symbols())
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().eaqhm_version() >= 100
    assert _lib.lib().eaqhm_analysis_scratch_bytes(100, 10, 40, 150) > 0
    assert _lib.lib().eaqhm_analysis_scratch_bytes(100, 0, 40, 150) < 0


def test_struct_layout_matches_header():
    # field order/sizes of the ctypes mirrors must follow include/eaqhm_b200.h
    assert ctypes.sizeof(_lib.Tracks) == 40 and ctypes.sizeof(_lib.Knots) == 32
    assert _lib.Tables.T.offset == 0 and _lib.Tables.maxlen.offset == 8 and _lib.Tables.G.offset == 16
    assert _lib.Tables.fs.offset == 32 and _lib.Tables.pos.offset == 40 and _lib.Tables.nA.offset == 40 + 17 * 8 and _lib.Tables.flist.offset == 40 + 18 * 8


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from eaqhm_b200 import engine
    g, kw = load_golden("synth_a_adpt0")
    inp = inputs_from_golden(g, kw)
    u = engine.Utterance(s=inp["s"], fs=inp["fs"], f0s=inp["f0s"], voiced5=inp["voiced5"],
                         frame_step=inp["frame_step"], f0min=inp["f0min"], Kmax=inp["Kmax"])
    with pytest.raises(_lib.EaqhmError):
        engine.Engine([u])
    amp = np.zeros(6)
    rc = _lib.lib().eaqhm_iqhmLS_complexamps(
        np.zeros(5).ctypes.data_as(_lib.c_f64p), 5, np.zeros(3).ctypes.data_as(_lib.c_f64p), 3,
        np.ones(5).ctypes.data_as(_lib.c_f64p), 16000.0, amp.ctypes.data_as(_lib.c_f64p),
        amp.ctypes.data_as(_lib.c_f64p))
    assert rc == -4    # EAQHM_E_NODEVICE: no CPU fallback


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) must print exactly one JSON line on stdout
    with the contract's keys, whatever the libraries print elsewhere."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-excerpt-seconds", "0.09"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["workload"].startswith("c2")


def test_wrapping_windows_are_reported_and_refused_by_the_streamed_engine():
    """A window that starts before sample 0 wraps to the END of the signal in the reference (functions.py:196 with a
    negative start index): 48 kHz, f0 = 80 Hz, default analysisWindow -> N = 900 > 480.  The monolithic engine
    reproduces it; a frame-range excerpt cannot, so FrameTables reports such frames for the streamed engine's guard."""
    import numpy as np
    from eaqhm_b200 import host
    fs, length = 48000, 48000
    n5 = len(np.arange(1, length, 240))
    t = np.arange(0, length - 1, 240) / fs
    f0s = np.stack([t, np.full(len(t), 80.0), np.ones(len(t))], axis=1)
    v5 = np.ones(n5, bool)
    v5[-6:] = False                  # the reference raises IndexError on a window past the END (FrameTables too)
    ft = host.FrameTables(length, fs, f0s, v5, 240, 300)
    w = ft.wrapping_frames()
    assert len(w) > 0 and (ft.ti0[w] < ft.N[w]).all() and ft.ti0[w].min() >= 480
    ft64 = host.FrameTables(length, fs, f0s, v5, 240, 300, analysisWindow=64)
    assert len(ft64.wrapping_frames()) == 0


def test_not_positive_definite_error_is_a_linalg_error():
    import numpy as np
    from eaqhm_b200 import engine
    assert issubclass(engine.NotPositiveDefiniteError, np.linalg.LinAlgError)
    engine.check_not_pd(0)
    import pytest
    with pytest.raises(np.linalg.LinAlgError):
        engine.check_not_pd(3)


def test_compact_parameter_file_round_trip(tmp_path):
    """params.save_result / load_result on a hand-made result: no GPU involved."""
    import numpy as np
    from eaqhm_b200 import engine, params
    rng = np.random.default_rng(0)
    length, step, K = 1000, 15, 7
    n = len(np.arange(1, length, step))
    voiced = rng.random(n) < 0.7
    am = rng.random((n, K)) * (rng.random((n, K)) < 0.6) * voiced[:, None]
    fm = (am != 0) * rng.uniform(100, 7000, (n, K))
    ph = (am != 0) * rng.uniform(-3, 3, (n, K))
    res = engine.Result(s_recon=rng.standard_normal(length), SRER=[10.0, 12.5], ti=np.arange(0, length - 1, step),
                        isSpeech=voiced | (rng.random(n) < 0.1), isVoiced=voiced, a0=rng.standard_normal(n) * voiced,
                        am=am, fm=fm, ph=ph, n_adapt_best=1, fs=16000, step=step)
    p = str(tmp_path / "r.npz")
    params.save_result(p, res, include_waveform=True)
    back = params.load_result(p)
    for name in ("s_recon", "ti", "isSpeech", "isVoiced", "a0", "am", "fm", "ph"):
        assert (getattr(back, name) == getattr(res, name)).all(), name
    assert back.SRER == res.SRER and back.n_adapt_best == 1 and back.fs == 16000 and back.step == step


def test_roofline_traffic_comes_from_a_committed_ncu_capture():
    """bench.py's roofline.traffic is parsed from profiles/ncu_traffic.json, which tools/ncu_traffic.py extracts from a
    committed raw ncu CSV (no hard-coded constant): the file it names exists, and a workload without capture gives null."""
    import json
    import bench
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    t = json.load(open(os.path.join(root, "profiles", "ncu_traffic.json")))
    assert os.path.exists(os.path.join(root, t["source"])) and t["workload"] == "c2"
    v, src = bench.measured_traffic("gram", "c2", 20.0)
    assert v == t["kernels"]["gram"]["dram_bytes_per_launch"] and v > 1e9 and t["source"] in src
    # the capture's two launches per adaptation against a run that processes the adaptation's frames in one chunk
    v1, src1 = bench.measured_traffic("gram", "c2", 20.0, chunks_now=1)
    assert v1 == v * t["chunks_per_adaptation"] and "rescaled" in src1
    assert bench.measured_traffic("gram", "c2", 20.0, chunks_now=t["chunks_per_adaptation"])[0] == v
    assert bench.measured_traffic("gram", "c3", 96.0)[0] is None
    assert bench.measured_traffic("gram", "c2", 5.0)[0] is None
