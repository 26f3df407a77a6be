"""GPU (-m gpu): parity of the CUDA path, called through the C ABI, against the golden fixtures made by
the reference and against the oracle on seeded inputs.

Tolerances (north_star): waveform 1e-7 absolute, SRER 0.01 dB, amplitudes / frequencies 1e-9 relative, all of
them asserted here through tests/common.tolerance():
  am_rel_framemax   amplitudes against the frame's largest amplitude                      1e-9
  am_rel_strong     amplitudes per harmonic (harmonics within 60 dB of the frame maximum) max(1e-9, 10 x floor)
  fm_rel_strong     frequencies per harmonic (same harmonics)                             max(1e-9, 10 x floor)
  ph_abs_strong     phases (rad, same harmonics)                                          max(1e-9, 10 x floor)
  a0_abs            DC component                                                          1e-9
`floor` is the reference's own run-to-run noise on the same fixture (the unmodified reference re-run with one
instead of eight OpenBLAS threads; tests/golden/make_floor.py -> tests/golden/floors.json): two runs of the
reference differ per harmonic by 0.5-1.7e-10 after one to four adaptations on the 0.4 s fixtures already, because
its explicit inverse (functions.py:465,530) is noisy at cond(R) up to 1e6 and the adaptation loop feeds that back.
Harmonics more than 60 dB below the frame maximum sit below that noise and are checked through am_rel_framemax only.
"""
import ctypes as C

import numpy as np
import pytest

import eaqhm_oracle as O
from common import GOLDEN, KEYS, compare_to_golden, inputs_from_golden, load_golden, tolerance

pytestmark = pytest.mark.gpu

TOL_S = 1e-7
TOL_SRER = 0.01
TOL_REL = 1e-9


def P(a):
    from eaqhm_b200 import _lib
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(_lib.c_f64p)


def utterance(inp):
    from eaqhm_b200 import engine
    return engine.Utterance(s=inp["s"], fs=inp["fs"], f0s=inp["f0s"], voiced5=inp["voiced5"],
                            frame_step=inp["frame_step"], f0min=inp["f0min"], Kmax=inp["Kmax"], det_part=inp["det_part"])


def run_cuda(inp, **kw):
    from eaqhm_b200 import engine
    eng = engine.Engine([utterance(inp)], step=inp["step"], maxAdpt=inp["maxAdpt"],
                        analysisWindow=inp["analysisWindow"], **kw)
    return eng.run()[0]


def run_oracle(inp):
    return O.eaqhm_oracle(inp["s"], inp["fs"], inp["f0s"], inp["voiced5"], inp["frame_step"], inp["f0min"],
                          inp["Kmax"], step=inp["step"], maxAdpt=inp["maxAdpt"], analysisWindow=inp["analysisWindow"],
                          det_part=inp["det_part"])


def assert_close_results(res, ref):
    assert len(res.SRER) == len(ref.SRER)
    assert np.abs(np.array(res.SRER) - np.array(ref.SRER)).max() <= TOL_SRER
    assert np.abs(res.s_recon - ref.s_recon).max() <= TOL_S
    v = ref.isVoiced
    assert (res.isVoiced == v).all() and (res.isSpeech == ref.isSpeech).all() and (res.ti == ref.ti).all()
    assert ((res.am != 0) == (ref.am != 0)).all()
    fmax = np.abs(ref.am[v]).max(axis=1, keepdims=True)
    assert (np.abs(res.am[v] - ref.am[v]) / fmax).max() <= TOL_REL
    strong = (ref.am[v] > 1e-3 * fmax) & (ref.fm[v] != 0)
    assert ((res.fm[v] == 0) == (ref.fm[v] == 0)).all()
    assert (np.abs(res.fm[v] - ref.fm[v])[strong] / np.abs(ref.fm[v][strong])).max() <= TOL_REL
    assert np.abs(res.a0 - ref.a0).max() <= TOL_REL
    # per-harmonic relative amplitude and absolute phase of the harmonics within 60 dB of the frame maximum
    sa = ref.am[v] > 1e-3 * fmax
    assert (np.abs(res.am[v] - ref.am[v])[sa] / ref.am[v][sa]).max() <= tolerance(None, "am_rel_strong")
    assert np.abs(res.ph[v] - ref.ph[v])[sa].max() <= tolerance(None, "ph_abs_strong")


# ---------------------------------------------------------------------------------------------
# helper seams (functions.py:420,472,537) through the host-buffer C entry points
# ---------------------------------------------------------------------------------------------
def call_iqhm(lib, s, f0range, win, fs):
    M = len(f0range)
    amp, slope = np.zeros(2 * M), np.zeros(2 * M)
    rc = lib.eaqhm_iqhmLS_complexamps(P(s), len(s), P(f0range), M, P(win), float(fs), P(amp), P(slope))
    return rc, amp.view(complex), slope.view(complex)


def call_eaqhm(lib, s, am, fm, win, fs):
    M = am.shape[1]
    amp, slope = np.zeros(2 * M), np.zeros(2 * M)
    # P() keeps the temporaries alive only for the call; pass contiguous copies explicitly
    am_c, fm_c = np.ascontiguousarray(am, float), np.ascontiguousarray(fm, float)
    rc = lib.eaqhm_eaqhmLS_complexamps(P(s), len(s), am_c.ctypes.data_as(C.POINTER(C.c_double)),
                                       fm_c.ctypes.data_as(C.POINTER(C.c_double)), M, P(win), float(fs),
                                       amp.ctypes.data_as(C.POINTER(C.c_double)),
                                       slope.ctypes.data_as(C.POINTER(C.c_double)))
    return rc, amp.view(complex), slope.view(complex)


def test_seams_match_reference_golden(lib):
    h = np.load(GOLDEN + "/helpers.npz")
    rc, a, b = call_iqhm(lib, h["iq_s"], h["iq_f0range"], h["iq_win"], 16000)
    assert rc == 0
    assert np.abs(a - h["iq_amp"]).max() <= TOL_REL * np.abs(h["iq_amp"]).max()
    assert np.abs(b - h["iq_slope"]).max() <= TOL_REL * np.abs(h["iq_slope"]).max()
    rc, a, b = call_eaqhm(lib, h["ea_s"], h["ea_am"], h["ea_fm"], h["ea_win"], 16000)
    assert rc == 0
    assert np.abs(a - h["ea_amp"]).max() <= TOL_REL * np.abs(h["ea_amp"]).max()
    assert np.abs(b - h["ea_slope"]).max() <= TOL_REL * np.abs(h["ea_slope"]).max()
    from eaqhm_b200 import _lib
    idx = np.ascontiguousarray(h["pi_idx"], np.int64)
    out = np.zeros(idx[-1] - idx[0] + 1)
    rc = lib.eaqhm_phase_integr_interpolation(P(h["pi_fm"]), P(h["pi_ph"]), len(h["pi_fm"]),
                                              idx.ctypes.data_as(_lib.c_i64p), len(idx), P(out))
    assert rc == 0 and np.abs(out - h["pi_out"]).max() <= 1e-12


@pytest.mark.parametrize("K,N,seed", [(1, 120, 0), (7, 120, 1), (31, 133, 2), (60, 150, 3), (86, 267, 4),
                                      (130, 400, 5)])       # the last one: P = 522 unknowns -> panel Cholesky
def test_iqhm_seam_recovers_known_stationary_harmonics(lib, K, N, seed):
    """Known-answer test: a stationary harmonic signal is recovered at adaptation 0."""
    rng = np.random.default_rng(seed)
    fs, f0 = 16000, 7800.0 / (K + 0.5)
    n = np.arange(-N, N + 1)
    ck = (rng.uniform(0.1, 1, K) / np.arange(1, K + 1)) * np.exp(1j * rng.uniform(-np.pi, np.pi, K))
    s = 0.05 + sum(2 * (ck[k] * np.exp(2j * np.pi * (k + 1) * f0 * n / fs)).real for k in range(K))
    f0range = np.arange(-K, K + 1) * f0
    win = np.blackman(2 * N + 1)
    rc, a, b = call_iqhm(lib, s, f0range, win, fs)
    assert rc == 0
    assert abs(a[K] - 0.05) <= 1e-10
    assert np.abs(a[K + 1:] - ck).max() <= 1e-10 and np.abs(a[:K][::-1] - ck.conj()).max() <= 1e-10
    ao, bo = O.iqhmLS_complexamps(s, f0range, win, fs)
    assert np.abs(a - ao).max() <= TOL_REL * np.abs(ao).max()


@pytest.mark.parametrize("Kp,N,seed", [(1, 120, 0), (12, 127, 1), (40, 150, 2)])
def test_eaqhm_seam_matches_oracle_on_chirps(lib, Kp, N, seed):
    rng = np.random.default_rng(seed)
    fs, f0 = 16000, 7000.0 / (Kp + 1)
    n = np.arange(-N, N + 1)
    fm = np.stack([(k + 1) * (f0 + 0.03 * n + 1e-4 * n ** 2) for k in range(Kp)], axis=1)
    am = np.stack([(1.0 / (k + 1)) * (1 + 0.002 * n * np.cos(k + seed)) for k in range(Kp)], axis=1)
    ph = 2 * np.pi * np.cumsum(fm, axis=0) / fs
    s = (2 * am * np.cos(ph - ph[N] + rng.uniform(-3, 3, Kp)[None, :])).sum(axis=1) + 1e-3 * rng.standard_normal(len(n))
    z = np.zeros((len(n), 1))
    fm_all = np.concatenate((-fm[::-1], z, fm), axis=1)         # functions.py:284-285 (time-reversed mirror, Q2)
    am_all = np.concatenate((am[::-1], z, am), axis=1)
    win = np.hamming(len(n))
    rc, a, b = call_eaqhm(lib, s, am_all, fm_all, win, fs)
    assert rc == 0
    ao, bo = O.eaqhmLS_complexamps(s, am_all, fm_all, win, fs)
    assert np.abs(a - ao).max() <= TOL_REL * np.abs(ao).max()
    assert np.abs(b - bo).max() <= 1e-8 * np.abs(bo).max()


def test_seam_reports_singular_system(lib):
    """An all-zero window makes R exactly singular: the reference's inv() raises LinAlgError
    (functions.py:530); the C ABI returns EAQHM_E_NOTPD."""
    N = 120
    n = np.arange(-N, N + 1)
    fm = np.stack([-200.0 + 0 * n, 0.0 * n, 200.0 + 0 * n], axis=1)
    am = np.ones_like(fm)
    rc, a, b = call_eaqhm(lib, np.cos(0.1 * n), am, fm, np.zeros(len(n)), 16000)
    assert rc == -3
    with pytest.raises(np.linalg.LinAlgError):
        O.eaqhmLS_complexamps(np.cos(0.1 * n), am, fm, np.zeros(len(n)), 16000)


def test_phase_seam_random_segments(lib):
    from eaqhm_b200 import _lib
    rng = np.random.default_rng(3)
    length = 2000
    idx = np.concatenate(([30], 30 + np.cumsum(rng.integers(3, 40, 60)))).astype(np.int64)
    fm = 2 * np.pi / 16000 * (3000 + 2500 * np.sin(np.arange(length) / 17.0))
    ph = np.zeros(length)
    ph[idx] = rng.uniform(-np.pi, np.pi, len(idx))
    out = np.zeros(idx[-1] - idx[0] + 1)
    rc = lib.eaqhm_phase_integr_interpolation(P(fm), P(ph), length, idx.ctypes.data_as(_lib.c_i64p), len(idx), P(out))
    assert rc == 0
    assert np.abs(out - O.phase_integr_interpolation(fm, ph, idx)).max() <= 1e-11


# ---------------------------------------------------------------------------------------------
# whole path against the reference's golden outputs
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["synth_a_adpt0", "synth_a_adpt1", "synth_a_adpt2", "synth_a_adpt3", "synth_low",
                                  "synth_child_vo", "synth_48k", "synth_silence", "c2_2s", "c4_015", "sa19"])
def test_cuda_path_matches_reference_golden(name):
    g, kw = load_golden(name)
    res = run_cuda(inputs_from_golden(g, kw))
    c = compare_to_golden(res, g)
    assert c["n_adapt"][0] == c["n_adapt"][1], c
    assert c["flags_ok"] and c["support_equal"] and c["fm_zero_equal"], c
    for key in KEYS:                                   # SRER, samples, amplitudes (two ways), frequencies, phases, DC
        assert c[key] <= tolerance(name, key), (key, c)
    assert res.stats["not_pd_frames"] == 0


def test_streamed_c2_shards_match_the_reference():
    """an excerpt of the c2 signal across shard boundaries against the REFERENCE (not against the
    monolithic CUDA run): 2 s streamed as four and as two frame-range shards."""
    g, kw = load_golden("c2_2s")
    inp = inputs_from_golden(g, kw)
    for shard_seconds, nsh in ((0.5, 4), (1.0, 2)):
        res, eng = run_streamed(inp, shard_seconds)
        assert len(eng.shards) == nsh
        c = compare_to_golden(res, g)
        assert c["n_adapt"][0] == c["n_adapt"][1] and c["flags_ok"] and c["support_equal"] and c["fm_zero_equal"], c
        for key in KEYS:
            assert c[key] <= tolerance("c2_2s", key), (shard_seconds, key, c)


def test_not_positive_definite_frames_raise():
    """A window shorter than the unknown count (pitchPeriods=1 at f0 ~ 100 Hz: L = 241 samples, P = 2(2K+1) = 314) makes
    E_w^H E_w singular.  The reference's inv() raises LinAlgError or returns noise (functions.py:465,530); this path raises
    a LinAlgError subclass instead of handing NaN or noise on."""
    from eaqhm_b200 import engine
    g, kw = load_golden("synth_low")
    inp = inputs_from_golden(g, kw)
    with pytest.raises(np.linalg.LinAlgError):
        engine.Engine([utterance(inp)], maxAdpt=1, pitchPeriods=1).run()


def test_sa19_srer_matches_published_values():
    """The reference's only published vector (img/SA19out.JPG, README.md:46-48)."""
    published = [17.86520945273994, 24.431728752204954, 24.67698055430504, 25.291120491477024, 25.497403658214047,
                 25.446628776435006]
    g, kw = load_golden("sa19")
    res = run_cuda(inputs_from_golden(g, kw))
    assert len(res.SRER) == 6 and res.n_adapt_best == 4
    assert np.abs(np.array(res.SRER) - np.array(published)).max() <= 1e-8


# ---------------------------------------------------------------------------------------------
# CUDA path against the oracle on fresh seeded inputs
# ---------------------------------------------------------------------------------------------
def synth_inputs(cfg, maxAdpt, voiced_only=False):
    from eaqhm_b200 import host, synthetic
    s, f0s = synthetic.make(cfg)
    fs = cfg["fs"]
    fstep = int(round(0.005 * fs))
    n5 = len(np.arange(1, len(s), fstep))
    Kmax = host.harmonic_capacity(fs, f0s)[1]
    if voiced_only:
        t5 = (np.arange(1, len(s), fstep) - 1) / fs
        voiced5 = np.asarray(cfg["voiced_mask_of_t"](t5), bool)
        det = host.voiced_only_target(s, voiced5, voiced5, fstep)
    else:
        voiced5, det = np.ones(n5, bool), None
    return dict(s=s, fs=fs, f0s=f0s, voiced5=voiced5, frame_step=fstep, f0min=float(cfg["gender"][0]), Kmax=Kmax,
                det_part=det, step=15, maxAdpt=maxAdpt, analysisWindow=32)


@pytest.mark.parametrize("case", ["c2", "c3", "c5"])
def test_cuda_path_matches_oracle_on_seeded_inputs(case):
    from eaqhm_b200 import synthetic
    if case == "c2":
        cfg, vo = synthetic.config_c2(duration=0.35), False
    elif case == "c3":
        cfg, vo = synthetic.config_c3(123, duration=0.4), False
    else:
        cfg, vo = synthetic.config_c5(5, duration=0.9), True
    inp = synth_inputs(cfg, maxAdpt=2, voiced_only=vo)
    assert_close_results(run_cuda(inp), run_oracle(inp))


def test_batch_equals_single_runs_bitwise():
    """Utterance batching (config c3) must not change any utterance's result."""
    from eaqhm_b200 import engine, synthetic
    inps = [synth_inputs(synthetic.config_c3(i, duration=d), maxAdpt=3) for i, d in ((0, 0.3), (1, 0.45), (2, 0.25))]
    singles = [run_cuda(i) for i in inps]
    batch = engine.Engine([utterance(i) for i in inps], maxAdpt=3).run()
    for a, b in zip(singles, batch):
        assert a.SRER == b.SRER
        assert (a.s_recon == b.s_recon).all() and (a.am == b.am[:, :a.am.shape[1]]).all()
        assert (a.fm == b.fm[:, :a.fm.shape[1]]).all() and (a.ph == b.ph[:, :a.ph.shape[1]]).all()


def test_corpus_runner_equals_single_runs_bitwise():
    """Config c3/c5 mechanics: a corpus dealt round-robin over ranks and analysed in memory-sized batches with host
    preprocessing running ahead in threads gives every utterance the result of its own single run."""
    from eaqhm_b200 import corpus, synthetic
    inps = [synth_inputs(synthetic.config_c3(20 + i, duration=0.25 + 0.05 * (i % 3)), maxAdpt=2) for i in range(5)]
    singles = [run_cuda(i) for i in inps]
    budget = corpus.utterance_bytes(utterance(inps[0])) * 2 + 1000         # forces several batches
    seen = {}
    for world in (1, 2):
        for rank in range(world):
            for i, r in corpus.analyze_corpus(len(inps), lambda i: utterance(inps[i]), rank=rank, world=world,
                                              batch_bytes=budget, maxAdpt=2, workers=2, lookahead=3):
                seen[(world, i)] = r
    for world in (1, 2):
        for i, a in enumerate(singles):
            b = seen[(world, i)]
            assert a.SRER == b.SRER and (a.s_recon == b.s_recon).all() and (a.am == b.am).all() and (a.fm == b.fm).all()


def test_chunked_analysis_equals_unchunked_bitwise():
    from eaqhm_b200 import synthetic
    inp = synth_inputs(synthetic.config_c3(9, duration=0.3), maxAdpt=2)
    a, b = run_cuda(inp), run_cuda(inp, chunk=37)
    assert a.SRER == b.SRER and (a.s_recon == b.s_recon).all() and (a.am == b.am).all()


def test_chunk_pipeline_equals_serial_chunks_bitwise():
    """EAQHM_PIPELINE=2 (experiment knob: the Cholesky of one chunk of frames on a second internal stream beside the
    basis + Gram of the next chunk, both kernels persistent on work queues, stratified chunk order) against the shipped
    serial chunk order: the same kernels run the same per-frame arithmetic, so every output is bit-identical.  The
    library reads the knob once per process, hence two child processes."""
    import json, os, subprocess, sys
    probe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pipeline_probe.py")
    out = {}
    for sets in ("1", "2"):
        env = dict(os.environ, EAQHM_PIPELINE=sets)
        r = subprocess.run([sys.executable, probe, "260"], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[sets] = json.loads(r.stdout.strip().splitlines()[-1])
    a, b = out["1"], out["2"]
    assert a["sets"] == 1 and b["sets"] == 2
    assert b["frames"] >= 4 * b["chunk"]                 # several phases: first Gram alone, mixed phases, last Cholesky alone
    for k in ("SRER", "s", "am", "fm", "ph", "a0"):
        assert a[k] == b[k], k


def test_one_warp_and_two_warp_cholesky_agree_bitwise():
    """Launches whose largest possible system has P <= 128 factor with one warp per frame (k_chol_solve<1>), the others
    with two: a frame's result must not depend on which instance its launch got (a shard or a batch of high-pitched
    frames gets the other instance than the whole signal), so the two instances perform the same operations in the same
    order - the back substitution's partial sums included.  Forced here through EAQHM_CHOL_WARPS on the same input."""
    import json, os, subprocess, sys
    probe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pipeline_probe.py")
    out = {}
    for w in ("1", "2"):
        env = dict(os.environ, EAQHM_CHOL_WARPS=w, EAQHM_PIPELINE="1")
        r = subprocess.run([sys.executable, probe], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[w] = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("SRER", "s", "am", "fm", "ph", "a0"):
        assert out["1"][k] == out["2"][k], k


def test_basis_fast_path_equals_running_sum_path():
    """The phasor-track fast path of the adaptive basis (k_basis) against the reference-shaped path (gap scan, running
    sum, sincos per sample), selected by passing no phasor tracks through the C ABI: same support, SRER and waveform far
    inside the parity tolerances (the two differ by the rounding of the running sum only)."""
    from eaqhm_b200 import synthetic
    inp = synth_inputs(synthetic.config_c2(duration=1.5), maxAdpt=4)
    fast, slow = run_cuda(inp), run_cuda(inp, phasor_tracks=False)
    assert len(fast.SRER) == len(slow.SRER) >= 3
    assert np.abs(np.array(fast.SRER) - np.array(slow.SRER)).max() <= 1e-9
    assert np.abs(fast.s_recon - slow.s_recon).max() <= 1e-10
    assert ((fast.am != 0) == (slow.am != 0)).all()
    v = fast.isVoiced
    fmax = np.abs(slow.am[v]).max(axis=1, keepdims=True)
    assert (np.abs(fast.am[v] - slow.am[v]) / fmax).max() <= 1e-10


def test_rerun_is_deterministic():
    from eaqhm_b200 import synthetic
    inp = synth_inputs(synthetic.config_c3(4, duration=0.3), maxAdpt=2)
    a, b = run_cuda(inp), run_cuda(inp)
    assert a.SRER == b.SRER and (a.s_recon == b.s_recon).all() and (a.fm == b.fm).all()


def test_properties_at_benchmark_shard_size():
    """Size-independent properties on a multi-second c2 shard the oracle cannot finish quickly: SRER rises until
    the stop, the returned waveform is the best adaptation's, every accepted frequency is positive and below Nyquist,
    amplitude support equals frequency support."""
    from eaqhm_b200 import synthetic
    inp = synth_inputs(synthetic.config_c2(duration=4.0), maxAdpt=10)
    res = run_cuda(inp)
    sr = np.array(res.SRER)
    assert (np.diff(sr[:-1]) > 0).all() or len(sr) <= 2
    assert res.n_adapt_best == (len(sr) - 2 if sr[-1] <= sr[-2] else len(sr) - 1)
    det = inp["s"]
    srer_from_signal = 20 * np.log10(np.std(det) / np.std(det - res.s_recon))
    assert abs(srer_from_signal - sr[res.n_adapt_best]) <= 1e-9
    v = res.isVoiced
    assert ((res.am[v] != 0) == (res.fm[v] != 0)).all()
    assert (res.fm[v][res.am[v] != 0] > 0).all() and (res.fm[v] < 8000).all()
    assert sr[res.n_adapt_best] > 15.0


def test_fallback_path_matches_reference_and_oracle():
    """Quirks Q6/Q10: frames inside digital silence accept nothing at adaptation 0 and analyse the 140 Hz / 10e-4
    fallback component afterwards (functions.py:204-242); the in-place amplitude write shows up in the returned
    parameters of the best adaptation (the am_fin alias)."""
    g, kw = load_golden("synth_silence")
    inp = inputs_from_golden(g, kw)
    res, ref = run_cuda(inp), run_oracle(inp)
    assert ref.n_fallback > 0 and res.stats["fallback_frames"] == ref.n_fallback
    assert_close_results(res, ref)
    sel, v = g["frame_sel"], g["det_isVoiced"]
    assert ((res.am[sel][v][:, 0] == 10e-4) == (g["det_am"][v][:, 0] == 10e-4)).all()
    assert (g["det_am"][v][:, 0] == 10e-4).sum() > 0


def assert_streamed_equals_mono(res, mono, tol_s=1e-9):
    assert len(res.SRER) == len(mono.SRER) and res.n_adapt_best == mono.n_adapt_best
    assert np.abs(np.array(res.SRER) - np.array(mono.SRER)).max() <= 1e-9
    assert np.abs(res.s_recon - mono.s_recon).max() <= tol_s
    v = mono.isVoiced
    assert (res.isVoiced == v).all() and ((res.am != 0) == (mono.am != 0)).all()
    assert ((res.fm != 0) == (mono.fm != 0)).all()
    if v.any():
        fmax = np.abs(mono.am[v]).max(axis=1, keepdims=True)
        assert (np.abs(res.am[v] - mono.am[v]) / fmax).max() <= TOL_REL
        strong = (mono.am[v] > 1e-3 * fmax) & (mono.fm[v] != 0)
        assert (np.abs(res.fm[v] - mono.fm[v])[strong] / mono.fm[v][strong]).max() <= TOL_REL
    assert np.abs(res.a0 - mono.a0).max() <= TOL_REL


def run_streamed(inp, shard_seconds, **kw):
    from eaqhm_b200 import longsignal
    eng = longsignal.LongSignalEngine(inp["s"], inp["fs"], inp["f0s"], inp["voiced5"], inp["frame_step"], inp["f0min"],
                                      inp["Kmax"], shard_seconds=shard_seconds, det_part=inp["det_part"],
                                      step=inp["step"], maxAdpt=inp["maxAdpt"], analysisWindow=inp["analysisWindow"], **kw)
    return eng.run(), eng


def test_frame_range_shards_equal_monolithic_run():
    """Config c2 mechanics: a signal streamed as three frame-range shards (per-frame parameter store, tracks rebuilt
    per shard from the store with an exchanged halo, one global SRER decision) must give the monolithic result: same
    adaptation count, SRER, waveform and parameters (to rounding: the halo truncates 0.268^32 spline influences)."""
    from eaqhm_b200 import synthetic
    inp = synth_inputs(synthetic.config_c2(duration=3.0), maxAdpt=3)
    mono = run_cuda(inp)
    res, eng = run_streamed(inp, 1.0)
    assert len(eng.shards) == 3 and res.stats["frames_solved"] == mono.stats["frames_solved"]   # no frame analysed twice
    assert_streamed_equals_mono(res, mono)
    one, _ = run_streamed(inp, 10.0)                     # a single shard keeps its tracks between the passes
    assert_streamed_equals_mono(one, mono)


@pytest.mark.parametrize("seed", [0, 1])
def test_streamed_shards_with_ragged_voicing_equal_monolithic_run(seed):
    """Shard boundaries through short runs, isolated frames and unvoiced holes: the short-run cubic keeps the
    reference's absolute sample coordinates (Q4, functions.py:369-371) inside an excerpt."""
    from eaqhm_b200 import host, synthetic
    rng = np.random.default_rng(300 + seed)
    cfg = synthetic.config_c3(60 + seed, duration=2.4)
    s, f0s = synthetic.make(cfg)
    fstep, step = 80, 80
    n5 = len(np.arange(1, len(s), fstep))
    voiced5 = np.ones(n5, bool)
    pos = 8
    while pos < n5 - 8:
        pos += int(rng.integers(2, 7))
        hole = int(rng.integers(1, 4))
        voiced5[pos:pos + hole] = False
        pos += hole
    det = host.voiced_only_target(s, voiced5, voiced5, fstep)
    inp = dict(s=s, fs=16000, f0s=f0s, voiced5=voiced5, frame_step=fstep, f0min=90.0,
               Kmax=host.harmonic_capacity(16000, f0s)[1], det_part=det, step=step, maxAdpt=3, analysisWindow=8)
    mono = run_cuda(inp)
    res, eng = run_streamed(inp, 0.5)
    assert len(eng.shards) >= 4
    assert_streamed_equals_mono(res, mono)


def test_streamed_shards_reproduce_the_fallback_quirks():
    """Q6/Q10 across shard boundaries: the 140 Hz fallback of a neighbour's frame is seen through an owned window, and
    its in-place amplitude lands in the best-adaptation store."""
    g, kw = load_golden("synth_silence")
    inp = inputs_from_golden(g, kw)
    mono = run_cuda(inp)
    assert mono.stats["fallback_frames"] > 0
    res, eng = run_streamed(inp, 0.15)
    assert len(eng.shards) >= 3
    assert_streamed_equals_mono(res, mono)
    assert ((res.am[:, 0] == 10e-4) == (mono.am[:, 0] == 10e-4)).all() and (mono.am[:, 0] == 10e-4).any()


def test_api_drop_in_on_sa19_with_supplied_preprocessing(tmp_path):
    """eaQHMAnalysisAndSynthesis (functions.py:35) from a wav path, f0 track and 5 ms flags supplied: the four return
    values have the reference's shapes and the published SRER."""
    from scipy.io import wavfile
    from eaqhm_b200.api import eaQHMAnalysisAndSynthesis
    g, kw = load_golden("sa19")
    wav = str(tmp_path / "SA19.WAV")
    wavfile.write(wav, int(g["fs"]), g["pcm"])
    s_recon, SRER, det, t = eaQHMAnalysisAndSynthesis(wav, "female", printPrompts=False, loadingScreen=False, f0s=g["f0s"],
                                                      frames=(g["speech5"], g["voiced5"], int(g["frame_step"])))
    assert s_recon.shape == (len(g["pcm"]),) and len(SRER) == 6 and t > 0
    assert np.abs(np.array(SRER) - g["SRER"]).max() <= 1e-8 and np.abs(s_recon - g["s_recon"]).max() <= TOL_S
    assert len(det) == 4233
    d = det[2000]
    assert d.isVoiced and d.ak == [] and d.amplitudes.dtype == object and d.amplitudes[0].shape == (1,)
    i = int(np.flatnonzero(g["frame_sel"] == 2000)[0])
    nk = int(g["det_nk"][i])
    assert len(d.amplitudes) == nk
    assert abs(float(d.amplitudes[0][0]) - g["det_am"][i, 0]) <= 1e-9 * g["det_am"][i].max()


def test_api_streams_long_signals_with_the_same_result():
    """analyze_arrays switches to the streamed long-signal engine when the dense tracks would be too large; forced
    here on SA19 (threshold 0, 1 s shards): the published SRER values and the reference's waveform come out."""
    from eaqhm_b200 import api
    g, kw = load_golden("sa19")
    s = g["pcm"].astype(float) / 32768
    res = api.analyze_arrays(s, int(g["fs"]), g["f0s"], g["speech5"], g["voiced5"], int(g["frame_step"]), "female",
                             stream_threshold_bytes=0, shard_seconds=1.0)
    assert res.stats["shards"] == 4 and len(res.SRER) == 6
    assert np.abs(np.array(res.SRER) - g["SRER"]).max() <= 1e-8 and np.abs(res.s_recon - g["s_recon"]).max() <= TOL_S
    out = compare_to_golden(res, g)
    assert out["flags_ok"] and out["support_equal"] and out["am_rel_framemax"] <= TOL_REL and out["fm_rel_strong"] <= TOL_REL


def test_api_drop_in_from_wav_only(tmp_path):
    """The full drop-in: wav path + gender, nothing supplied; host preprocessing (SWIPE', flags) restated in
    preprocess.py, hot path on the GPU.  Must give the published SA19 SRER values."""
    from scipy.io import wavfile
    from eaqhm_b200.api import eaQHMAnalysisAndSynthesis
    g, kw = load_golden("sa19")
    wav = str(tmp_path / "SA19.WAV")
    wavfile.write(wav, int(g["fs"]), g["pcm"])
    s_recon, SRER, det, t = eaQHMAnalysisAndSynthesis(wav, "female", printPrompts=False, loadingScreen=False)
    published = [17.86520945273994, 24.431728752204954, 24.67698055430504, 25.291120491477024, 25.497403658214047,
                 25.446628776435006]
    assert len(SRER) == 6 and np.abs(np.array(SRER) - np.array(published)).max() <= 1e-8
    assert np.abs(s_recon - g["s_recon"]).max() <= TOL_S and len(det) == 4233


@pytest.mark.parametrize("seed", [0, 1])
def test_ragged_voicing_short_runs_and_isolated_frames_match_oracle(seed):
    """Voiced-only mode with randomly punched voicing: runs of 1, 2, 3 and more frames per harmonic exercise the
    isolated-frame impulses (functions.py:350-358), the short-run cubic with knots at samples 0/step (Q4,
    functions.py:369-371), run boundaries inside analysis windows (gap fill, functions.py:250-278) and the
    shrinking active set."""
    from eaqhm_b200 import host, synthetic
    rng = np.random.default_rng(100 + seed)
    cfg = synthetic.config_c3(50 + seed, duration=1.2)
    s, f0s = synthetic.make(cfg)
    fstep, step = 80, 80          # hop = one 5 ms frame: a burst of b voiced 5 ms frames gives a run of b-1 analysis frames
    n5 = len(np.arange(1, len(s), fstep))
    voiced5 = np.ones(n5, bool)
    pos = 8
    while pos < n5 - 8:                                   # bursts of 2..6 voiced 5 ms frames, holes of 1..3
        burst = int(rng.integers(2, 7))
        pos += burst
        hole = int(rng.integers(1, 4))
        voiced5[pos:pos + hole] = False
        pos += hole
    det = host.voiced_only_target(s, voiced5, voiced5, fstep)
    inp = dict(s=s, fs=16000, f0s=f0s, voiced5=voiced5, frame_step=fstep, f0min=90.0,
               Kmax=host.harmonic_capacity(16000, f0s)[1], det_part=det, step=step, maxAdpt=3, analysisWindow=8)
    res, ref = run_cuda(inp), run_oracle(inp)
    assert_close_results(res, ref)
    v = ref.isVoiced
    runs = np.diff(np.concatenate(([0], v.astype(int), [0])))
    lens = np.flatnonzero(runs == -1) - np.flatnonzero(runs == 1)
    assert lens.min() == 1 and {2, 3} & set(lens.tolist()) and lens.max() >= 4   # isolated, short and long runs


@pytest.mark.parametrize("kw", [dict(step=10), dict(step=20, analysisWindow=24), dict(pitchPeriods=4),
                                dict(partials=20)])
def test_non_default_parameters_match_oracle(kw):
    """step / pitchPeriods / analysisWindow / partials (functions.py:35-37) other than the defaults."""
    from eaqhm_b200 import engine, host, synthetic
    cfg = synthetic.config_c3(77, duration=0.5)
    s, f0s = synthetic.make(cfg)
    n5 = len(np.arange(1, len(s), 80))
    step, aw, pp = kw.get("step", 15), kw.get("analysisWindow", 32), kw.get("pitchPeriods", 3)
    Kmax = host.harmonic_capacity(16000, f0s, kw.get("partials", 0))[1]
    u = engine.Utterance(s=s, fs=16000, f0s=f0s, voiced5=np.ones(n5, bool), frame_step=80, f0min=90.0, Kmax=Kmax)
    res = engine.Engine([u], step=step, maxAdpt=2, pitchPeriods=pp, analysisWindow=aw).run()[0]
    ref = O.eaqhm_oracle(s, 16000, f0s, np.ones(n5, bool), 80, 90.0, Kmax, step=step, maxAdpt=2, pitchPeriods=pp,
                         analysisWindow=aw)
    assert_close_results(res, ref)


def test_resynthesis_from_parameters_and_compact_format(tmp_path):
    """the dense parameter tensors -> compact file -> back -> waveform through eaqhm_interpolate +
    eaqhm_synthesize.  The returned phases are the reference's pm_fin (functions.py:402): at the last frame of a run that is
    the INTEGRATED phase, not the measured one, so the last hop of every run is rebuilt from slightly different knots (the
    sinusoidal correction of functions.py:565-569 is re-derived against its own end value: 0.4 % of that correction);
    everywhere else the waveform is the analysis' own s_recon to rounding."""
    from eaqhm_b200 import params
    g, kw = load_golden("synth_a_adpt3")
    inp = inputs_from_golden(g, kw)
    res = run_cuda(inp)
    path = str(tmp_path / "p.npz")
    params.save_result(path, res)
    back = params.load_result(path)
    for a, b in ((res.am, back.am), (res.fm, back.fm), (res.ph, back.ph), (res.a0 * res.isVoiced, back.a0)):
        assert (a == b).all()
    assert (back.isVoiced == res.isVoiced).all() and (back.isSpeech == res.isSpeech).all() and (back.ti == res.ti).all()
    assert back.SRER == [float(x) for x in res.SRER] and back.fs == 16000 and back.step == 15
    s = params.resynthesize(back)
    d = np.abs(s - res.s_recon)
    assert np.median(d) <= 1e-14 and (d <= 1e-12).mean() >= 0.8     # exact away from the run ends
    assert d.max() <= 2e-3 * np.abs(res.s_recon).max()
    srer = 20 * np.log10(np.std(res.s_recon) / np.std(s - res.s_recon))
    assert srer > 60.0


def test_api_returns_tensors_on_request(tmp_path):
    from scipy.io import wavfile
    from eaqhm_b200 import engine
    from eaqhm_b200.api import eaQHMAnalysisAndSynthesis
    g, kw = load_golden("synth_a_adpt1")
    wav = str(tmp_path / "a.wav")
    wavfile.write(wav, int(g["fs"]), g["pcm"])
    s_recon, SRER, res, t = eaQHMAnalysisAndSynthesis(wav, (90, 250), maxAdpt=1, printPrompts=False, f0s=g["f0s"],
                                                      frames=(g["speech5"], g["voiced5"], int(g["frame_step"])),
                                                      return_tensors=True)
    assert isinstance(res, engine.Result) and res.am.shape == (len(res.ti), int(g["Kmax"]))
    assert np.abs(s_recon - g["s_recon"]).max() <= TOL_S and np.abs(np.array(SRER) - g["SRER"]).max() <= TOL_SRER


# ---------------------------------------------------------------------------------------------
# preprocessing stages on the GPU: same pitch track and flags as the reference
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,plim", [("sa19", (160, 300)), ("synth_child_vo", (250, 500)), ("synth_low", (70, 180))])
def test_swipe_on_the_gpu_gives_the_cpu_pitch_track(name, plim):
    """csrc/preproc.cu (hand-written FFT -> ERB loudness -> kernel products -> interpolation -> argmax) against the numpy
    restatement of SWIPE.py:14-132, which tests/test_preprocess_cpu.py pins to the reference: identical pitch in every 1 ms
    frame (the refinement snaps it to a 1/768-octave grid), strengths to rounding."""
    from eaqhm_b200 import preprocess
    g, _ = load_golden(name)
    s = g["pcm"].astype(float) / 32768
    fs = int(g["fs"])
    cpu = preprocess.swipep(s, fs, list(plim))
    gpu = preprocess.swipep_gpu(s, fs, list(plim))
    assert cpu.shape == gpu.shape and (cpu[:, 0] == gpu[:, 0]).all()
    assert ((cpu[:, 1] == gpu[:, 1]) | (np.isnan(cpu[:, 1]) & np.isnan(gpu[:, 1]))).all()
    both = np.isfinite(cpu[:, 2]) & np.isfinite(gpu[:, 2])
    assert (np.isfinite(cpu[:, 2]) == np.isfinite(gpu[:, 2])).all() and np.abs(cpu[both, 2] - gpu[both, 2]).max() <= 1e-12


def test_sa19_pitch_from_the_gpu_is_the_reference_track():
    """The 5 ms f0 grid the hot path consumes (functions.py:111-113), from the GPU SWIPE', equals the one captured from the
    unmodified reference run (fixture sa19.npz)."""
    from eaqhm_b200 import preprocess
    g, _ = load_golden("sa19")
    s = g["pcm"].astype(float) / 32768
    grid = np.arange(0, len(s) - 1, 80) / 16000
    f0s = preprocess.getLinear(preprocess.swipep_gpu(s, 16000, [160, 300]), grid)
    assert f0s.shape == g["f0s"].shape and np.abs(f0s[:, 1] - g["f0s"][:, 1]).max() <= 1e-9


@pytest.mark.parametrize("name,gender", [("sa19", "female"), ("synth_child_vo", (250, 500)), ("synth_silence", (90, 250))])
def test_speech_voiced_flags_on_the_gpu_equal_the_reference(name, gender):
    """k_vuv_energy + thresholds + 4-tap median (functions.py:577-642, misc.py:184-206) against the flags the unmodified
    reference estimated for the fixture."""
    from eaqhm_b200 import preprocess
    g, _ = load_golden(name)
    s = g["pcm"].astype(float) / 32768
    speech, voiced, fstep = preprocess.voicedUnvoicedFrames_gpu(s, int(g["fs"]), gender)
    assert fstep == int(g["frame_step"])
    assert (np.asarray(speech, bool) == g["speech5"]).all() and (np.asarray(voiced, bool) == g["voiced5"]).all()
    sc, vc, _ = preprocess.voicedUnvoicedFrames(s, int(g["fs"]), gender)
    assert (speech == sc).all() and (voiced == vc).all()


def test_batched_elliptic_filter_on_the_gpu_equals_scipy_bitwise():
    """misc.py:167-182 (zero-phase 6th-order elliptic filter through scipy.signal.filtfilt) for a batch of
    utterances in one launch, in scipy's operation order: the same bits as the host call, for the high-pass the reference
    applies when fc > 0 (functions.py:90-91) and the two filters of its speech / voiced detector (functions.py:579,582)."""
    from eaqhm_b200 import preprocess
    rng = np.random.default_rng(4)
    sigs = [rng.standard_normal(n) * 0.1 + 0.05 for n in (4000, 777, 16000, 48001)]
    g, _ = load_golden("sa19")
    sigs.append(g["pcm"].astype(float) / 32768)
    for fs, fc, ftype in ((16000, 30, "highpass"), (16000, 1500, "lowpass"), (48000, 1000, "lowpass"), (16000, 70, "highpass")):
        ys = preprocess.ellipFilter_batch_gpu(sigs, fs, fc, ftype)
        for s, y in zip(sigs, ys):
            assert (y == preprocess.ellipFilter(s, fs, fc, ftype)).all()


# ---------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------
def test_windows_that_wrap_to_the_end_of_the_signal_match_the_oracle_and_are_refused_by_the_streamed_engine():
    """A window that starts before sample 0 is read from the END of the signal by the reference's fancy index
    (functions.py:196,244-245 with negative start): half window N = 533 > analysisWindow*step = 480 at f0 = 45 Hz.  (The
    reference only survives such a signal when its last frames are unvoiced - a window past the end raises IndexError -
    hence the voiced-only flags.)  The monolithic engine reproduces the wrap; a frame range cannot."""
    from eaqhm_b200 import host, longsignal, synthetic
    cfg = dict(fs=16000, duration=0.2, seed=31, f0_of_t=lambda t: 45.0 + 0 * t)
    s, f0s = synthetic.make(cfg)
    n5 = len(np.arange(1, len(s), 80))
    voiced5 = np.zeros(n5, bool)
    voiced5[:14] = True
    Kmax = host.harmonic_capacity(16000, f0s)[1]
    ft = host.FrameTables(len(s), 16000, f0s, voiced5, 80, Kmax)
    assert len(ft.wrapping_frames()) >= 3
    det = host.voiced_only_target(s, voiced5, voiced5, 80)
    inp = dict(s=s, fs=16000, f0s=f0s, voiced5=voiced5, frame_step=80, f0min=40.0, Kmax=Kmax, det_part=det, step=15,
               maxAdpt=1, analysisWindow=32)
    assert_close_results(run_cuda(inp), run_oracle(inp))
    with pytest.raises(ValueError):
        longsignal.LongSignalEngine(s, 16000, f0s, voiced5, 80, 40.0, Kmax, shard_seconds=0.1, det_part=det, maxAdpt=1)


def test_batch_with_an_utterance_without_voiced_frames():
    """Ragged batch: one utterance has no analysed frame at all (its reconstruction is the zero DC spline, SRER 0 dB, stop
    after the second adaptation like functions.py:394-396); its neighbours in the batch are not disturbed."""
    from eaqhm_b200 import engine, synthetic
    a = synth_inputs(synthetic.config_c3(5, duration=0.3), maxAdpt=2)
    b = synth_inputs(synthetic.config_c3(6, duration=0.25), maxAdpt=2)
    b["voiced5"] = np.zeros_like(b["voiced5"])
    ra, rb = engine.Engine([utterance(a), utterance(b)], maxAdpt=2).run()
    alone = run_cuda(a)
    assert ra.SRER == alone.SRER and (ra.s_recon == alone.s_recon).all()
    ob = run_oracle(b)
    assert len(rb.SRER) == len(ob.SRER) == 2 and np.abs(np.array(rb.SRER) - np.array(ob.SRER)).max() <= 1e-12
    assert not rb.isVoiced.any() and np.abs(rb.s_recon).max() == 0.0 and (rb.am == 0).all()


def test_fixed_work_mode_keeps_the_reference_result():
    """bench.py's fixed-work mode (run_device(fixed_adaptations=n): n adaptations, convergence stop disabled) must leave the
    result the data-dependent loop returns: same SRER values up to the stop, same best adaptation, same waveform, in both
    engines."""
    from eaqhm_b200 import engine, longsignal, synthetic
    inp = synth_inputs(synthetic.config_c2(duration=1.5), maxAdpt=10)
    a = engine.Engine([utterance(inp)], maxAdpt=10)
    ra = a.results(*a.run_device())[0]
    b = engine.Engine([utterance(inp)], maxAdpt=10)
    n = len(ra.SRER) + 2
    rb = b.results(*b.run_device(fixed_adaptations=n))[0]
    assert len(rb.SRER) == n and rb.SRER[:len(ra.SRER)] == ra.SRER and rb.n_adapt_best == ra.n_adapt_best
    assert (rb.s_recon == ra.s_recon).all() and (rb.am == ra.am).all() and (rb.ph == ra.ph).all()
    la = longsignal.LongSignalEngine(inp["s"], 16000, inp["f0s"], inp["voiced5"], 80, inp["f0min"], inp["Kmax"], shard_seconds=0.5)
    xa = la.results(*la.run_device())
    lb = longsignal.LongSignalEngine(inp["s"], 16000, inp["f0s"], inp["voiced5"], 80, inp["f0min"], inp["Kmax"], shard_seconds=0.5)
    xb = lb.results(*lb.run_device(fixed_adaptations=len(xa.SRER) + 2))
    assert xb.SRER[:len(xa.SRER)] == xa.SRER and xb.n_adapt_best == xa.n_adapt_best and (xb.s_recon == xa.s_recon).all()
    assert (xb.am == xa.am).all()
