"""The five BASELINE.json configurations on one GPU at their named shapes where one GPU holds them (c1 whole file,
c4 one full second at 48 kHz, c3 / c5 through the corpus runner on `--utterances` utterances of the 10,000, c2 as one 20 s
frame range; the whole hour: tools/run_full_hour.py).  Device time of the adaptation loop (CUDA events, second run,
inputs resident), data-dependent stop like the reference.  Prints a markdown table and one JSON line per configuration.

    python tools/config_timings.py [--utterances 256] [--c4-seconds 1.0]
"""
import argparse, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from common import load_golden, inputs_from_golden
from eaqhm_b200 import _lib, corpus, engine, host, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--utterances", type=int, default=256)
ap.add_argument("--c4-seconds", type=float, default=1.0)
args = ap.parse_args()
L = _lib.lib()


def timed(utts, **kw):
    eng = engine.Engine(utts, **kw)
    eng.run_device()
    L.eaqhm_profile_reset(); L.eaqhm_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); ev0.record()
    SRER, best, solves = eng.run_device()
    ev1.record(); torch.cuda.synchronize()
    L.eaqhm_profile_enable(0)
    prof = {k: round(v[0], 2) for k, v in _lib.profile_read().items() if v[0] >= 0.5}
    fl = eng.useful_flops
    return ev0.elapsed_time(ev1) * 1e-3, SRER, solves, prof, fl


def corpus_timed(n, load, **kw):
    """Corpus runner end to end (host preprocessing in threads, batches sized by memory, results on the host)."""
    t0 = time.perf_counter()
    srer, solves, audio = [], 0, 0.0
    for i, r in corpus.analyze_corpus(n, load, **kw):
        srer.append(r.SRER)
        solves += int(r.isVoiced.sum()) * len(r.SRER)
        audio += len(r.s_recon) / r.fs
    torch.cuda.synchronize()
    return time.perf_counter() - t0, srer, solves, audio


rows = []
g, kw = load_golden("sa19")
inp = inputs_from_golden(g, kw)
u = engine.Utterance(s=inp["s"], fs=inp["fs"], f0s=inp["f0s"], voiced5=inp["voiced5"], frame_step=inp["frame_step"],
                     f0min=inp["f0min"], Kmax=inp["Kmax"])
t, S, n, prof, fl = timed([u])
rows.append(("c1 SA19.WAV whole (3.97 s, female, defaults)", len(inp["s"]) / 16000, t, S, n, prof, fl, "device"))
t, S, n, prof, fl = timed([engine.Utterance(**bench.long_signal_inputs(20.0))])
rows.append(("c2 first 20 s of the 1-hour AM-FM signal (whole hour: tools/run_full_hour.py)", 20.0, t, S, n, prof, fl, "device"))

c3_cache = {}
def load_c3(i):
    return engine.Utterance(**bench.corpus_inputs(i, 1)[0])
t, S, n, audio = corpus_timed(args.utterances, load_c3)
rows.append(("c3 %d of the 10,000 3-second utterances, corpus runner (wall, synthesis of the inputs and read-back included)"
             % args.utterances, audio, t, S, n, {}, {}, "wall"))

cfg = synthetic.config_c4(duration=args.c4_seconds)
s, f0s = synthetic.make(cfg)
n5 = len(np.arange(1, len(s), 240))
u4 = engine.Utterance(s=s, fs=48000, f0s=f0s, voiced5=np.ones(n5, bool), frame_step=240, f0min=70.0,
                      Kmax=host.harmonic_capacity(48000, f0s)[1])
t, S, n, prof, fl = timed([u4], analysisWindow=64, scratch_budget_bytes=80 << 30)
rows.append(("c4 48 kHz f0~80 Hz (P=1190, L=1801), %.2f s, analysisWindow=64, default maxAdpt" % args.c4_seconds,
             args.c4_seconds, t, S, n, prof, fl, "device"))


def load_c5(i):
    cfg = synthetic.config_c5(i, duration=3.0)
    s, f0s = synthetic.make(cfg)
    t5 = (np.arange(1, len(s), 80) - 1) / 16000
    v5 = np.asarray(cfg["voiced_mask_of_t"](t5), bool)
    return engine.Utterance(s=s, fs=16000, f0s=f0s, voiced5=v5, frame_step=80, f0min=250.0,
                            Kmax=host.harmonic_capacity(16000, f0s)[1], det_part=host.voiced_only_target(s, v5, v5, 80))
t, S, n, audio = corpus_timed(args.utterances, load_c5)
rows.append(("c5 voiced-only, child-range f0, %d utterances x 3 s, corpus runner (wall)" % args.utterances, audio, t, S, n,
             {}, {}, "wall"))

print("| config | audio s | seconds (device or wall) | audio-s/s | frame solves | solves/s | adaptations | mean best SRER dB | TFLOP/s (conventional count) |")
print("|---|---:|---:|---:|---:|---:|---|---:|---:|")
for name, a, t, S, n, prof, fl, kind in rows:
    ad = sorted(set(len(x) for x in S))
    tf = (fl.get("gram", 0) + fl.get("chol", 0) + fl.get("basis", 0)) / t / 1e12 if fl else float("nan")
    print("| %s | %.2f | %.4f (%s) | %.1f | %d | %.0f | %s | %.2f | %.1f |"
          % (name, a, t, kind, a / t, n, n / t, "/".join(map(str, ad)) if len(ad) < 8 else "%d..%d" % (ad[0], ad[-1]),
             float(np.mean([max(x) for x in S])), tf))
print()
for name, a, t, S, n, prof, fl, kind in rows:
    print(json.dumps({"config": name, "audio_s": a, "seconds": t, "timing": kind, "audio_s_per_s": a / t, "frame_solves": int(n),
                      "kernel_ms": prof, "algorithmic_flops": fl}))
