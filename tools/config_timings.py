"""One pass over the five BASELINE.json configurations (scaled to a few seconds each) on one GPU: device time of the
adaptation loop (CUDA events, second run), audio-s/s, adaptations, best SRER.  Prints a markdown table."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from common import load_golden, inputs_from_golden
from eaqhm_b200 import engine, host, synthetic


def timed(utts, **kw):
    eng = engine.Engine(utts, **kw)
    eng.run_device()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); ev0.record()
    SRER, best, solves = eng.run_device()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) * 1e-3, SRER, solves, eng


rows = []
g, kw = load_golden("sa19")
inp = inputs_from_golden(g, kw)
u = engine.Utterance(s=inp["s"], fs=inp["fs"], f0s=inp["f0s"], voiced5=inp["voiced5"], frame_step=inp["frame_step"],
                     f0min=inp["f0min"], Kmax=inp["Kmax"])
t, S, n, e = timed([u])
rows.append(("c1 SA19.WAV (3.97 s, female, defaults)", len(inp["s"]) / 16000, t, S, n, e))
t, S, n, e = timed([engine.Utterance(**bench.shard_inputs(0, 10.0))])
rows.append(("c2 1-hour AM-FM signal, one 10 s shard", 10.0, t, S, n, e))
utts = [engine.Utterance(**i) for i in bench.corpus_inputs(0, 16)]
t, S, n, e = timed(utts)
rows.append(("c3 16 utterances x 3 s", 48.0, t, S, n, e))
cfg = synthetic.config_c4(duration=0.5)
s, f0s = synthetic.make(cfg)
n5 = len(np.arange(1, len(s), 240))
u4 = engine.Utterance(s=s, fs=48000, f0s=f0s, voiced5=np.ones(n5, bool), frame_step=240, f0min=70.0,
                      Kmax=host.harmonic_capacity(48000, f0s)[1])
t, S, n, e = timed([u4], analysisWindow=64, maxAdpt=2, scratch_budget_bytes=60 << 30)
rows.append(("c4 48 kHz f0~80 Hz (P=1190, L=1801), 0.5 s, maxAdpt=2", 0.5, t, S, n, e))
utts5 = []
for i in range(8):
    cfg = synthetic.config_c5(i, duration=3.0)
    s, f0s = synthetic.make(cfg)
    t5 = (np.arange(1, len(s), 80) - 1) / 16000
    v5 = np.asarray(cfg["voiced_mask_of_t"](t5), bool)
    utts5.append(engine.Utterance(s=s, fs=16000, f0s=f0s, voiced5=v5, frame_step=80, f0min=250.0,
                                  Kmax=host.harmonic_capacity(16000, f0s)[1], det_part=host.voiced_only_target(s, v5, v5, 80)))
t, S, n, e = timed(utts5)
rows.append(("c5 voiced-only, child-range f0, 8 utterances x 3 s", 24.0, t, S, n, e))
print("| config | audio s | device s | audio-s/s | frame solves | solves/s | adaptations | best SRER dB |")
print("|---|---:|---:|---:|---:|---:|---|---:|")
for name, a, t, S, n, e in rows:
    ad = sorted(set(len(x) for x in S))
    print("| %s | %.2f | %.4f | %.1f | %d | %.0f | %s | %.2f |" % (name, a, t, a / t, n, n / t, "/".join(map(str, ad)),
                                                                  float(np.mean([max(x) for x in S]))))
