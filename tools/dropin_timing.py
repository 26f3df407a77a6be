"""Wall time of the drop-in call on SA19.WAV (wav path + gender in, everything out): GPU preprocessing vs the numpy
restatements, tensor-form result vs the reference's list of Deterministic objects.  Needs tests/golden/sa19.npz."""
import os, sys, tempfile, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scipy.io import wavfile
from common import load_golden
from eaqhm_b200 import preprocess
from eaqhm_b200.api import eaQHMAnalysisAndSynthesis

g, _ = load_golden("sa19")
td = tempfile.mkdtemp()
wav = os.path.join(td, "SA19.WAV")
wavfile.write(wav, int(g["fs"]), g["pcm"])
s = g["pcm"].astype(float) / 32768
for label, kw in (("gpu preprocessing, tensors", dict(return_tensors=True)),
                  ("gpu preprocessing, Deterministic list", dict()),
                  ("host preprocessing, tensors", dict(return_tensors=True, host_preprocessing=True)),
                  ("host preprocessing, Deterministic list", dict(host_preprocessing=True))):
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        s_recon, SRER, det, _ = eaQHMAnalysisAndSynthesis(wav, "female", printPrompts=False, loadingScreen=False, **kw)
        ts.append(time.perf_counter() - t0)
    print("%-42s %s ms   SRER[-2] %.9f" % (label, " ".join("%.1f" % (1e3 * t) for t in ts), SRER[-2]))
for name, fn in (("swipep (numpy)", lambda: preprocess.swipep(s, 16000, [160, 300])),
                 ("swipep_gpu", lambda: preprocess.swipep_gpu(s, 16000, [160, 300])),
                 ("voicedUnvoicedFrames (numpy)", lambda: preprocess.voicedUnvoicedFrames(s, 16000, "female")),
                 ("voicedUnvoicedFrames_gpu", lambda: preprocess.voicedUnvoicedFrames_gpu(s, 16000, "female"))):
    ts = []
    for rep in range(4):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print("%-42s %s ms" % (name, " ".join("%.1f" % (1e3 * t) for t in ts)))
