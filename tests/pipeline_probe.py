"""Helper of test_chunk_pipeline_equals_serial_chunks_bitwise: one c2 excerpt through the CUDA path in THIS process
(the library reads EAQHM_PIPELINE once per process), result digests as one JSON line."""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from eaqhm_b200 import _lib, engine, synthetic

cfg = synthetic.config_c2(duration=1.5)
s, f0s = synthetic.make(cfg)
n5 = len(np.arange(1, len(s), 80))
u = engine.Utterance(s=s, fs=16000, f0s=f0s, voiced5=np.ones(n5, bool), frame_step=80, f0min=90.0,
                     Kmax=int(round(7800 / f0s[:, 1].min()) + 10))
eng = engine.Engine([u], maxAdpt=3, chunk=int(sys.argv[1]) if len(sys.argv) > 1 else None)
r = eng.run()[0]
h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
print(json.dumps({"sets": _lib.lib().eaqhm_analysis_pipeline_sets(), "chunk": eng.chunk, "frames": int(eng.tab.nA),
                  "SRER": [float(x) for x in r.SRER], "s": h(r.s_recon), "am": h(r.am), "fm": h(r.fm), "ph": h(r.ph),
                  "a0": h(r.a0)}))
