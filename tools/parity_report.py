"""Parity of the CUDA path against every reference-made fixture, all metrics, next to the measured
reference-vs-reference noise floor (tests/golden/floors.json): prints the markdown table kept as profiles/r02_parity.md.

    python tools/parity_report.py [fixture ...] > profiles/r02_parity.md
"""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import FLOOR_FACTOR, KEYS, compare_to_golden, inputs_from_golden, load_floors, load_golden, tolerance  # noqa: E402
from eaqhm_b200 import engine  # noqa: E402

names = sys.argv[1:] or ["synth_a_adpt0", "synth_a_adpt1", "synth_a_adpt2", "synth_a_adpt3", "synth_low", "synth_child_vo",
                         "synth_48k", "synth_silence", "c2_2s", "c4_015", "sa19"]
floors = load_floors()
print("# CUDA path vs the reference's golden outputs, and the reference's own noise floor (round 2)\n")
print("Every cell: `CUDA-vs-reference / reference-vs-reference floor` (floor = the unmodified reference re-run with one "
      "OpenBLAS thread against its own 8-thread run, tests/golden/make_floor.py). Test tolerance per metric = "
      "max(north-star tolerance, %g x floor).\n" % FLOOR_FACTOR)
print("| fixture | adaptations | " + " | ".join(KEYS) + " |")
print("|---|---:|" + "---:|" * len(KEYS))
worst = {}
for name in names:
    if not os.path.exists(os.path.join(ROOT, "tests", "golden", name + ".npz")):
        continue
    g, kw = load_golden(name)
    inp = inputs_from_golden(g, kw)
    u = engine.Utterance(s=inp["s"], fs=inp["fs"], f0s=inp["f0s"], voiced5=inp["voiced5"], frame_step=inp["frame_step"],
                         f0min=inp["f0min"], Kmax=inp["Kmax"], det_part=inp["det_part"])
    res = engine.Engine([u], step=inp["step"], maxAdpt=inp["maxAdpt"], analysisWindow=inp["analysisWindow"]).run()[0]
    c = compare_to_golden(res, g)
    fl = floors.get(name, {})
    cells = []
    for k in KEYS:
        ok = c[k] <= tolerance(name, k)
        cells.append("%.1e / %s%s" % (c[k], ("%.1e" % fl[k]) if k in fl else "-", "" if ok else " **FAIL**"))
        worst[k] = max(worst.get(k, 0.0), c[k])
    print("| %s | %d | %s |" % (name, len(res.SRER), " | ".join(cells)))
print("\nWorst over fixtures: " + ", ".join("%s %.1e" % (k, v) for k, v in worst.items()))
print("\n```json\n" + json.dumps({"floor_factor": FLOOR_FACTOR, "worst": worst}) + "\n```")
