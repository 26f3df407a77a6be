#!/bin/bash
# usage: tools/quick_bench.sh [shard_seconds]  -- runs the GPU parity tests (fail fast) and a short bench, prints the kernel table
S=${1:-5}
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 2 --warmup 1 --shard-seconds $S --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_tmp.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_tmp.json"))
print("value %.2f audio-s/s  ms/step %.1f  e2e %.2f  adapt %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["adaptations_run"]))
r = d["roofline"]
print({k: round(v, 2) for k, v in r["kernel_ms_per_step"].items() if v > 0.2})
fl = r["algorithmic_flops_per_step"]; kt = r["kernel_ms_per_step"]
print("gram %.1f TF/s  chol %.1f TF/s  (useful)" % (fl["gram"] / kt["gram"] / 1e9, fl["chol"] / kt["chol_solve"] / 1e9))
PY
