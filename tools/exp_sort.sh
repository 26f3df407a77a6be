#!/bin/bash
for g in 0 1; do
  EAQHM_SORT_FRAMES=$g python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_tmp.json
  python - "$g" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_tmp.json")); kt = d["roofline"]["kernel_ms_per_step"]
print("sort %s: step %.1f ms  chol %.1f gram %.1f basis %.1f  e2e %.1f ms" % (sys.argv[1], d["ms_per_step"], kt["chol_solve"], kt["gram"], kt["basis"], d["e2e"]["ms_per_step"]), d["config"]["SRER_dB"][-1])
PY
done
