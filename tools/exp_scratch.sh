#!/bin/bash
for g in 6 16 40 80; do
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --scratch-gb $g 2>&1 | tail -1 > gpurun_out/bench_tmp.json
  python - "$g" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_tmp.json")); kt = d["roofline"]["kernel_ms_per_step"]
print("scratch %s GB: step %.1f ms  chol %.1f gram %.1f basis %.1f  e2e %.1f ms launches %d" % (sys.argv[1], d["ms_per_step"], kt["chol_solve"], kt["gram"], kt["basis"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
PY
done
