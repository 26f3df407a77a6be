#!/bin/bash
for w in 0 1 2 3; do
  EAQHM_CHOL_VARIANT=$w python bench.py --steps 2 --warmup 1 --shard-seconds 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_tmp.json
  python - "$w" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_tmp.json")); kt = d["roofline"]["kernel_ms_per_step"]
print("variant %s: chol %.1f ms  gram %.1f  basis %.1f step %.1f  adapt %d" % (sys.argv[1], kt["chol_solve"], kt["gram"], kt["basis"], d["ms_per_step"], d["config"]["adaptations_run"]), d["config"]["SRER_dB"][-1])
PY
done
