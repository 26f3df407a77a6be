#!/bin/bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for cfg in "0 2" "1 1" "1 2"; do
  set -- $cfg
  EAQHM_OVERLAP=$1 EAQHM_GRAM_CTAS_OVERLAP=$2 python bench.py --steps 2 --warmup 1 --shard-seconds 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_tmp.json
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_tmp.json")); kt = d["roofline"]["kernel_ms_per_step"]
print("overlap %s gramCTAs %s: step %.1f ms  (events: chol %.1f gram %.1f basis %.1f)  e2e %.1f  adapt %d" % (sys.argv[1], sys.argv[2], d["ms_per_step"], kt["chol_solve"], kt["gram"], kt["basis"], d["e2e"]["ms_per_step"], d["config"]["adaptations_run"]), d["config"]["SRER_dB"][-1])
PY
done
