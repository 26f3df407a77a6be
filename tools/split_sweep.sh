#!/bin/bash
# Size split of the Cholesky (one warp per frame up to P = EAQHM_CHOL_SPLIT_P): small-system configurations and the c2 step
O=gpurun_out/split_sweep.jsonl; : > $O
for p in 0 96 128 160 192; do EAQHM_CHOL_SPLIT_P=$p python tools/small_p_timing.py >> $O 2>>gpurun_out/split_sweep.err; done
for p in 0 128 160; do echo "# c2 split $p" >> $O; EAQHM_CHOL_SPLIT_P=$p python tools/kernel_times.py 20 1 >> $O 2>>gpurun_out/split_sweep.err; done
echo "# c2 real warps 1" >> $O; EAQHM_CHOL_REAL_WARPS=1 python tools/kernel_times.py 20 1 >> $O 2>>gpurun_out/split_sweep.err
EAQHM_CHOL_REAL_WARPS=1 python tools/small_p_timing.py >> $O 2>>gpurun_out/split_sweep.err
