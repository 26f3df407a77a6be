#!/bin/bash
# Chunk-pipeline experiments on the c2 default step: tools/pipe_sweep.sh <out.jsonl>  (one JSON line per variant)
out=${1:-gpurun_out/pipe_sweep.jsonl}
: > $out
run() { echo "# $*" >> $out; env "$@" python tools/kernel_times.py 20 2 >> $out 2>>${out%.jsonl}.err; }
rm -f gpurun_out/trace_pipe3.txt
run EAQHM_PIPELINE=2 EAQHM_TRACE=gpurun_out/trace_pipe3.txt
run EAQHM_PIPELINE=2 EAQHM_PIPE_CHUNKS=4
run EAQHM_PIPELINE=2 EAQHM_PIPE_CHUNKS=9
run EAQHM_PIPELINE=2 EAQHM_PIPE_CHUNKS=12 EAQHM_PIPE_MIN_CHUNK=1200
run EAQHM_PIPELINE=2 EAQHM_PIPE_HELPERS=0
run EAQHM_PIPELINE=2 EAQHM_PIPE_CHOL_CTAS=3
run EAQHM_PIPELINE=2 EAQHM_PIPE_GRAM_CTAS=2 EAQHM_PIPE_CHOL_CTAS=8
