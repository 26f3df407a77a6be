#!/bin/bash
# Timeline of the chunk pipeline: tools/pipe_trace.sh  -> gpurun_out/trace_<variant>.txt (class stream t0 t1 per launch)
run() { name=$1; shift; rm -f gpurun_out/trace_$name.txt; env EAQHM_TRACE=gpurun_out/trace_$name.txt "$@" python tools/kernel_times.py 20 1 >> gpurun_out/pipe_trace.jsonl 2>>gpurun_out/pipe_trace.err; }
: > gpurun_out/pipe_trace.jsonl
run serial EAQHM_PIPELINE=1
run pipe EAQHM_PIPELINE=2
run nohelper EAQHM_PIPELINE=2 EAQHM_PIPE_HELPERS=0
run prio0 EAQHM_PIPELINE=2 EAQHM_PIPE_PRIO=0
