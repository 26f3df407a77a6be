#!/bin/bash
# Kernel-variant comparison on the c2 default step: tools/variant_sweep.sh out.jsonl name1 name2 ...  (build/lib_<name>.so; "base" = in-tree)
out=$1; shift
: > $out
for v in "$@"; do
  if [ "$v" = base ]; then python tools/kernel_times.py 20 2 >> $out 2>>${out%.jsonl}.err
  else EAQHM_LIB=build/lib_$v.so python tools/kernel_times.py 20 2 >> $out 2>>${out%.jsonl}.err; fi
done
