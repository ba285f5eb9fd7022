#!/bin/bash
# A/B run of kernel variants on the GPU box: parity probe + throughput probe per library.
# usage: tools/scl_ab.sh <outfile> <variant names...>   ("main" = the in-tree library)
out=$1; shift
mkdir -p gpurun_out
: > $out
for v in "$@"; do
  if [ "$v" = main ]; then unset ES_B200_LIB; else export ES_B200_LIB=$PWD/echoseal_b200/_variants/lib_$v.so; fi
  echo "=== $v" >> $out
  timeout 300 python tools/scl_check.py >> $out 2>&1 || echo "CHECK rc=$?" >> $out
  timeout 300 python tools/scl_perf.py 151552 8 >> $out 2>&1 || echo "PERF rc=$?" >> $out
done
cat $out
