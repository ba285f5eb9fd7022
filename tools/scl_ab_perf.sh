#!/bin/bash
# throughput-only A/B run of kernel variants: tools/scl_ab_perf.sh <outfile> <variant names...>
out=$1; shift
mkdir -p gpurun_out
: > $out
for v in "$@"; do
  if [ "$v" = main ]; then unset ES_B200_LIB; else export ES_B200_LIB=$PWD/echoseal_b200/_variants/lib_$v.so; fi
  echo "=== $v" >> $out
  timeout 300 python tools/scl_perf.py 151552 8 2>&1 | grep -v "^list_decode.*" | tail -4 >> $out || echo "PERF rc=$?" >> $out
done
cat $out
