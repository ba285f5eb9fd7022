#!/bin/bash
# ncu --set full capture of the 5th scl_list launch (detector pairing) of tools/scl_perf.py for each variant.
# usage: tools/scl_prof.sh <variant names...>
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = main ]; then unset ES_B200_LIB; else export ES_B200_LIB=$PWD/echoseal_b200/_variants/lib_$v.so; fi
  timeout 300 python tools/scl_perf.py 37888 8 > gpurun_out/prof_plain_$v.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:scl_list -s 4 -c 1 -f -o gpurun_out/scl_$v \
      python tools/scl_perf.py 37888 8 > gpurun_out/prof_ncu_$v.log 2>&1
  echo "$v rc=$?"; tail -3 gpurun_out/prof_plain_$v.log
done
