#!/bin/bash
# DRAM bytes and L2 hit rate of the 5th scl_list launch (detector pairing) per variant: tools/scl_dram.sh <variants...>
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = main ]; then unset ES_B200_LIB; else export ES_B200_LIB=$PWD/echoseal_b200/_variants/lib_$v.so; fi
  echo "=== $v"
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none \
      -k regex:scl_list -s 4 -c 1 python tools/scl_perf.py 37888 8 2>&1 | grep -E "dram__|lts__|gpu__time"
done
