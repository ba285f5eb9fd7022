#!/bin/bash
# ncu --set full of every shipped RX kernel on one 300-clip sub-batch (plain run first, as the profiling guide asks)
mkdir -p gpurun_out
python tools/rx_stage_times.py 300 > gpurun_out/prof_rx_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bandpass|ncc_|peaks2|frames_kernel|llr_kernel|scl_hard|collect_hits|resample' -c 12 -f \
    -o gpurun_out/rx_kernels_r02 python tools/rx_stage_times.py 300 > gpurun_out/prof_rx_ncu.log 2>&1
echo "rx kernels rc=$?"
