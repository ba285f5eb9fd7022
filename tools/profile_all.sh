#!/bin/bash
# Round evidence: (1) launch list of the default bench command, (2) ncu --set full of every shipped RX kernel on one
# 300-clip sub-batch, (3) ncu --set full of scl_list_kernel in the detector pairing (tools/scl_prof.sh).
# Every profiled command first runs plain (exit code checked) as the profiling guide asks.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/prof_bench_plain.json 2> gpurun_out/prof_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/prof_bench_ncu.json 2> gpurun_out/prof_bench_ncu.err
echo "launch list rc=$?"
python tools/rx_stage_times.py 300 > gpurun_out/prof_rx_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bandpass|ncc_|peaks2|frames_kernel|llr_kernel|scl_hard|collect_hits|resample' -c 12 -f \
    -o gpurun_out/rx_kernels_r02 python tools/rx_stage_times.py 300 > gpurun_out/prof_rx_ncu.log 2>&1
echo "rx kernels rc=$?"
bash tools/scl_prof.sh main
