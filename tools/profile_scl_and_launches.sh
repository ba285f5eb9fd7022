#!/bin/bash
# Round evidence, short form: launch list of the default bench command + ncu --set full of scl_list_kernel (detector pairing).
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/prof_bench_plain.json 2> gpurun_out/prof_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/prof_bench_ncu.json 2> gpurun_out/prof_bench_ncu.err
echo "launch list rc=$?"
bash tools/scl_prof.sh main
