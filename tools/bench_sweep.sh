#!/bin/bash
# bench.py at several sub-batch sizes: tools/bench_sweep.sh 1000 2000 ...
for sb in "$@"; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-secondary --sub-batch $sb 2>/dev/null > gpurun_out/sweep_$sb.json
  python - "$sb" <<'PY'
import sys, json
sb = sys.argv[1]
d = json.loads(open(f"gpurun_out/sweep_{sb}.json").read().strip().split("\n")[-1])
print("sub_batch", sb, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), {k: round(v, 1) for k, v in d["kernel_ms_per_step"].items()})
PY
done
