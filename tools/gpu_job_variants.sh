#!/bin/bash
# A/B of executor body variants (AOCR_VARIANT bits, dec_bodies.cuh): per-command trace + config-2 bench for each value
tag=${1:-var}; shift
mkdir -p gpurun_out
for v in "$@"; do
  (AOCR_VARIANT=$v AOCR_PERSIST_TRACE=1 AOCR_GRAPHS=0 STEP_N=2 STEP_DECODE=1 timeout 120 python tools/one_step.py > gpurun_out/${tag}_trace_v$v.log 2>&1
   echo "== variant $v"; grep -E "persist trace\] (80|121|2[0-9][0-9]) cmds|body stamps" gpurun_out/${tag}_trace_v$v.log | cut -c1-520 | head -4)
  (AOCR_VARIANT=$v timeout 300 python bench.py --config 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_v$v.log 2>&1
   python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_v$v.log").read().strip().splitlines()[-1])
print("variant $v: train ms", round(d["ms_per_step"],3), "decode ms", round(d["decode"]["ms_per_batch"],3))
PY
  )
done
