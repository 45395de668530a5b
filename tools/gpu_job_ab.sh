#!/bin/bash
# A/B of environment switches: config-2 (and optionally config-4) bench + phase breakdown for each setting
# usage: gpu_job_ab.sh <tag> "<ENV=.. ENV=..>" ["<ENV..>" ...]     ("-" = no switches)
tag=${1:-ab}; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  [ "$envs" = "-" ] && envs=""
  echo "== [$i] ${envs:-default}"
  (env $envs AOCR_PHASES=1 AOCR_GRAPHS=0 STEP_N=3 timeout 100 python tools/one_step.py 2>&1 | grep phases | tail -1)
  for c in ${AB_CONFIGS:-2}; do
    (env $envs timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_${i}_c$c.log 2>&1
     python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_${i}_c$c.log").read().strip().splitlines()[-1])
o=d["roofline"].get("other_class",{})
print("  config $c: train ms", round(d["ms_per_step"],3), "decode ms", round(d.get("decode",{}).get("ms_per_batch",0),3), "| conv/GEMM class", round(o.get("achieved",0),1), "TFLOP/s, ms in class", round(o.get("ms_per_step_in_class",0),3))
PY
    )
  done
  i=$((i+1))
done
