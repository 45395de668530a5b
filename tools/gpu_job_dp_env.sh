#!/bin/bash
# bench lines of config 2 / 4 at N ranks for several environment settings (e.g. NCCL channel limits)
N=${1:-2}; tag=${2:-dpenv}; shift; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
i=0
for envs in "$@"; do
  [ "$envs" = "-" ] && envs=""
  for c in ${DP_CONFIGS:-2 4}; do
    (env $envs timeout 300 $TR bench.py --gpus $N --config $c --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_${i}_c${c}_n$N.log 2>&1
     grep '"metric"' gpurun_out/${tag}_${i}_c${c}_n$N.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[${envs:-default}] config $c N=$N:', round(d['value']), d['unit'], 'ms/step', round(d['ms_per_step'],3))")
  done
  i=$((i+1))
done
