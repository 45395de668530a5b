#!/bin/bash
# multi-GPU check (run under gpurun --gpus N): data-parallel parity against the single-device oracle, the bench lines of
# config 2 and config 4 at N ranks, and the per-phase breakdown of a data-parallel step
N=${1:-2}; tag=${2:-dp}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
(timeout 300 $TR tools/dp_check.py > gpurun_out/${tag}_dpcheck_n$N.log 2>&1; echo dp_check rc=$?; grep -E "world|grad L2" gpurun_out/${tag}_dpcheck_n$N.log)
for c in 2 4; do
  (timeout 300 $TR bench.py --gpus $N --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_c${c}_n$N.log 2>&1; echo bench c$c rc=$?
   grep '"metric"' gpurun_out/${tag}_bench_c${c}_n$N.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  config $c N=$N:', round(d['value']), d['unit'], 'ms/step', round(d['ms_per_step'],3))")
  (AOCR_PHASES=1 timeout 300 $TR bench.py --gpus $N --config $c --steps 4 --warmup 2 --no-cpu-baseline 2>&1 | grep "aocr phases" | tail -2)
done
