#!/bin/bash
# Full GPU check of a build (run under gpurun): per-command trace, CNN-gradient diagnostics, the GPU test-suite, the
# driver's ncu pass over smoke(), and the bench lines of BASELINE configs 2-5.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-run}
mkdir -p gpurun_out
(AOCR_PERSIST_TRACE=1 AOCR_GRAPHS=0 STEP_N=2 STEP_DECODE=1 timeout 120 python tools/one_step.py > gpurun_out/${tag}_trace.log 2>&1; echo trace rc=$?)
(DIAG_B=64 DIAG_MODES=0 DIAG_VERBOSE=1 timeout 600 python tools/diag_cnn_grads.py > gpurun_out/${tag}_diag_cnn64.log 2>&1; grep "B=64\|dsrc rel" gpurun_out/${tag}_diag_cnn64.log)
(DIAG_B=8 DIAG_MODES=0,2 timeout 600 python tools/diag_cnn_grads.py > gpurun_out/${tag}_diag_cnn8.log 2>&1; grep "B=8\|dsrc rel" gpurun_out/${tag}_diag_cnn8.log)
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo pytest rc=$?; tail -25 gpurun_out/${tag}_pytest.log)
(timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/${tag}_smoke_ncu.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke_ncu.log 2>&1; echo ncu rc=$?; tail -3 gpurun_out/${tag}_smoke_ncu.log; grep -c persist_kernel gpurun_out/${tag}_smoke_ncu.csv)
for c in 2 3 4 5; do
  (timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c$c.log 2>&1; echo bench c$c rc=$?; tail -1 gpurun_out/${tag}_bench_c$c.log | cut -c1-300)
done
