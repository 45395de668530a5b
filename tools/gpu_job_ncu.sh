#!/bin/bash
# ncu evidence of a build (run under gpurun, one GPU): launch list of the bench command, `--set full` captures of the
# executor programs and of the GEMM / convolution kernels of one config-2 train step; raw CSV pages land in gpurun_out/
tag=${1:-r02}
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,launch__registers_per_thread,launch__grid_size"
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches_bench.ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1; echo launches rc=$?; grep -c persist_kernel gpurun_out/${tag}_launches_bench.ncu.csv)
(AOCR_GRAPHS=0 STEP_N=3 STEP_DECODE=1 timeout 600 ncu --set full --clock-control none -k regex:persist_kernel -s 12 -c 9 -o gpurun_out/${tag}_executor -f python tools/one_step.py > gpurun_out/${tag}_executor_ncu.log 2>&1; echo executor rc=$?
 ncu -i gpurun_out/${tag}_executor.ncu-rep --page raw --csv > gpurun_out/${tag}_executor_raw.csv 2>/dev/null; ls -la gpurun_out/${tag}_executor.ncu-rep)
(AOCR_GRAPHS=0 STEP_N=2 timeout 600 ncu --metrics $M --clock-control none -k regex:"tc_gemm" -s 40 -c 40 --csv --log-file gpurun_out/${tag}_gemm_metrics.csv python tools/one_step.py > gpurun_out/${tag}_gemm_ncu.log 2>&1; echo gemm rc=$?)
rm -f gpurun_out/${tag}_executor.ncu-rep
