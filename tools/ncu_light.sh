#!/bin/bash
# Light ncu pass (source counters + warp states + totals) of one kernel of a script: tools/ncu_light.sh KERNEL_REGEX SKIP OUT script.py [args...]
cd "$(dirname "$0")/.."
k=$1; skip=$2; out=$3; shift 3
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section MemoryWorkloadAnalysis \
  --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/$out python "$@" > gpurun_out/ncu_$out.log 2>&1
tail -1 gpurun_out/ncu_$out.log
