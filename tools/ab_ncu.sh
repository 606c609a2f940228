#!/bin/bash
# Light ncu pass (source counters + warp states + a few totals) per experiment library: tools/ab_ncu.sh NAME...
cd "$(dirname "$0")/.."
for n in "$@"; do
  OFDM_B200_LIB=$PWD/ofdm-course_b200/lib/exp/$n.so ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats \
    --metrics smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none --import-source on -k regex:rx4096 -s 3 -c 1 -f -o gpurun_out/ab_$n \
    python bench.py --streams 8192 --steps 2 --warmup 3 --no-cpu --e2e-streams 512 > gpurun_out/ab_ncu_$n.log 2>&1
  tail -1 gpurun_out/ab_ncu_$n.log
done
