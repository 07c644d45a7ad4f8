#!/bin/bash
# Completes tools/r2_ncu.sh for the kernels whose capture step timed out in the final call: block_jacobi_kernel (cfg5) and
# dense_matvec_kernel (the coarse pinv apply), same metric list, same library.
mkdir -p gpurun_out/ncu
O=gpurun_out/ncu
cut -c1-16 pyamg_b200/libpyamg_b200.so.sha256 > $O/so_sha16_cfg5.txt
NCU="ncu --clock-control none"
MET="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__throughput.avg.pct_of_peak_sustained_elapsed"
AMGB_NCU_SELECT="0:5:1,1:5:1,0:1:1,0:0:1,0:2:1" timeout 200 $NCU --metrics $MET --profile-from-start off -o $O/cfg5 python tools/ncu_targets.py cfg5 > $O/cfg5.log 2>&1
timeout 120 $NCU --metrics $MET -k regex:dense_matvec_kernel -c 1 -o $O/dense python tools/ncu_targets.py cfg5 > $O/dense.log 2>&1
for f in cfg5 dense; do [ -f $O/$f.ncu-rep ] && ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.raw.csv 2>/dev/null; done
rm -f $O/*.ncu-rep
ls -la $O | tail -8; tail -2 $O/cfg5.log $O/dense.log
