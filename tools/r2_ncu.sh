#!/bin/bash
# Round-2 ncu evidence for the kernels AS SHIPPED.  One gpurun call, every step bounded.
# The engine brackets chosen launches of an un-graphed cycle with cudaProfilerStart/Stop (AMGB_NCU_SELECT=level:op:count),
# so ncu (--profile-from-start off) sees ~20 kernels instead of ~1000.  tools/ncu_summarise.py turns the raw pages
# into profiles/r02_ncu_*.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out/ncu
O=gpurun_out/ncu
cut -c1-16 pyamg_b200/libpyamg_b200.so.sha256 > $O/so_sha16.txt
NCU="ncu --clock-control none"
MET="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,launch__shared_mem_per_block_dynamic,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__throughput.avg.pct_of_peak_sustained_elapsed"
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $1"; }
step "fine-level kernels in isolation: SpMV, residual, fused Jacobi+residual, Jacobi (metric list)"
timeout 200 $NCU --metrics $MET -k regex:csr_tile_kernel -s 4 -c 4 -o $O/fine python tools/ncu_targets.py fine > $O/fine.log 2>&1
step "fused Jacobi+residual: --set full with source"
timeout 200 $NCU --set full --import-source on -k regex:csr_tile_kernel -s 6 -c 1 -o $O/fine_jacobi_full python tools/ncu_targets.py fine > $O/fine_full.log 2>&1
step "one kernel per (level, op) of the 256^3 V-cycle (metric list)"
AMGB_NCU_SELECT="0:4:1,0:1:1,0:0:1,1:4:2,1:1:1,1:0:1,2:4:2,2:1:1,2:0:1,2:2:1,1:2:1,0:2:1,3:4:2,4:4:2,5:4:1" \
  timeout 420 $NCU --metrics $MET --profile-from-start off -o $O/cycle python tools/ncu_targets.py cycle > $O/cycle.log 2>&1
step "dominant kernel (level-1 GS wave): --set full with source"
AMGB_NCU_SELECT="1:4:1" timeout 300 $NCU --set full --import-source on --profile-from-start off -o $O/l1_gs_full python tools/ncu_targets.py cycle > $O/l1_gs_full.log 2>&1
step "block Jacobi (cfg5), Jacobi cycle (cfg2)"
AMGB_NCU_SELECT="0:5:1,1:5:1,0:1:1,0:0:1,0:2:1" timeout 120 $NCU --metrics $MET --profile-from-start off -o $O/cfg5 python tools/ncu_targets.py cfg5 > $O/cfg5.log 2>&1
AMGB_GPU_RHO=0 AMGB_NCU_SELECT="0:3:1,1:3:1,2:3:1,0:1:1,0:0:1,0:2:1" timeout 200 $NCU --metrics $MET --profile-from-start off -o $O/cfg2 python tools/ncu_targets.py cfg2 > $O/cfg2.log 2>&1
step "launch list of one bench run (headline only; first 2500 launches)"
timeout 400 $NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs "" --cpu-sample 1 > $O/launches_bench.log 2>&1
step "export"
for f in fine fine_jacobi_full cycle l1_gs_full cfg5 cfg2; do
    [ -f $O/$f.ncu-rep ] && ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.raw.csv 2>/dev/null
done
for f in fine_jacobi_full l1_gs_full; do
    [ -f $O/$f.ncu-rep ] && ncu -i $O/$f.ncu-rep --page source --csv > $O/$f.source.csv 2>/dev/null
done
rm -f $O/*.ncu-rep          # the raw / source pages carry what profiles/ needs; reports can exceed the 64 MiB limit
du -sh $O; ls -la $O
tail -2 $O/*.log
step "done"
