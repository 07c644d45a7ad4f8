#!/bin/bash
# Round-2 ncu evidence for the kernels AS SHIPPED.  One gpurun call; the raw pages are
# exported to CSV on the box (the .ncu-rep files of ~100 full captures would exceed the 64 MiB gpurun_out limit),
# tools/ncu_summarise.py turns them into profiles/r02_ncu_*.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out/ncu
O=gpurun_out/ncu
SHA=$(cut -c1-16 pyamg_b200/libpyamg_b200.so.sha256)
echo "$SHA" > $O/so_sha16.txt
NCU="ncu --clock-control none"
M="--set full"
echo "=== launch list of one bench run (headline only)"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --configs "" --cpu-sample 1 > $O/launches_bench.log 2>&1
echo "=== full: fine-level kernels in isolation (SpMV, residual, fused Jacobi+residual, Jacobi)"
timeout 600 $NCU $M --import-source on -k regex:csr_tile_kernel -s 4 -c 4 -o $O/fine python tools/ncu_targets.py fine > $O/fine.log 2>&1
echo "=== full: tile kernels of one V-cycle of the 256^3 hierarchy (down-leg levels 0-2, first up-leg launches)"
timeout 1200 $NCU $M -k regex:csr_tile_kernel -c 96 -o $O/cycle_tiles python tools/ncu_targets.py cycle > $O/cycle_tiles.log 2>&1
echo "=== full: lanes-per-row waves of the small levels (level 3/4), dense coarse solve"
timeout 900 $NCU $M -k regex:csr_rows_kernel -s 120 -c 6 -o $O/cycle_rows python tools/ncu_targets.py cycle > $O/cycle_rows.log 2>&1
timeout 900 $NCU $M -k regex:dense_matvec_kernel -c 1 -o $O/cycle_dense python tools/ncu_targets.py cycle > $O/cycle_dense.log 2>&1
echo "=== full: block Jacobi (cfg5), Jacobi cycle (cfg2)"
timeout 600 $NCU $M -k regex:block_jacobi_kernel -c 4 -o $O/cfg5_bj python tools/ncu_targets.py cfg5 > $O/cfg5.log 2>&1
timeout 600 $NCU $M -k regex:csr_tile_kernel -c 12 -o $O/cfg2_tiles python tools/ncu_targets.py cfg2 > $O/cfg2.log 2>&1
for f in fine cycle_tiles cycle_rows cycle_dense cfg5_bj cfg2_tiles; do
    [ -f $O/$f.ncu-rep ] && ncu -i $O/$f.ncu-rep --page raw --csv > $O/$f.raw.csv 2>/dev/null
done
# keep only the small reports (source-level pages of the fine-level kernels); the others stay as CSV
rm -f $O/cycle_tiles.ncu-rep $O/cycle_rows.ncu-rep $O/cfg2_tiles.ncu-rep
ls -la $O
tail -3 $O/*.log
