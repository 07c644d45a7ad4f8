#!/bin/bash
# Round-2: A/B of the per-operator tile geometry / lanes-per-row rule / flat gathers for R and P on the 256^3 hierarchy,
# parity of the new paths, timing of the device colouring.
mkdir -p gpurun_out
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $1"; }
step "parity: every kernel path + device colouring"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_widening.py -q -m gpu -x -k "every_kernel_path or device_mis" 2>&1 | tail -4
step "A/B"
S="AMGB_NO_PDL=0"
S="$S;AMGB_TILE_LANE_ENTRIES=8"
S="$S;AMGB_TILE_DENSE_CFG=7"
S="$S;AMGB_TILE_DENSE_CFG=7,AMGB_TILE_LANE_ENTRIES=8"
S="$S;AMGB_TILE_FLAT=2"
S="$S;AMGB_TILE_FLAT=2,AMGB_TILE_DENSE_CFG=7"
S="$S;AMGB_TILE_DENSE_CFG=7,AMGB_TILE_DENSE_AVG=24"
S="$S;AMGB_TILE_LANE_ENTRIES=6"
timeout 1200 python tools/tune_tiles.py --grid 256 --settings "$S" 2>&1 | grep -E "cycle_ms|Error|error|rror" | cut -c1-1200 | tee gpurun_out/r2_ab3.jsonl
step "timing: widening rows (incl. device colouring)"
timeout 600 python tools/time_widening.py --grid 128 2>&1 | grep -E "coloring|rho|galerkin|Error" | tee gpurun_out/r2_widening2.jsonl
step "done"
