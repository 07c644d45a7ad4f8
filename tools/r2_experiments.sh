#!/bin/bash
# One gpurun call that validates and A/B-times every opt-in path left unmeasured at the end of round 1.
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_experiments.sh 256 > gpurun_out/r2_exp.log 2>&1'
#
# 1. parity of the default suite's newest test and of the experimental paths (fails loudly, keeps going);
# 2. graphed V-cycle time + per-(level, op) GB/s of the 256^3 hierarchy under each setting -- the hierarchy is
#    built once (tools/tune_tiles.py re-uploads per setting);
# 3. a fresh launch list (ncu, bounded) for the default setting.
G=${1:-256}
mkdir -p gpurun_out
echo "=== parity: newest default-suite test"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "out_buffer" 2>&1 | tail -3
echo "=== parity: SURVEY 8(f) widening (smoothers, AMLI, GMRES/FGMRES, GPU Galerkin) -- first hardware run"
timeout 1500 python -m pytest tests/test_zz_gpu_widening.py -q -m gpu 2>&1 | tail -8
echo "=== timing: widening rows"
timeout 1500 python tools/time_widening.py --grid 128 2>&1 | tail -20 | tee gpurun_out/r2_widening.jsonl
echo "=== parity: experimental paths"
AMGB_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_experimental.py -q -m gpu_experimental 2>&1 | tail -15
echo "=== A/B (graphed cycle ms, small_levels ms, per-level GB/s)"
S="AMGB_NO_PDL=0"                                              # baseline (default settings)
S="$S;AMGB_TILE_PDL=1"                                         # tile kernels programmatically launched
S="$S;AMGB_RESIDENT=1"                                         # levels <= 65536 rows: one cluster launch per smoother
S="$S;AMGB_RESIDENT=1,AMGB_TILE_PDL=1"
S="$S;AMGB_RESIDENT=1,AMGB_RESIDENT_MAX_ROWS=200000"           # + the 176 k-row level
S="$S;AMGB_RESIDENT=1,AMGB_RESIDENT_MAX_ROWS=5000"             # only the smallest levels
S="$S;AMGB_RESIDENT=1,AMGB_TILE_PDL=1,AMGB_TILE_MIN_NNZ=400000"
S="$S;AMGB_TILE_FLAT=1"                                        # flat-gather tile kernel on every tiled operator
S="$S;AMGB_TILE_FLAT=1,AMGB_TILE_MIN_NNZ=300000"               # ... also on the 176 k-row level's waves
S="$S;AMGB_TILE_FLAT=1,AMGB_TILE_PDL=1"
S="$S;AMGB_TILE_FLAT=1,AMGB_TILE_PDL=1,AMGB_RESIDENT=1"
S="$S;AMGB_TILE_FLAT=1,AMGB_TILE_CTAS=5"
S="$S;AMGB_TILE_FLAT=1,AMGB_TILE_CTAS=7"
timeout 1500 python tools/tune_tiles.py --grid $G --settings "$S" 2>&1 | grep -E "cycle_ms|Error|error" | cut -c1-600
echo "=== bench line (default settings)"
timeout 900 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/r2_bench_default.json | cut -c1-400
echo "=== ncu: launch list of the widened rows + one full capture of the Galerkin SpGEMM and of a block-GS wave"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_widening_launches.csv \
    python tools/time_widening.py --grid 48 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spgemm_row_kernel -c 3 \
    -o gpurun_out/r2_spgemm python tools/time_widening.py --grid 64 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:schwarz_kernel|kaczmarz_kernel" -c 4 \
    -o gpurun_out/r2_conflict_waves python -m pytest tests/test_zz_gpu_widening.py -q -m gpu -k "vcycle_matches_reference_golden and (cfg13 or cfg15)" > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:block_gs_kernel -c 2 \
    -o gpurun_out/r2_block_gs python -m pytest tests/test_zz_gpu_widening.py -q -m gpu -k "vcycle_matches_reference_golden and cfg10" > /dev/null 2>&1

