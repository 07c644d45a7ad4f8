#!/bin/bash
# Round-2, GPU call 1: parity of everything (reference importable from baseline/_ref), A/B of the opt-in kernel
# variants on the 256^3 hierarchy, first numbers for BASELINE configs[1] and [4].
mkdir -p gpurun_out
echo "=== pytest -m gpu (whole suite; the reference is importable: no from_pyamg skips expected)"
timeout 900 python -m pytest tests -q -m gpu -x -rs 2>&1 | tail -15
echo "=== parity: experimental paths"
AMGB_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -q -m gpu_experimental 2>&1 | tail -15
echo "=== timing: widening rows"
timeout 900 python tools/time_widening.py --grid 128 2>&1 | tail -30 | tee gpurun_out/r2_widening.jsonl
echo "=== A/B (graphed cycle ms, small_levels ms, per-level GB/s)"
S="AMGB_NO_PDL=0"
S="$S;AMGB_TILE_PDL=1"
S="$S;AMGB_RESIDENT=1"
S="$S;AMGB_RESIDENT=1,AMGB_RESIDENT_MAX_ROWS=200000"
S="$S;AMGB_RESIDENT=1,AMGB_RESIDENT_MAX_ROWS=5000"
S="$S;AMGB_TAIL_NNZ=300000"
S="$S;AMGB_TILE_FLAT=1"
S="$S;AMGB_TILE_FLAT=1,AMGB_TILE_MIN_NNZ=300000"
S="$S;AMGB_TILE_FLAT=1,AMGB_TILE_PDL=1,AMGB_RESIDENT=1"
S="$S;AMGB_TILE_FLAT=1,AMGB_TILE_CTAS=5"
S="$S;AMGB_TILE_FLAT=1,AMGB_TILE_CTAS=7"
timeout 1500 python tools/tune_tiles.py --grid 256 --settings "$S" 2>&1 | grep -E "cycle_ms|Error|error" | cut -c1-700 | tee gpurun_out/r2_ab.jsonl
echo "=== bench cfg2 / cfg5"
timeout 600 python bench.py --workload cfg2 --grid 2000 --steps 20 --warmup 3 2>gpurun_out/r2_cfg2.err | tail -1 | tee gpurun_out/r2_bench_cfg2.json | cut -c1-3000
timeout 600 python bench.py --workload cfg5 --grid 300 --steps 20 --warmup 3 2>gpurun_out/r2_cfg5.err | tail -1 | tee gpurun_out/r2_bench_cfg5.json | cut -c1-3000
tail -5 gpurun_out/r2_cfg2.err gpurun_out/r2_cfg5.err
