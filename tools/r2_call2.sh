#!/bin/bash
# Round-2, GPU call 2: the persistent grid kernel for the coarse part (grid_kernel.cuh) -- parity, then A/B on 256^3.
mkdir -p gpurun_out
echo "=== parity: every kernel path (incl. AMGB_TAIL_GRID and forced tile paths)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "every_kernel_path" 2>&1 | tail -5
echo "=== A/B"
S="AMGB_NO_PDL=0"
S="$S;AMGB_TAIL_GRID=1"
S="$S;AMGB_TAIL_GRID=1,AMGB_TAIL_NNZ=6000000"
S="$S;AMGB_TAIL_GRID=1,AMGB_TAIL_NNZ=60000000"
S="$S;AMGB_TAIL_GRID=1,AMGB_TAIL_SOLO_BYTES=100000"
S="$S;AMGB_TAIL_GRID=1,AMGB_TAIL_SOLO_BYTES=1600000"
S="$S;AMGB_TAIL_GRID=1,AMGB_TAIL_NNZ=6000000,AMGB_TAIL_SOLO_BYTES=1600000"
S="$S;AMGB_TAIL_GRID=1,AMGB_GRID_CTAS=74"
S="$S;AMGB_TAIL_GRID=1,AMGB_TAIL_NNZ=1000000"
timeout 1500 python tools/tune_tiles.py --grid 256 --settings "$S" 2>&1 | grep -E "cycle_ms|Error|error|rror" | cut -c1-700 | tee gpurun_out/r2_ab2.jsonl
echo "=== timing: widening rows"
timeout 900 python tools/time_widening.py --grid 128 2>&1 | tail -30 | tee gpurun_out/r2_widening.jsonl
