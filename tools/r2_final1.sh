#!/bin/bash
# Round-2 final single-GPU call: whole gpu suite + smoke on the frozen library, then the bench line the driver will run.
mkdir -p gpurun_out
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $1"; }
step "pytest -m gpu"
timeout 900 python -m pytest tests -q -m gpu -x -rs 2>&1 | tail -6
step "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
step "bench (ours)"
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2_bench_n1.err | tail -1 > gpurun_out/r2_bench_n1.json
grep "^\[bench\]" gpurun_out/r2_bench_n1.err | cut -c1-400
grep -iE "error|traceback" gpurun_out/r2_bench_n1.err | head
step "bench (reference arm, 3 steps)"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2> gpurun_out/r2_bench_ref.err | tail -1 > gpurun_out/r2_bench_ref.json
cut -c1-600 gpurun_out/r2_bench_ref.json
step "done"
