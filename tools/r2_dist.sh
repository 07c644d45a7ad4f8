#!/bin/bash
# Round-2 multi-GPU call (gpurun --gpus N): parity of the partitioned cycle (peer-memory halo, graphed), then bench.
N=${1:-2}
CONFIGS=${2:-}
mkdir -p gpurun_out
T0=$(date +%s)
step() { echo "=== [$(( $(date +%s) - T0 )) s] $1"; }
if [ "${3:-tests}" != "notest" ]; then
step "pytest tests/test_dist_gpu.py (2 ranks: peer, peer + graph, allgather)"
timeout 900 python -m pytest tests/test_dist_gpu.py -q -m gpu -x -s 2>&1 | grep -E "dist-gpu|passed|failed|Error|error" | tail -30
fi
step "bench --gpus $N (configs: '$CONFIGS')"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N --steps 10 --warmup 3 --configs "$CONFIGS" 2> gpurun_out/r2_dist_n$N.err | tail -1 > gpurun_out/r2_dist_n$N.json
grep -E "^\[bench\]" gpurun_out/r2_dist_n$N.err | tail -20
grep -iE "error|traceback|trap|illegal" gpurun_out/r2_dist_n$N.err | head -10
cut -c1-1500 gpurun_out/r2_dist_n$N.json
step "done"
