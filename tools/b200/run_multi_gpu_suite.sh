#!/usr/bin/env bash
# Everything that needs the whole node, in one go (so an 8-GPU box is held for minutes, not hours):
#   collective sweeps at 2/4/8 GPUs (in-kernel + NCCL baseline + burst + p2p overlap), ZeRO-1 parity check, flagship bench.
# Usage: tools/b200/run_multi_gpu_suite.sh [max_gpus]      outputs -> gpurun_out/
set -uo pipefail
cd "$(dirname "$0")/../.."
MAXG="${1:-$(nvidia-smi -L | wc -l)}"
OUT=gpurun_out
mkdir -p "$OUT"
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
for n in 2 4 8; do
  [ "$n" -le "$MAXG" ] || continue
  echo "== collectives n=$n"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29600 + n)) \
    -m simumax_b200.calib.collectives --out "$OUT/coll_sweep_$n.json" 2>&1 | grep "^\[fit\|^\[coll.*262144 KiB\|Error\|error\|timeout" | tail -12
done
echo "== ZeRO-1 parity (smx vs nccl) at $MAXG GPUs"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$MAXG" --master-addr 127.0.0.1 --master-port 29621 tools/dp_check.py 2>&1 | grep "^{" | cut -c1-600
echo "== bench at $MAXG GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$MAXG" --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus "$MAXG" --steps 4 --warmup 3 2>&1 | tail -1 | tee "$OUT/bench_${MAXG}gpu.json" | cut -c1-1200
