#!/usr/bin/env bash
# The whole on-hardware validation in the order this repo was validated (outputs -> gpurun_out/; copy what you keep to profiles/):
#   tools/b200/run_validation_suite.sh [max_gpus]
#   1 GPU : GPU test tier, smoke, both bench arms, the single-GPU predicted-vs-measured rows, GEMM vs cuBLAS column
#   N GPUs: multi-GPU test tier (collectives, ZeRO-1, fused kernels, TP / PP / CP / MoE / DeepSeek parity), bench at N (gates inside),
#           the predicted-vs-measured rows that fit N
set -uo pipefail
cd "$(dirname "$0")/../.."
MAXG="${1:-$(nvidia-smi -L | wc -l)}"
OUT=gpurun_out
mkdir -p "$OUT"
trun() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$2" --master-addr 127.0.0.1 --master-port "$3" "${@:4}"; }
echo "== 1 GPU: tests, smoke, bench (both arms)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 2>/dev/null | grep '^{' > "$OUT/bench_n1.json"
timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 --impl reference 2>/dev/null | grep '^{' > "$OUT/bench_ref_n1.json"
echo "== 1 GPU: predicted vs measured rows, cuBLAS column"
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/b200/perf_vs_real.py --out "$OUT/perf_vs_real_1gpu" 2>&1 | grep "real .* ms\|wrote"
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m simumax_b200.calib.calibrate ops --limit-per-op 50 --library-column --out "$OUT/calib_ops_library_column.json" 2>&1 | grep "vs cuBLAS"
for n in 2 4 8; do
  [ "$n" -le "$MAXG" ] || continue
  echo "== $n GPUs: bench (correctness gates, collectives table, BASELINE rows), predicted vs measured rows"
  trun 900 "$n" $((29700 + n)) bench.py --gpus "$n" --steps 3 --warmup 3 2>/dev/null | grep '^{' > "$OUT/bench_n$n.json"
  trun 900 "$n" $((29710 + n)) bench.py --gpus "$n" --steps 3 --warmup 3 --impl reference 2>/dev/null | grep '^{' > "$OUT/bench_ref_n$n.json"
  trun 1500 "$n" $((29720 + n)) tools/b200/perf_vs_real.py --out "$OUT/perf_vs_real_${n}gpu" 2>&1 | grep "real .* ms\|wrote"
done
if [ "$MAXG" -ge 2 ]; then
  echo "== multi-GPU test tier"
  timeout 900 python -m pytest tests/test_collectives_gpu.py -m gpu -q 2>&1 | tail -3
fi
