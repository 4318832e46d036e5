set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
echo "== trainer tests"
timeout 300 python -m pytest tests/test_trainer_gpu.py tests/test_elementwise_gpu.py -q -m gpu -x 2>&1 | tail -12
echo "== deepseek check (1 GPU)"
timeout 300 python tools/deepseek_check.py 2>&1 | tail -25
