# 2 GPUs: every multi-GPU test of the suite on the final tree
set -uo pipefail
mkdir -p gpurun_out
timeout 560 python -m pytest tests/test_collectives_gpu.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/final_pytest_2gpu.log
