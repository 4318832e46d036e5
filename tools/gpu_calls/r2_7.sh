# 1 GPU: second-generation attention forward, pipelined norm kernels, split backward
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
echo "== attention + elementwise + trainer tests"
timeout 400 python -m pytest tests/test_attention_gpu.py tests/test_elementwise_gpu.py tests/test_trainer_gpu.py -q -m gpu -x 2>&1 | tail -15
echo "== fwd bench v2 (auto cluster)"; timeout 120 python tools/fmha_bench.py 2>&1 | tail -6
echo "== fwd bench v2 (no cluster)"; SMX_FMHA_CLUSTER=1 timeout 120 python tools/fmha_bench.py 2>&1 | tail -6 | cut -c1-200
echo "== fwd bench v1"; SMX_FMHA_V1=1 timeout 120 python tools/fmha_bench.py 2>&1 | tail -6 | cut -c1-200
echo "== bwd bench"; timeout 200 python tools/fmha_bwd_bench.py 2>&1 | tail -8
echo "== bandwidth ops"; timeout 100 python - <<'PY'
import torch, json
from simumax_b200.calib.measure import measure_bandwidth_ops, measure_launch_latency_us
r = measure_bandwidth_ops()
print(json.dumps({k: {kk: round(vv, 1) for kk, vv in v.items()} for k, v in r.items()}))
print("launch latency us", measure_launch_latency_us())
PY
