set -uo pipefail
echo "== elementwise tests"; timeout 200 python -m pytest tests/test_elementwise_gpu.py tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -3
echo "== fwd2 trace"; timeout 100 python tools/fmha_fwd_trace.py 2>&1 | tail -12
echo "== fwd2 trace MHA S8192"; timeout 100 python tools/fmha_fwd_trace.py 1 8192 16 16 2>&1 | tail -12
echo "== wgrad rmw"; timeout 100 python tools/gemm_wgrad_bench.py 2>&1 | tail -6
echo "== wgrad red"; SMX_GEMM_RED=1 timeout 100 python tools/gemm_wgrad_bench.py 2>&1 | tail -6
echo "== bandwidth ops"; timeout 100 python - <<'PY'
import torch, json
from simumax_b200.calib.measure import measure_bandwidth_ops, measure_launch_latency_us, _time, _rand
from simumax_b200.ops import elementwise as E
r = measure_bandwidth_ops()
print(json.dumps({k: {kk: round(vv, 1) for kk, vv in v.items()} for k, v in r.items()}))
print("launch latency us (graph)", measure_launch_latency_us())
T, H = 16384, 4096
xs = [_rand((T, H)) for _ in range(4)]; ds = [_rand((T, H)) for _ in range(4)]; outs = [torch.empty(T, H, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
g = torch.ones(H, device="cuda", dtype=torch.bfloat16); rstd = torch.rand(T, device="cuda") + 0.5; dg = torch.zeros(H, device="cuda")
ms = _time(lambda i: E.rmsnorm_bwd(xs[i % 4], g, ds[i % 4], rstd, dg, dx=outs[i % 4]))
print("rmsnorm_bwd GB/s (x + dy read, dx write)", 3 * T * H * 2 / ms / 1e6, "ms", ms)
ms = _time(lambda i: E.rmsnorm_bwd(xs[i % 4], g, ds[i % 4], rstd, dg, dx=outs[i % 4], accumulate_dx=True))
print("rmsnorm_bwd accumulate GB/s (x + dy + dx read, dx write)", 4 * T * H * 2 / ms / 1e6, "ms", ms)
PY
