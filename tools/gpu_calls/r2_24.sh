set -uo pipefail
mkdir -p gpurun_out
echo "== bandwidth ops (sustained)"; SMX_CALIB_SUSTAINED_MS=100 timeout 120 python - <<'PY'
import json
from simumax_b200.calib.measure import measure_bandwidth_ops
r = measure_bandwidth_ops()
json.dump(r, open("gpurun_out/bandwidth_ops_final.json", "w"), indent=1)
print(json.dumps({k: round(v["gbs"], 1) for k, v in r.items()}))
PY
echo "== bench own arm (primary)"; timeout 200 python bench.py --gpus 1 --steps 3 --warmup 3 --cases primary 2>/dev/null | grep '^{' > gpurun_out/final_bench_n1_v2.json
python -c "import json; j=json.loads(open('gpurun_out/final_bench_n1_v2.json').read().splitlines()[-1]); print(j['ms_per_step'], j['value'], j['e2e']['value'], j['time_error_pct'], j['peak_mem_error_pct'], j['predicted']['iter_ms'], j['clocks'])"
