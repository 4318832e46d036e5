# after restoring the RED wgrad epilogue as the default: DeepSeek 1-GPU row on a cool GPU (A/B of the epilogue), final bench of both
# arms, final 1-GPU predicted-vs-measured rows with calibration v10
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
for red in 1 0; do
  echo "== deepseek 1-GPU row, SMX_GEMM_RED=$red"
  SMX_GEMM_RED=$red timeout 200 python tools/b200/perf_vs_real.py --cases deepseekv2_l2_ep1_pp1_dp1_mbc2 --out $OUT/pvr_ds_red$red 2>&1 | grep "real .* ms" | tail -2
done
echo "== gemm + trainer tests"; timeout 200 python -m pytest tests/test_gemm_gpu.py tests/test_trainer_gpu.py -q -m gpu -x 2>&1 | tail -2
echo "== bench own arm"; timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 2>/dev/null | grep '^{' > $OUT/final_bench_n1.json; cut -c1-260 $OUT/final_bench_n1.json
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 2>/dev/null | grep '^{' > $OUT/final_bench_ref_n1.json; cut -c1-260 $OUT/final_bench_ref_n1.json
python - <<'PY'
import json
a=json.loads(open('gpurun_out/final_bench_n1.json').read().splitlines()[-1]); b=json.loads(open('gpurun_out/final_bench_ref_n1.json').read().splitlines()[-1])
print("own", a["value"], "e2e", a["e2e"]["value"], "ref", b["value"], "e2e", b["e2e"]["value"], "ratio e2e", a["e2e"]["value"]/b["e2e"]["value"])
print({k: a.get(k) for k in ("gpu_launches","clocks","ms_per_step","time_error_pct","peak_mem_error_pct")})
PY
echo "== perf vs real, 1-GPU rows"; timeout 500 python tools/b200/perf_vs_real.py --out $OUT/perf_vs_real_final_1gpu 2>&1 | grep "real .* ms\|wrote" | tail -10
