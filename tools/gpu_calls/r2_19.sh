# final validation: full 1-GPU test tier, smoke, both bench arms with the final code, 1-GPU predicted-vs-measured rows with the
# final calibration, wgrad epilogue A/B on the whole step, cuBLAS column on a slice of the sweep
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/final_pytest_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 2>/dev/null | grep '^{' > $OUT/final_bench_ref_n1.json; cut -c1-300 $OUT/final_bench_ref_n1.json
echo "== bench own arm"; timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 2>/dev/null | grep '^{' > $OUT/final_bench_n1.json; cut -c1-500 $OUT/final_bench_n1.json
for red in 0 1; do
  echo "== bench own arm, SMX_GEMM_RED=$red (primary case only)"
  SMX_GEMM_RED=$red timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --cases primary --skip-checks 2>/dev/null | grep '^{' > $OUT/final_bench_n1_red$red.json
  python -c "import json; j=json.loads(open('$OUT/final_bench_n1_red$red.json').read().splitlines()[-1]); print(j['ms_per_step'], j['value'], j['clocks'])"
done
python - <<'PY'
import json
a=json.loads(open('gpurun_out/final_bench_n1.json').read().splitlines()[-1]); b=json.loads(open('gpurun_out/final_bench_ref_n1.json').read().splitlines()[-1])
print("own", a["value"], "e2e", a["e2e"]["value"], "ref", b["value"], "e2e", b["e2e"]["value"], "ratio e2e", a["e2e"]["value"]/b["e2e"]["value"])
print({k: a.get(k) for k in ("gpu_launches","clocks","ms_per_step")})
print({k: v for k, v in a.items() if "pred" in k or "err" in k})
PY
echo "== perf vs real, 1-GPU rows"; timeout 500 python tools/b200/perf_vs_real.py --out $OUT/perf_vs_real_final_1gpu 2>&1 | tail -14
echo "== cuBLAS column (burst clocks, first 50 keys per op)"; timeout 200 python -m simumax_b200.calib.calibrate ops --limit-per-op 50 --library-column --out $OUT/calib_ops_library_column.json 2>&1 | grep "vs cuBLAS\|Traceback\|Error" | head -5
