# final validation: full 1-GPU test tier, smoke, both bench arms with the final code
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/final_pytest_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -3
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 2>/dev/null | grep '^{' > $OUT/final_bench_ref_n1.json; cut -c1-400 $OUT/final_bench_ref_n1.json
echo "== bench own arm"; timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 2>/dev/null | grep '^{' > $OUT/final_bench_n1.json; cut -c1-700 $OUT/final_bench_n1.json
python - <<'PY'
import json
a=json.loads(open('gpurun_out/final_bench_n1.json').read().splitlines()[-1]); b=json.loads(open('gpurun_out/final_bench_ref_n1.json').read().splitlines()[-1])
print("own", a["value"], "e2e", a["e2e"]["value"], "ref", b["value"], "e2e", b["e2e"]["value"], "ratio e2e", a["e2e"]["value"]/b["e2e"]["value"])
print({k: a.get(k) for k in ("gpu_launches","clocks")}); print(a.get("accuracy") or a.get("prediction"))
PY
