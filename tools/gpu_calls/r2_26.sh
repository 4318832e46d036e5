set -uo pipefail
mkdir -p gpurun_out
timeout 100 python bench.py --gpus 1 --steps 3 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/final_bench_n1_v3.json; python -c "import json; j=json.loads(open('gpurun_out/final_bench_n1_v3.json').read().splitlines()[-1]); print('own', j['ms_per_step'], j['value'], j['e2e']['value'], j['time_error_pct'], j['peak_mem_error_pct'], j['gpu_launches'], j['bench_wall_s'], j['clocks']['reasons'])"
timeout 80 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/final_bench_ref_n1_v3.json; python -c "import json; j=json.loads(open('gpurun_out/final_bench_ref_n1_v3.json').read().splitlines()[-1]); print('ref', j['ms_per_step'], j['value'], j['e2e']['value'], j['bench_wall_s'])"
