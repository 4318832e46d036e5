set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
echo "== gpu tests (elementwise, trainer)"
timeout 300 python -m pytest tests/test_elementwise_gpu.py tests/test_trainer_gpu.py -q -m gpu -x 2>&1 | tail -15
echo "== lm_head debug"
timeout 120 python tools/debug_lm_head.py 2>&1 | tail -12
echo "== bench own arm N=1"
timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 > $OUT/bench_r2_n1.json 2> $OUT/bench_r2_n1.err; echo rc=$?; tail -c 1500 $OUT/bench_r2_n1.err; head -c 3000 $OUT/bench_r2_n1.json
echo "== bench reference arm N=1"
timeout 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > $OUT/bench_r2_ref_n1.json 2> $OUT/bench_r2_ref_n1.err; echo rc=$?; tail -c 1500 $OUT/bench_r2_ref_n1.err; head -c 3000 $OUT/bench_r2_ref_n1.json
