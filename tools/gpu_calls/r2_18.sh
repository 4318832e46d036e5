set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
echo "== attention tests gen3"; SMX_FMHA_GEN=3 timeout 300 python -m pytest tests/test_attention_gpu.py -q -m gpu -x 2>&1 | tail -4
echo "== fwd bench gen3"; SMX_FMHA_GEN=3 timeout 120 python tools/fmha_bench.py 2>&1 | grep -v '"D": 192' | tail -4 | cut -c1-230
echo "== gen3 trace"; SMX_FMHA_GEN=3 timeout 100 python tools/fmha_fwd_trace.py 2>&1 | tail -10
echo "== gemm tests (red epilogue default)"; timeout 200 python -m pytest tests/test_gemm_gpu.py tests/test_trainer_gpu.py -q -m gpu -x 2>&1 | tail -3
echo "== recalibrate fp32-accumulating wgrads"
cp profiles/calib_ops_v7_without_wgrad.json $OUT/calib_ops_v8.json
timeout 500 python -m simumax_b200.calib.calibrate ops --default-sweep --sustained-ms 100 --cache-mode supplement --out $OUT/calib_ops_v8.json > $OUT/calib_v8.log 2>&1; echo rc=$?; grep -c "accumulate=True\|bwd_grad_w" $OUT/calib_v8.log; tail -2 $OUT/calib_v8.log
