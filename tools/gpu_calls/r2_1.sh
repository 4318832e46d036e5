set -uo pipefail
OUT=gpurun_out
nvidia-smi -L
echo "== sustained calibration"
timeout 420 python -m simumax_b200.calib.calibrate ops --sustained-ms 200 --out $OUT/calib_ops_sustained.json > $OUT/calib_ops_sustained.log 2>&1; echo rc=$?; tail -3 $OUT/calib_ops_sustained.log
echo "== P in TMEM fwd test"
SMX_FMHA_P_TMEM=1 timeout 200 python -m pytest tests/test_attention_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 100 python tools/fmha_bench.py > $OUT/fmha_fwd_smem_p.log 2>&1; SMX_FMHA_P_TMEM=1 timeout 100 python tools/fmha_bench.py > $OUT/fmha_fwd_tmem_p.log 2>&1
tail -4 $OUT/fmha_fwd_smem_p.log $OUT/fmha_fwd_tmem_p.log
echo "== launch list of a 4-layer step"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_l4.csv python bench.py --layers 4 --steps 1 --warmup 1 --skip-e2e > $OUT/launches_l4.log 2>&1; echo rc=$?
echo "== memcheck elementwise"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 66 --print-limit 5 python -m pytest tests/test_elementwise_gpu.py -q -m gpu -x > $OUT/sanitize_memcheck_elementwise.log 2>&1; echo rc=$?; tail -3 $OUT/sanitize_memcheck_elementwise.log
