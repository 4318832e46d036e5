# 1 GPU: big sustained calibration sweep, 1-GPU predicted-vs-measured zoo, ncu of the uncovered kernels, sanitizer
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
echo "== perf vs real (1 GPU zoo)"
timeout 500 python tools/b200/perf_vs_real.py --out $OUT/perf_vs_real_1gpu 2>&1 | grep -v Warn | tail -12
echo "== sustained calibration, default sweep"
timeout 900 python -m simumax_b200.calib.calibrate ops --default-sweep --sustained-ms 100 --cache-tag sweep --out $OUT/calib_ops.json > $OUT/calib_sweep.log 2>&1; echo rc=$?; tail -4 $OUT/calib_sweep.log
echo "== ncu misc"
for k in grouped mxfp8 elementwise ce permute optimizer embedding; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:'smx' -s 2 -c 14 -f -o $OUT/prof_$k python tools/ncu_misc.py $k > $OUT/ncu_$k.log 2>&1; echo "$k rc=$?"
done
echo "== sanitizer (memcheck, api errors off: the cudart lazy-load probe is not a kernel error)"
timeout 400 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 66 --print-limit 5 python -m pytest tests/test_elementwise_gpu.py tests/test_trainer_gpu.py -q -m gpu -x > $OUT/sanitize_memcheck.log 2>&1; echo rc=$?; tail -3 $OUT/sanitize_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --report-api-errors no --error-exitcode 66 --print-limit 5 python -m pytest tests/test_elementwise_gpu.py -q -m gpu -x > $OUT/sanitize_racecheck.log 2>&1; echo rc=$?; tail -3 $OUT/sanitize_racecheck.log
