# pair-vs-single kernel choice A/B on few-tile shapes (sustained clock, cuBLAS next to it); ncu of the two wgrad epilogues
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
echo "== gemm tests with the new heuristic"; timeout 200 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -2
echo "== pair fill A/B"; timeout 200 python tools/gemm_pair_fill_ab.py 2>&1 | tail -18
for red in 0 1; do
  echo "== ncu wgrad epilogue RED=$red"
  SMX_GEMM_RED=$red timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:smx::gemm -s 2 -c 1 -f -o $OUT/prof_wgrad_red$red python tools/ncu_misc.py wgrad > $OUT/ncu_wgrad_red$red.log 2>&1; echo rc=$?
done
ls -la $OUT/prof_wgrad_red*.ncu-rep
