set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:fmha_fwd2 -s 1 -c 1 -f -o $OUT/prof_fwd2 python tools/ncu_fmha.py > $OUT/ncu_fwd2.log 2>&1; echo rc=$?; tail -3 $OUT/ncu_fwd2.log
for k in grouped mxfp8 elementwise ce permute optimizer embedding; do
  timeout 150 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:smx:: -s 2 -c 12 -f -o $OUT/prof_$k python tools/ncu_misc.py $k > $OUT/ncu_$k.log 2>&1; echo "$k rc=$?"
done
ls -la $OUT/*.ncu-rep
