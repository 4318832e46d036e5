#!/usr/bin/env bash
# Code written after the round-1 GPU budget ran out: everything here is verified on CPU as far as that goes (schedule logic under
# gloo, syntax, nvcc build) but has never executed on a GPU.  Run this first in the next session, on a node with >= 4 GPUs:
#   tools/b200/run_pending_gpu_checks.sh [max_gpus]            outputs -> gpurun_out/
set -uo pipefail
cd "$(dirname "$0")/../.."
MAXG="${1:-$(nvidia-smi -L | wc -l)}"
OUT=gpurun_out
mkdir -p "$OUT"
run() { timeout "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$2" --master-addr 127.0.0.1 --master-port "$3" "${@:4}" 2>&1 \
          | grep -v "^W0\|^\*\*\*\|OMP_NUM\|FutureWarning\|enable_symm_mem\|Recompute is\|warnings.warn"; }
echo "== 1. pipeline-parallel trainer (1F1B, NCCL p2p on a side stream) vs single GPU"
for n in 2 4; do [ "$n" -le "$MAXG" ] && run 200 "$n" $((29800 + n)) tools/pp_check.py | tail -2 | cut -c1-1200; done
echo "== 2. TP x DP (ZeRO-1 over replicas): the reference's tp2_dp4 / tp4_dp2 rows"
if [ "$MAXG" -ge 8 ]; then
  for tp in 2 4; do
    run 300 8 $((29810 + tp)) tools/b200/perf_vs_real.py --tp --tp-size "$tp" --cases llama3_70b_l12_s4096_tp \
        --out "$OUT/perf_vs_real_tp${tp}_dp$((8 / tp)).json" | tail -2
  done
fi
echo "== 2b. the other tp8_dp1 rows of the reference's published table (70B-l12 mbc8 / mbc32, 405B-l4): measured vs Megatron's published ms"
[ "$MAXG" -ge 8 ] && run 400 8 29820 tools/b200/perf_vs_real.py --tp --out "$OUT/perf_vs_real_tp8_published.json" \
    --cases llama3_70b_l12_s4096_tp llama3_70b_l12_s4096_mbc8_tp llama3_405b_l4_s4096_tp llama3_405b_l4_s4096_mbc8_tp | tail -5
echo "== 3. sustained-clock calibration of the bench shapes (1 GPU), then predicted vs measured with it"
timeout 600 python -m simumax_b200.calib.calibrate ops --sustained-ms 300 --limit-per-op 40 --out "$OUT/calib_ops_sustained.json" | tail -5
python -m simumax_b200.calib.calibrate merge --ops "$OUT/calib_ops_sustained.json" --base simumax_b200/configs/system/b200_sm100a_calibrated.json \
    --out simumax_b200/configs/system/b200_sm100a_sustained.json --name b200_sm100a_sustained | tail -1
timeout 300 python tools/b200/perf_vs_real.py --system b200_sm100a_sustained --cases llama3_8b_full_s4096 --out "$OUT/perf_vs_real_sustained.json" | tail -2
echo "== 4. ncu capture of the fused GEMM -> reduce-scatter epilogue + consumer (1 GPU, world = 1)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemm_2cta|rs_consume' -c 6 -f -o "$OUT/prof_fused" \
    python tools/ncu_fused.py > "$OUT/ncu_fused.log" 2>&1 && python tools/ncu_summary.py "$OUT/prof_fused.ncu-rep" profiles/gemm_rs_fused_ncu_summary.json
echo "== 4b. compute-sanitizer memcheck over the elementwise kernels (the GEMM / attention passes take ~10 min each: tools/sanitize_kernels.sh)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 66 --print-limit 5 python -m pytest tests/test_elementwise_gpu.py -q -m gpu -x \
    > "$OUT/sanitize_memcheck_elementwise.log" 2>&1; echo "   rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' "$OUT/sanitize_memcheck_elementwise.log") clean"
echo "== 4c. experimental attention forward with P in TMEM (tcgen05.st + TMEM-A MMA): correctness, then A/B timing"
SMX_FMHA_P_TMEM=1 timeout 300 python -m pytest tests/test_attention_gpu.py -q -m gpu -x 2>&1 | tail -2
timeout 200 python tools/fmha_bench.py > "$OUT/fmha_fwd_smem_p.log" 2>&1; SMX_FMHA_P_TMEM=1 timeout 200 python tools/fmha_bench.py > "$OUT/fmha_fwd_tmem_p.log" 2>&1
tail -3 "$OUT/fmha_fwd_smem_p.log" "$OUT/fmha_fwd_tmem_p.log"
echo "== 5. opt-in GPU test"
SMX_TEST_PP=1 timeout 300 python -m pytest tests/test_collectives_gpu.py -q -m gpu -k pipeline 2>&1 | tail -2
