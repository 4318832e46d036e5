#!/usr/bin/env bash
# Calibrate THIS machine end to end and write simumax_b200/configs/system/<name>.json.
#   tools/b200/build_current_machine_system_config.sh [name] [max_gpus]
#   env: SWEEP=default (built-in zoo: ~800 GEMM / 35 + 35 attention / 96 grouped / 60 fp8 / 60 MXFP8 keys; unset = the example workloads only),
#        PARAM_FILE=<sweep.json>, SUSTAINED_MS=100 (time every op at the power-capped steady clock; 0 = burst), CACHE_MODE=rebuild|supplement,
#        CACHE_TAG=<name>, LIBRARY_COLUMN=1 (cuBLAS on the same operands next to every bf16 GEMM key)
# Steps: (1) single-GPU operator sweeps with the sm_100a kernels, (2) collective sweeps + burst-window probe at
# 2/4/8 GPUs (in-kernel NVLink collectives and the NCCL baseline), (3) merge into a system config, (4) replay the
# example workloads with the new config and print hit/miss statistics.
# (reference: tools/b200/build_current_machine_system_config.py)
set -euo pipefail
cd "$(dirname "$0")/../.."
NAME="${1:-b200_sm100a_calibrated}"
MAXG="${2:-$(nvidia-smi -L | wc -l)}"
OUT=gpurun_out
mkdir -p "$OUT"
python -c "from simumax_b200.ops import _lib; _lib.build()"
ARGS=(--sustained-ms "${SUSTAINED_MS:-100}" --cache-mode "${CACHE_MODE:-rebuild}")
[ "${SWEEP:-}" = "default" ] && ARGS+=(--default-sweep)
[ -n "${PARAM_FILE:-}" ] && ARGS+=(--param-file "$PARAM_FILE")
[ -n "${CACHE_TAG:-}" ] && ARGS+=(--cache-tag "$CACHE_TAG")
[ "${LIBRARY_COLUMN:-0}" = "1" ] && ARGS+=(--library-column)
python -m simumax_b200.calib.calibrate ops --out "$OUT/calib_ops.json" "${ARGS[@]}" | tee "$OUT/calib_ops.log"
OPS="$OUT/calib_ops${CACHE_TAG:+_$CACHE_TAG}.json"
COLL=()
for n in 2 4 8; do
  if [ "$n" -le "$MAXG" ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      -m simumax_b200.calib.collectives --out "$OUT/coll_sweep_$n.json" | tee "$OUT/coll_sweep_$n.log"
    COLL+=("$OUT/coll_sweep_$n.json")
  fi
done
python -m simumax_b200.calib.calibrate merge --ops "$OPS" --coll "${COLL[@]}" \
  --out "simumax_b200/configs/system/$NAME.json" --name "$NAME"
SIMU_SYSTEM="$NAME" python tools/b200/check_calibrated_config.py "$NAME"
