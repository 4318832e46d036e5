#!/usr/bin/env bash
# Calibrate THIS machine end to end and write simumax_b200/configs/system/<name>.json.
#   tools/b200/build_current_machine_system_config.sh [name] [max_gpus]
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
python -m simumax_b200.calib.calibrate ops --out "$OUT/calib_ops.json" | tee "$OUT/calib_ops.log"
COLL=()
for n in 2 4 8; do
  if [ "$n" -le "$MAXG" ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      -m simumax_b200.calib.collectives --out "$OUT/coll_sweep_$n.json" | tee "$OUT/coll_sweep_$n.log"
    COLL+=("$OUT/coll_sweep_$n.json")
  fi
done
python -m simumax_b200.calib.calibrate merge --ops "$OUT/calib_ops.json" --coll "${COLL[@]}" \
  --out "simumax_b200/configs/system/$NAME.json" --name "$NAME"
SIMU_SYSTEM="$NAME" python tools/b200/check_calibrated_config.py "$NAME"
