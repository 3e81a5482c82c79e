#!/usr/bin/env bash
# Runs the GPU parity tests group by group, each in its own process (a trapped kernel poisons its CUDA context),
# and leaves the logs under gpurun_out/diag/. Usage: tools/gpu_diag.sh [group ...]
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/diag
mkdir -p "$out"
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,memory.total --format=csv > "$out/gpu.csv" 2>&1
groups=("$@")
if [ ${#groups[@]} -eq 0 ]; then
  groups=(conv1x1_gemm conv3x3_fwd channel_slices dgrad conv3x3_wgrad transpose large pack first_conv bn_finalize bn_relu head seg_loss "adamw or copy or argument")
fi
for g in "${groups[@]}"; do
  name=$(echo "$g" | tr ' ' '_')
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -k "$g" -p no:cacheprovider > "$out/$name.log" 2>&1
  echo "== $g: exit $? :: $(tail -n 1 "$out/$name.log")"
done
