#!/usr/bin/env bash
# ncu captures of one round (run under gpurun, one GPU): launch list of two host-launched training steps and
# --set full captures of the bandwidth kernels. Usage: tools/ncu_round.sh <tag>
# Per-launch times under ncu are cold-cache and serialised; bench numbers are never taken from these runs.
set -uo pipefail
tag="${1:-rX}"
out=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph"
$B > $out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 $out/plain_$tag.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 420 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_launches_$tag.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'c1_|head_|seg_loss|adamw|pack_all' -s 27 -c 9 \
    -f -o $out/prof_${tag}_ew $B > $out/ncu_ew_$tag.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'bn_bwd_kernel|bn_apply_kernel' -s 228 -c 10 \
    -f -o $out/prof_${tag}_bn $B > $out/ncu_bn_$tag.log 2>&1
ls -la $out/prof_${tag}_*.ncu-rep $out/launches_$tag.csv
# dominant tensor-core kernels: one launch each of the row-halo (64 -> 64 @256^2) and tile-pair (512 -> 512 @32^2) convs
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv2_tc -s 4 -c 1 -f -o $out/prof_${tag}_halo64 \
    python tools/conv_probe.py --iters 4 --only "fwd\[64->64@256" > $out/ncu_halo64_$tag.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv2_tc -s 4 -c 1 -f -o $out/prof_${tag}_pair256 \
    python tools/conv_probe.py --iters 4 --only "fwd\[512->512@32" > $out/ncu_pair256_$tag.log 2>&1
ls -la $out/prof_${tag}_*.ncu-rep
