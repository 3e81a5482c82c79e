#!/usr/bin/env bash
# convT (k2 s2) kernels at the two HBM-bound UNet levels: tile sweep + ncu --set full (run under gpurun, one GPU)
set -uo pipefail
out=gpurun_out; mkdir -p $out
P="python tools/conv_probe.py --concat --iters 20"
for mode in fwd dgrad wgrad; do
  for tn in 0 64 128 256; do
    echo "== mode=$mode tile_n=$tn"; $P --only 'convT\[(128->64|256->128)@' --convt-mode $mode --tile-n $tn 2>&1 | tail -2
  done
done
for mode in fwd dgrad wgrad; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:'conv2_tc|wgrad_tc|conv_tc' -s 5 -c 2 -f -o $out/prof_convt_$mode \
     $P --only 'convT\[128->64@' --convt-mode $mode > $out/ncu_convt_$mode.log 2>&1
done
ls -la $out/prof_convt_*.ncu-rep
