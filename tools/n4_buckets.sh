#!/usr/bin/env bash
# 4-GPU exposed-communication vs gradient bucket size
run() { tag=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 60 --warmup 3 --no-extras "$@" > gpurun_out/n4_$tag.json 2> gpurun_out/n4_$tag.err; echo "$tag rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/n4_$tag.json'));print('$tag',round(b['ms_per_step'],3),round(b['value'],1),b['extra'].get('exposed_comm_ms'),b['extra']['dp_check']['ok'],b['clocks']['sm_mhz'])"; }
mkdir -p gpurun_out
run b32
run b64 --bucket-mb 64
run b200 --bucket-mb 200
run b8 --bucket-mb 8
