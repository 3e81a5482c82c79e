#!/usr/bin/env bash
# 8-GPU exposed-communication variants (NCCL CTA bound, bucket size); prints ms/step, img/s, exposed_comm_ms per variant
run() { tag=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 60 --warmup 3 --no-extras "$@" > gpurun_out/n8_$tag.json 2> gpurun_out/n8_$tag.err; echo "$tag rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/n8_$tag.json'));print('$tag',round(b['ms_per_step'],3),round(b['value'],1),b['extra'].get('exposed_comm_ms'),b['extra']['dp_check']['ok'],b['clocks']['sm_mhz'])"; }
mkdir -p gpurun_out
run ctas8 --nccl-max-ctas 8
run ctas16 --nccl-max-ctas 16
run default
