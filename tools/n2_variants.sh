run() { tag=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 60 --warmup 3 --no-extras "$@" > gpurun_out/n2_$tag.json 2> gpurun_out/n2_$tag.err; echo "$tag rc=$?"; python -c "
import json;b=json.load(open('gpurun_out/n2_$tag.json'));print('$tag',round(b['ms_per_step'],3),round(b['value'],1),b['extra'].get('exposed_comm_ms'),b['extra']['dp_check']['ok'],b['clocks']['sm_mhz'])"; }
run A
run B --nccl-max-ctas 8
run C --bucket-mb 32
run D --bucket-mb 32 --nccl-max-ctas 4
timeout 200 python -m pytest tests/test_ddp_gpu.py tests/test_unet_gpu.py tests/test_parity_full_gpu.py -m gpu -q -k "nccl and 64 or graph" 2>&1 | tail -4
