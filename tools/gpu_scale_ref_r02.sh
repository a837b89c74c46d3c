#!/bin/bash
# the driver's two arms at N GPUs: reference first, then the product
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r02_scale_ref_$N.json 2> gpurun_out/r02_scale_ref_$N.err; echo "reference N=$N rc=$?"; cut -c1-300 gpurun_out/r02_scale_ref_$N.json
bash tools/gpu_scale_r02.sh $N
