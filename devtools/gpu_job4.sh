#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/m4_bench4.json 2> gpurun_out/m4_bench4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/m4_ref4.json 2> gpurun_out/m4_ref4.err
tail -c 300 gpurun_out/m4_bench4.err; head -c 300 gpurun_out/m4_bench4.json; echo; head -c 300 gpurun_out/m4_ref4.json
