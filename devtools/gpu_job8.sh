#!/bin/bash
# 8-GPU job: the N = 8 bench line (weak + strong C3, sharding self-check, C4 / C5 legs over eight ranks) + per-piece rendezvous timing
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/m8_bench8.json 2> gpurun_out/m8_bench8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 devtools/time_rendezvous_mp.py > gpurun_out/m8_rdv.txt 2> gpurun_out/m8_rdv.err
tail -c 300 gpurun_out/m8_bench8.err; head -c 300 gpurun_out/m8_bench8.json; cat gpurun_out/m8_rdv.txt
