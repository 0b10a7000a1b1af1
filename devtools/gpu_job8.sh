#!/bin/bash
# 8-GPU job: per-piece timing of an exchange + rendezvous of the C5 leg (the N = 8 bench line is profiles/r2_bench_8gpu.json)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 devtools/time_rendezvous_mp.py > gpurun_out/m8_rdv.txt 2> gpurun_out/m8_rdv.err
tail -c 300 gpurun_out/m8_rdv.err; cat gpurun_out/m8_rdv.txt
