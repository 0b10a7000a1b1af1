#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 devtools/time_rendezvous_mp.py > gpurun_out/m8_rdv.txt 2> gpurun_out/m8_rdv.err
tail -c 600 gpurun_out/m8_rdv.err; cat gpurun_out/m8_rdv.txt
