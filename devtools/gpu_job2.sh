#!/bin/bash
# 2-GPU job: multi-rank parity tests + the N = 2 bench line (strong-scaling leg, sharding self-check, C4 / C5 legs over two ranks)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/m2_smi.txt
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/m2_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/m2_bench2.json 2> gpurun_out/m2_bench2.err
tail -c 300 gpurun_out/m2_pytest.txt; tail -c 300 gpurun_out/m2_bench2.err; head -c 300 gpurun_out/m2_bench2.json
