#!/bin/bash
mkdir -p gpurun_out
timeout 300 python devtools/time_rendezvous_mp.py > gpurun_out/j45_rdv1.txt 2>&1
tail -10 gpurun_out/j45_rdv1.txt
