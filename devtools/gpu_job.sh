#!/bin/bash
mkdir -p gpurun_out
timeout 600 python devtools/x_streaming.py > gpurun_out/j39_xstream.txt 2>&1
cat gpurun_out/j39_xstream.txt
