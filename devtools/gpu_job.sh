#!/bin/bash
mkdir -p gpurun_out
timeout 600 python devtools/tc_shapes.py > gpurun_out/j37_tc_shapes.txt 2>&1
cat gpurun_out/j37_tc_shapes.txt
