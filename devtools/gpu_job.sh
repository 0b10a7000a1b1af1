#!/bin/bash
# last job of round 2: smoke() and the bench line of the final build (profiles/traffic.json is current)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j55_smoke.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/j55_bench.json 2> gpurun_out/j55_bench.err
tail -1 gpurun_out/j55_smoke.txt; tail -c 200 gpurun_out/j55_bench.err
