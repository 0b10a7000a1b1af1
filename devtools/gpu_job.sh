#!/bin/bash
# last job of round 2: the bench line of the final build (profiles/traffic.json is current), smoke()
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j53_smoke.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/j53_bench.json 2> gpurun_out/j53_bench.err
tail -1 gpurun_out/j53_smoke.txt; tail -c 200 gpurun_out/j53_bench.err
