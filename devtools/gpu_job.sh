#!/bin/bash
# last job of round 2: smoke(), the plain-C host, and the bench line of the final build (traffic.json is current)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j43_smoke.txt 2>&1
(cd examples/c_host && ./run_pt) > gpurun_out/j43_chost.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/j43_bench.json 2> gpurun_out/j43_bench.err
tail -2 gpurun_out/j43_smoke.txt; tail -3 gpurun_out/j43_chost.txt; tail -c 200 gpurun_out/j43_bench.err
