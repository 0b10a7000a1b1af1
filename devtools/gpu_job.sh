#!/bin/bash
# last job of round 2: full GPU test-suite, smoke(), the plain-C host, the bench line of the final build
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/j46_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j46_smoke.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/j46_bench.json 2> gpurun_out/j46_bench.err
tail -3 gpurun_out/j46_pytest.txt; tail -1 gpurun_out/j46_smoke.txt; tail -c 200 gpurun_out/j46_bench.err
