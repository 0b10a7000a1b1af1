#!/bin/bash
# soak: the two-stream batch-Langevin kernel and the likelihood kernel, repeated (rare races / lost hand-overs would show as a
# wrong gradient, a trap from the bounded barrier waits, or a hang)
mkdir -p gpurun_out
rm -f gpurun_out/j48_soak.txt
for i in 1 2 3 4 5 6; do
  timeout 600 python -m pytest tests -m gpu -q -x -k "langevin or loglik or lg_ or c4 or sweep or swap" 2>&1 | tail -1 >> gpurun_out/j48_soak.txt
done
timeout 600 python - >> gpurun_out/j48_soak.txt 2>&1 <<'PY'
import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from sapt_b200 import _lib
from sapt_b200.problem import Problem
rs = np.random.RandomState(0)
D, H, O, N, R = 54, 64, 7, 200_000, 96
x = rs.standard_normal((N, D)).astype(np.float32); y = rs.randint(0, O, N).astype(np.int32)
prob = Problem.from_f32([D, H, O], x, y)
th = torch.as_tensor((rs.standard_normal((R, prob.P)) * 0.4).astype(np.float32)).cuda()
ref_g = prob.langevin(th, 0.01, mode=1, batch_scale=1.0 / N).clone(); ref_l = prob.loglik(th, 0, path=_lib.PATH_TC)[0].clone()
bad = 0
for it in range(300):
    g = prob.langevin(th, 0.01, mode=1, batch_scale=1.0 / N); l = prob.loglik(th, 0, path=_lib.PATH_TC)[0]
    if not (torch.equal(g, ref_g) and torch.equal(l, ref_l)): bad += 1
torch.cuda.synchronize()
print("300 repeats of batch-Langevin + likelihood at 54-64-7 x 200000 rows x 96 chains: results bit-identical to the first run in", 300 - bad, "of 300")
PY
cat gpurun_out/j48_soak.txt
