import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import golden_util as gu
from sapt_b200.problem import Problem
for name in ("unit_ions", "unit_cancer", "unit_pendigit"):
    g = gu.load(name); topo = [int(v) for v in g["topology"]]
    prob = Problem(topo, g["train"], g["test"])
    th = torch.as_tensor(np.tile(g["W"][:2], (5, 1)).astype(np.float32)).cuda()
    out = prob.langevin(th, 0.01, mode=0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): out = prob.langevin(th, 0.01, mode=0, out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"sgd_exact {name} {topo} {g['train'].shape[0]} rows x 10 chains: {us:.1f} us per call, {us / g['train'].shape[0] * 1e3:.0f} ns/row")
