import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import golden_util as gu
from sapt_b200 import _lib
from sapt_b200.engine import EngineConfig, PTEngine
from sapt_b200.ladder import assign_temperatures
from sapt_b200.problem import Problem
g = gu.load("unit_cancer"); topo = [int(v) for v in g["topology"]]
for path, nm in ((_lib.PATH_F32, "f32"), (_lib.PATH_TC, "tc")):
    prob = Problem(topo, g["train"], g["test"], path=path)
    for graph in (False, True):
        cfg = EngineConfig(variant="rw", samples=3001, swap_interval=50, surrogate_interval=50, surrogate_prob=0.0, use_surrogate=False, seed=3, path=path, record_pos=False)
        eng = PTEngine(prob, assign_temperatures(8, 4.0), cfg)
        eng.initialize(g["W"][0])
        torch.cuda.synchronize(); t0 = time.time()
        eng.run(graph=graph); torch.cuda.synchronize()
        dt = time.time() - t0
        print(nm, "graph" if graph else "eager", "us/step %.1f" % (dt / 3000 * 1e6), flush=True)
