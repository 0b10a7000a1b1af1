"""C2 (Ionosphere-shape, Langevin fork, no surrogate): step-by-step Python chaining vs CUDA-graph replay."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sapt_b200.engine import EngineConfig, PTEngine
from sapt_b200.ladder import assign_temperatures
from sapt_b200.problem import Problem
D, H, O, R, S = 34, 50, 2, 10, 600
rs = np.random.RandomState(7)
mk = lambda n: np.concatenate([rs.standard_normal((n, D)), rs.randint(0, O, n)[:, None].astype(np.float64)], axis=1)
tr, te = mk(245), mk(109)
prob = Problem([D, H, O], tr, te)
for graph in (False, True, False, True):
    cfg = EngineConfig(variant="lg", samples=S, swap_interval=100000, surrogate_interval=100000, use_surrogate=False, use_langevin=True,
                       learn_rate=0.01, seed=3, record_pos=False)
    eng = PTEngine(prob, assign_temperatures(R, 4.0), cfg)
    eng.initialize(np.random.RandomState(3).randn(prob.P) * 0.5)
    torch.cuda.synchronize(); t0 = time.time()
    eng.run(graph=graph)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"graph={graph}: {dt:.3f} s, {(S - 1) / dt:.0f} steps/s, naccept {eng.naccept.cpu().numpy()[:4]}", flush=True)
