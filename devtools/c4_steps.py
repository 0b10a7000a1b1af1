"""A few steps of the C4 leg of bench.py alone (for `ncu --metrics gpu__time_duration.sum` launch lists). usage: python devtools/c4_steps.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sapt_b200.engine import EngineConfig, PTEngine
from sapt_b200.ladder import assign_temperatures
from sapt_b200.problem import Problem
W = bench.C4
D, H, O, N, Rl = W["D"], W["H"], W["O"], W["n_train"], W["chains_per_gpu"]
x, y = bench.synth(D, O, N, 1); xt, yt = bench.synth(D, O, W["n_test"], 2)
prob = Problem.from_f32([D, H, O], x, y, xt, yt, device=0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = EngineConfig(variant="lg", samples=3 + n + 2, swap_interval=100000, surrogate_interval=100000, surrogate_prob=0.0, use_surrogate=False, use_langevin=True,
                   learn_rate=0.5, langevin_mode="batch", langevin_batch_scale=1.0 / N, seed=W["seed"], record_pos=False)
eng = PTEngine(prob, assign_temperatures(Rl, W["maxtemp"]), cfg)
eng.initialize(np.random.RandomState(3).randn(prob.P) * 0.5)
for i in range(3):
    eng.step(i); eng.rendezvous(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
e0.record()
for i in range(3, 3 + n):
    eng.step(i); eng.rendezvous(i)
e1.record(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"C4 leg: {e0.elapsed_time(e1) / n:.3f} ms per step over {n} steps")
