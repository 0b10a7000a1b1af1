"""A few steps of the C5 leg of bench.py alone (for `ncu --metrics gpu__time_duration.sum` launch lists). usage: python devtools/c5_steps.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sapt_b200.engine import EngineConfig, PTEngine
from sapt_b200.ladder import assign_temperatures
from sapt_b200.problem import Problem
W = bench.C5
D, H, O, N, Rl = W["D"], W["H"], W["O"], W["n_train"], W["chains_per_gpu"]
x, y = bench.synth(D, O, N, 5); xt, yt = bench.synth(D, O, W["n_test"], 6)
prob = Problem.from_f32([D, H, O], x, y, xt, yt, device=0)
w = bench._fixed_surrogate(prob.P, 200, 50, 7)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = EngineConfig(variant="rw", samples=12 + n + 2, swap_interval=W["swap_interval"], surrogate_interval=W["swap_interval"], surrogate_prob=0.5, use_surrogate=True,
                   surrogate_topology=3, seed=W["seed"], record_pos=False, train_cap=W["swap_interval"])
eng = PTEngine(prob, assign_temperatures(Rl, W["maxtemp"]), cfg, surrogate_trainer=bench.FixedTrainer(w))
eng.keep_batches = False
eng.initialize(np.random.RandomState(3).randn(prob.P) * 0.5)
for i in range(12):
    eng.step(i); eng.rendezvous(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
e0.record()
for i in range(12, 12 + n):
    eng.step(i); eng.rendezvous(i)
e1.record(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"C5 leg: {e0.elapsed_time(e1) / n:.3f} ms per step over {n} steps")
