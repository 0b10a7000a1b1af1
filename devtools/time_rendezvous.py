"""Times the pieces of one rendezvous (exchange + surrogate training-set assembly) at 65 536 chains on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sapt_b200.engine import EngineConfig, PTEngine
from sapt_b200.ladder import assign_temperatures
from sapt_b200.problem import Problem
from oracle import sapt_oracle as so
D, H, O, N = 6, 25, 18, 2048
rs = np.random.RandomState(5)
x = rs.standard_normal((N, D)).astype(np.float32); y = rs.randint(0, O, N).astype(np.int32)
prob = Problem.from_f32([D, H, O], x, y, x[:512], y[:512])
w = so.make_surrogate_weights(prob.P, 3, 7)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = EngineConfig(variant="rw", samples=64, swap_interval=4, surrogate_interval=4, surrogate_prob=0.5, use_surrogate=True, surrogate_topology=3,
                   seed=4, record_pos=False, train_cap=4)
eng = PTEngine(prob, assign_temperatures(R, 4.0), cfg, surrogate_trainer=lambda batch, sig: (w if sig == 1 else None))
eng.surr_batches = type("Drop", (list,), {"append": lambda self, x: None})()
eng.initialize(np.random.RandomState(3).randn(prob.P) * 0.5)
def timed(fn, n=1):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)
for i in range(12):
    eng.step(i); eng.rendezvous(i)
for i in range(12, 20):
    eng.step(i)
    if i % 4 == 0:
        t_swap = timed(lambda: eng.swap(i))
        t_train = timed(lambda: eng.surrogate_rendezvous(i))
        th = eng.train_theta
        t_mul = timed(lambda: th.mul_(1.0))
        print(f"R={R} step {i}: swap {t_swap:.3f} ms, surrogate_rendezvous {t_train:.3f} ms, in-place pass over train_theta {t_mul:.3f} ms ({th.numel()*4/1e6:.0f} MB)", flush=True)
# the sweep kernel alone (C-ABI call), without the host-side bookkeeping of PTEngine.swap
from sapt_b200 import _lib
from sapt_b200.problem import _ptr, current_stream_ptr
sw = torch.zeros(R - 1, dtype=torch.uint8, device="cuda")
def raw():
    _lib.check(eng.L.sapt_swap_sweep(R, _ptr(eng.lik_post), _ptr(eng.T), _ptr(eng.row_of_slot), eng.flags, 4, 99, None, _ptr(sw),
                                     _ptr(eng.counters[3:4]), None, eng.row0, eng.R, _ptr(eng.slot_of_row), current_stream_ptr()))
raw()
print(f"R={R}: sapt_swap_sweep alone {timed(raw):.3f} ms, swapped fraction {sw.float().mean().item():.3f}", flush=True)
