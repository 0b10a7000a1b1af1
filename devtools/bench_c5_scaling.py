"""C5 weak scaling (Chess-shape 6-25-18, surrogate_prob 0.5, fixed surrogate MLP): R chains PER GPU, one process per
GPU (torchrun). Prints one JSON line on rank 0: ms/step and samples/s over all GPUs."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sapt_b200.engine import EngineConfig, PTEngine
from sapt_b200.ladder import assign_temperatures
from sapt_b200.problem import Problem
from oracle import sapt_oracle as so   # only for the fixed random surrogate weights

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
R_per = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D, H, O, N = 6, 25, 18, 16_833
rs = np.random.RandomState(5)
x = rs.standard_normal((N, D)).astype(np.float32); y = rs.randint(0, O, N).astype(np.int32)
xt = rs.standard_normal((11_223, D)).astype(np.float32); yt = rs.randint(0, O, 11_223).astype(np.int32)
prob = Problem.from_f32([D, H, O], x, y, xt, yt, device=local)
w = so.make_surrogate_weights(prob.P, 3, 7)
R = R_per * world
cfg = EngineConfig(variant="rw", samples=64, swap_interval=4, surrogate_interval=4, surrogate_prob=0.5, use_surrogate=True, surrogate_topology=3,
                   seed=4, record_pos=False, train_cap=4, rank=rank, world=world)
eng = PTEngine(prob, assign_temperatures(R, 4.0), cfg, surrogate_trainer=lambda batch, sig: (w if sig == 1 else None))
eng.surr_batches = type("Drop", (list,), {"append": lambda self, x: None})()
eng.initialize(np.random.RandomState(3).randn(prob.P) * 0.5)
for i in range(12):
    eng.step(i); eng.rendezvous(i)
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 16
e0.record()
for i in range(12, 12 + steps):
    eng.step(i); eng.rendezvous(i)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps(dict(config="C5 chess-shape 6-25-18, surrogate_prob 0.5", n_gpus=world, chains_per_gpu=R_per, chains=R, ms_per_step=ms.item(),
                          samples_per_s=R / (ms.item() * 1e-3))), flush=True)
if world > 1: dist.destroy_process_group()
