"""Times the pieces of an exchange + surrogate rendezvous of the C5 leg of bench.py on N ranks (torchrun), CUDA events per rank, max over ranks.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 devtools/time_rendezvous_mp.py [chains_per_gpu]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from sapt_b200 import _lib, parallel
from sapt_b200.engine import EngineConfig, PTEngine
from sapt_b200.ladder import assign_temperatures
from sapt_b200.problem import Problem, _ptr, current_stream_ptr

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W = bench.C5
Rl = int(sys.argv[1]) if len(sys.argv) > 1 else W["chains_per_gpu"]
D, H, O, N = W["D"], W["H"], W["O"], W["n_train"]
x, y = bench.synth(D, O, N, 5); xt, yt = bench.synth(D, O, W["n_test"], 6)
prob = Problem.from_f32([D, H, O], x, y, xt, yt, device=local)
w = bench._fixed_surrogate(prob.P, 200, 50, 7)
cfg = EngineConfig(variant="rw", samples=64, swap_interval=W["swap_interval"], surrogate_interval=W["swap_interval"], surrogate_prob=0.5, use_surrogate=True,
                   surrogate_topology=3, seed=W["seed"], record_pos=False, train_cap=W["swap_interval"], rank=rank, world=world)
eng = PTEngine(prob, assign_temperatures(Rl * world, W["maxtemp"]), cfg, surrogate_trainer=bench.FixedTrainer(w))
eng.keep_batches = False
eng.initialize(np.random.RandomState(3).randn(prob.P) * 0.5)
for i in range(12):
    eng.step(i); eng.rendezvous(i)

def timed(fn):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = fn(); b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), r

out = {}
acc = lambda k, v: out.setdefault(k, []).append(v)
for i in range(12, 28):
    t, _ = timed(lambda: eng.step(i)); acc("step", t)
    if i % W["swap_interval"] == 0:
        t, _ = timed(eng.sync_slots); acc("swap.sync_slots", t)
        t, _ = timed(eng._sync_training_counts); acc("swap.sync_training_counts", t)
        t, _ = timed(lambda: eng.swap(i)); acc("swap.rest (sweep + bookkeeping; slots / counts already synced)", t)
        # the rendezvous, piece by piece (mirrors PTEngine.gather_training_rows / surrogate_rendezvous)
        t, plan = timed(lambda: parallel.training_row_plan(eng.train_count, eng.train_mine, eng.train_npad, skip_first=False, zero_row=eng.spec.ship_zero_row)); acc("train.plan", t)
        src, dest, zero_rows, n_out = plan
        t, _ = timed(lambda: _lib.check(eng.L.sapt_train_rows_pack(_ptr(eng.train_theta), _ptr(eng.train_L), _ptr(src), _ptr(dest), eng.train_npad, eng.P, _ptr(eng.train_payload), current_stream_ptr()))); acc("train.pack", t)
        t, allp = timed(lambda: parallel.gather_payloads(eng.train_payload, eng.trainer_rank, world, eng.pg)); acc("train.gather", t)
        del allp
        t, _ = timed(lambda: eng.gather_training_rows()); acc("train.gather_training_rows (plan + pack + gather + scatter, again)", t)
        t, _ = timed(lambda: eng.surrogate_rendezvous(i)); acc("train.surrogate_rendezvous (whole, on the emptied buffers: trainer + broadcast)", t)
if rank == 0:
    print(f"world {world}, {Rl} chains per GPU, payload {eng.train_payload.numel() * 4 / 1e6:.0f} MB per rank, valid rows fraction {float(n_out) / (Rl * world * (W['swap_interval'] + 1)):.2f}")
    for k, v in out.items():
        print(f"  {k:70s} {np.mean(v):8.3f} ms (n={len(v)})")
if world > 1:
    dist.destroy_process_group()
