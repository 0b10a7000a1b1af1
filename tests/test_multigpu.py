"""2-GPU test (NCCL): sharding the chains over two GPUs must not change the chains.
Same seed, same data: accept/reject decisions, swap decisions and posterior samples of the
2-rank run equal the 1-GPU run (SURVEY.md §4 iv). Skipped when fewer than 2 GPUs are visible."""
import os
import socket

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _cfg(world, rank):
    from sapt_b200.engine import EngineConfig
    return EngineConfig(variant="rw", samples=30, swap_interval=3, surrogate_interval=3, surrogate_prob=0.0,
                        use_surrogate=False, seed=21, record_history=True, record_pos=True, rank=rank, world=world)


def _run(rank, world, port, q):
    import torch.distributed as dist
    from sapt_b200.engine import PTEngine
    from sapt_b200.ladder import assign_temperatures
    from sapt_b200.problem import Problem
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    g = gu.load("unit_cancer")
    topo = [int(v) for v in g["topology"]]
    prob = Problem(topo, g["train"], g["test"], device=rank)
    eng = PTEngine(prob, assign_temperatures(8, 4.0), _cfg(world, rank))
    eng.initialize(g["W"][0])
    eng.run()
    torch.cuda.synchronize()
    res = eng.results()
    out = {k: res[k] for k in ("pos_w", "accepted", "accept_list", "lhood_list", "acc_train", "naccept", "row_of_slot")}
    out["swaps"] = np.array([s for s in res["swap_decisions"]])
    out["num_swap"] = res["num_swap"]
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        q.put(out)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_run_equals_one_gpu_run():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_run, args=(0, 1, 0, q)); p.start(); one = q.get(timeout=300); p.join()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    two = q.get(timeout=300)
    for p in procs: p.join()
    assert np.array_equal(one["accepted"], two["accepted"])
    assert np.array_equal(one["swaps"], two["swaps"]) and one["num_swap"] == two["num_swap"]
    assert np.array_equal(one["row_of_slot"], two["row_of_slot"])
    assert np.array_equal(one["pos_w"], two["pos_w"])
    assert np.array_equal(one["naccept"], two["naccept"])
    np.testing.assert_allclose(one["lhood_list"], two["lhood_list"], rtol=1e-12)
    assert one["swaps"].sum() > 0
