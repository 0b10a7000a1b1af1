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


def _cfg(world, rank, variant="rw"):
    from sapt_b200.engine import EngineConfig
    if variant == "ls":   # local-surrogate fork: Philox streams, networks and history windows follow the chains across ranks
        return EngineConfig(variant="ls", samples=40, swap_interval=4, surrogate_interval=10, surrogate_prob=0.5,
                            use_surrogate=True, seed=22, record_history=True, record_pos=False, rank=rank, world=world)
    if variant == "rw_sur":   # global surrogate: the interval's training rows of all chains are assembled on every rank
        return EngineConfig(variant="rw", samples=40, swap_interval=4, surrogate_interval=4, surrogate_prob=0.5, use_surrogate=True,
                            surrogate_topology=1, seed=23, record_history=True, record_pos=True, rank=rank, world=world)
    if variant == "rw_sur_all":   # surrogate_prob = 1: once a model exists no chain makes a true evaluation, so the rendezvous moves NO rows
        return EngineConfig(variant="rw", samples=24, swap_interval=4, surrogate_interval=4, surrogate_prob=1.0, use_surrogate=True,
                            surrogate_topology=1, seed=25, record_history=True, record_pos=True, rank=rank, world=world)
    if variant == "lg_sur":   # Langevin fork: exchanges fall INSIDE a surrogate interval, training rows are captured by lg_first_eval
        return EngineConfig(variant="lg", samples=36, swap_interval=3, surrogate_interval=5, surrogate_prob=0.4, use_surrogate=True,
                            use_langevin=True, learn_rate=0.05, surrogate_topology=1, seed=24, record_history=True, record_pos=True,
                            rank=rank, world=world)
    return EngineConfig(variant="rw", samples=30, swap_interval=3, surrogate_interval=3, surrogate_prob=0.0,
                        use_surrogate=False, seed=21, record_history=True, record_pos=True, rank=rank, world=world)


def _run(rank, world, port, q, variant="rw"):
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
    trainer = None
    if variant in ("rw_sur", "lg_sur", "rw_sur_all"):
        from oracle import sapt_oracle as so      # fixed MLPs stand in for the trained surrogate (one per signature)
        ws = [so.make_surrogate_weights(prob.P, 1, 50 + s) for s in range(12)]
        trainer = lambda batch, sig: ws[sig - 1]
    eng = PTEngine(prob, assign_temperatures(8, 4.0), _cfg(world, rank, variant), surrogate_trainer=trainer)
    eng.initialize(g["W"][0])
    eng.run()
    torch.cuda.synchronize()
    res = eng.results()
    keys = ("pos_w", "accepted", "accept_list", "lhood_list", "acc_train", "naccept", "row_of_slot")
    if variant == "ls":
        keys = ("accepted", "accept_list", "lhood_list", "acc_train", "naccept", "row_of_slot", "surg_likeh_list", "is_surrogate_step",
                "ls_weights", "ls_limit")
    if variant in ("rw_sur", "lg_sur", "rw_sur_all"):
        keys = keys + ("is_surrogate_step", "surg_likeh_list")
    out = {k: res[k] for k in keys}
    if variant in ("rw_sur", "lg_sur", "rw_sur_all"):
        out["surr_batches"] = res["surr_batches"]
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


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_local_surrogate_equals_one_gpu():
    """Variant 'ls' over two ranks: the per-chain networks, their limits and the surg_likeh_list windows they read
    must follow the chains through the exchanges (engine._ls_sync_history, sync_slots)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_run, args=(0, 1, 0, q, "ls")); p.start(); one = q.get(timeout=300); p.join()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, 2, port, q, "ls")) for r in range(2)]
    for p in procs: p.start()
    two = q.get(timeout=300)
    for p in procs: p.join()
    assert one["is_surrogate_step"].sum() > 0 and one["swaps"].sum() > 0
    for k in ("accepted", "is_surrogate_step", "swaps", "row_of_slot", "naccept"):
        assert np.array_equal(one[k], two[k]), k
    a, b = one["surg_likeh_list"], two["surg_likeh_list"]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_allclose(b[~np.isnan(a)], a[~np.isnan(a)], rtol=1e-12)
    np.testing.assert_allclose(two["lhood_list"][np.isfinite(one["lhood_list"])], one["lhood_list"][np.isfinite(one["lhood_list"])], rtol=1e-12)
    np.testing.assert_allclose(two["ls_limit"], one["ls_limit"], rtol=1e-12)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("variant", ["rw_sur", "lg_sur", "rw_sur_all"])
def test_two_gpu_global_surrogate_equals_one_gpu(variant):
    """PT with the global surrogate over two ranks (random-walk fork: exchange and training share the rendezvous;
    Langevin fork: exchanges inside a surrogate interval): the training rows every rank assembles at the
    rendezvous (RW:1105-1116) and every decision that follows must equal the single-GPU run."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_run, args=(0, 1, 0, q, variant)); p.start(); one = q.get(timeout=300); p.join()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, 2, port, q, variant)) for r in range(2)]
    for p in procs: p.start()
    two = q.get(timeout=300)
    for p in procs: p.join()
    assert one["is_surrogate_step"].sum() > 0 and one["swaps"].sum() > 0
    if variant == "rw_sur_all":      # from the first model on every step is a surrogate step: later batches hold the zero rows only
        assert one["is_surrogate_step"][:, 8:20].all() and any(not np.any(b) for b in one["surr_batches"][1:])
    for k in ("accepted", "is_surrogate_step", "swaps", "row_of_slot", "naccept", "pos_w"):
        assert np.array_equal(one[k], two[k]), k
    assert len(one["surr_batches"]) == len(two["surr_batches"]) > 0
    for a, b in zip(one["surr_batches"], two["surr_batches"]):
        assert np.array_equal(a, b)
