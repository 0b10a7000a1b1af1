"""world_size=2 gloo tests (CPU) of the multi-GPU host logic: sharding, the masked all-reduce
that re-synchronises per-chain scalars before an exchange, the deterministic sweep every rank
replays, the local re-binding, and the history merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sapt_oracle as so


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sapt_b200 import parallel
        Rg = 8
        row0, R = parallel.shard_bounds(Rg, rank, world)
        assert (row0, R) == (rank * 4, 4)
        # binding after some earlier exchange: slots are NOT contiguous per rank any more
        row_of_slot = torch.tensor([1, 4, 0, 6, 2, 7, 3, 5], dtype=torch.int32)
        slot_of_row = parallel.rebind_local(row_of_slot, row0, R)
        expect = {0: [2, 0, 4, 6], 1: [1, 7, 3, 5]}[rank]
        assert slot_of_row.tolist() == expect
        mask = parallel.hosted_mask(slot_of_row, Rg)
        # ground truth per slot; each rank only knows its hosted slots, the rest is stale garbage (incl. NaN)
        truth = torch.arange(parallel.PACK_ROWS * Rg, dtype=torch.float64).reshape(parallel.PACK_ROWS, Rg) * 0.5 - 7
        truth[4, 3] = float("inf")             # an 'inf' history value of a surrogate step must survive
        truth[11] = torch.arange(Rg)           # naccept
        mine = torch.where(mask[None, :] > 0, truth, torch.full_like(truth, float("nan")))
        synced = parallel.sync_hosted(mine, mask)
        assert torch.equal(synced, truth), "masked all-reduce did not reconstruct the chain table"
        lik_cur = torch.zeros(Rg, dtype=torch.float64); lik_post = torch.zeros_like(lik_cur); prior = torch.zeros_like(lik_cur)
        lwt = torch.zeros_like(lik_cur); last = torch.zeros((4, Rg), dtype=torch.float64); mov = torch.zeros((Rg, 3), dtype=torch.float64)
        nacc = torch.zeros(Rg, dtype=torch.int32)
        parallel.unpack_state(synced, lik_cur, lik_post, prior, lwt, last, mov, nacc)
        assert torch.equal(parallel.pack_state(lik_cur, lik_post, prior, lwt, last, mov, nacc), truth)
        assert nacc.tolist() == list(range(Rg))
        # every rank replays the same sweep with counter-based uniforms -> identical new binding
        u = so.philox_uniform(123, so.KIND_SWAP, 5, np.zeros(Rg - 1, dtype=np.uint32), np.arange(Rg - 1)).astype(np.float64)
        src, swapped = so.swap_sweep(lik_post.numpy() * 1.0, u)
        new_ros = row_of_slot[torch.as_tensor(src).long()]
        gathered = [torch.zeros_like(new_ros) for _ in range(world)]
        dist.all_gather(gathered, new_ros)
        assert all(torch.equal(g, new_ros) for g in gathered)
        new_local = parallel.rebind_local(new_ros, row0, R)
        assert (new_local >= 0).all()
        allslots = [torch.zeros_like(new_local) for _ in range(world)]
        dist.all_gather(allslots, new_local)
        assert sorted(torch.cat(allslots).tolist()) == list(range(Rg)), "every chain must be hosted exactly once"
        # history merge: hosted entries written, others zero
        hist = torch.zeros((3, Rg), dtype=torch.float64); hist[:, slot_of_row.long()] = 1.0 + rank
        merged = parallel.merge_history(hist)
        want = torch.zeros((3, Rg), dtype=torch.float64)
        want[:, [2, 0, 4, 6]] = 1.0; want[:, [1, 7, 3, 5]] = 2.0
        assert torch.equal(merged, want)
        h8 = torch.zeros((2, Rg), dtype=torch.uint8); h8[:, slot_of_row.long()] = 1
        assert parallel.merge_history(h8).sum().item() == 2 * Rg
        # surrogate training rows follow the chains: each rank holds the rows of the slots it hosts (stale garbage elsewhere)
        cap, P = 3, 5
        full_cnt = torch.tensor([3, 0, 2, 1, 3, 3, 0, 2], dtype=torch.int32)
        full_th = torch.arange(Rg * cap * P, dtype=torch.float32).reshape(Rg, cap, P) + 1.0
        full_L = -torch.arange(Rg * cap, dtype=torch.float64).reshape(Rg, cap) - 0.25
        hosted = mask > 0
        th = torch.where(hosted[:, None, None], full_th, torch.full_like(full_th, -777.0))
        tl = torch.where(hosted[:, None], full_L, torch.full_like(full_L, 555.0))
        cnt = torch.where(hosted, full_cnt, torch.full_like(full_cnt, 9))
        parallel.sync_training_buffers(th, tl, cnt, mask)
        assert torch.equal(th, full_th) and torch.equal(tl, full_L) and torch.equal(cnt, full_cnt)
        # history rows written since the last exchange: non-zero on the writing rank only, NaN markers included
        rows = torch.zeros((4, Rg, 3), dtype=torch.float64)
        rows[:, slot_of_row.long(), :] = float(rank + 1)
        rows[1, slot_of_row.long()[0], 1] = float("nan")
        parallel.sync_history_rows(rows)
        want_rows = torch.zeros((4, Rg, 3), dtype=torch.float64)
        want_rows[:, [2, 0, 4, 6], :] = 1.0; want_rows[:, [1, 7, 3, 5], :] = 2.0
        nan_at = [(1, 2, 1), (1, 1, 1)]                      # first hosted slot of rank 0 / rank 1
        for (a_, b_, c_) in nan_at:
            assert torch.isnan(rows[a_, b_, c_])
            rows[a_, b_, c_] = want_rows[a_, b_, c_]
        assert torch.equal(rows, want_rows)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"FAIL {type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_bounds_rejects_uneven():
    from sapt_b200 import parallel
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 0, 4)
    assert parallel.shard_bounds(1024, 3, 8) == (384, 128)
