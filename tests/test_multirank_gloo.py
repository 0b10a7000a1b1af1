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
        # surrogate rendezvous: training rows STAY on the rank that wrote them, the per-chain counts follow the chains, the
        # trainer rank assembles the chain-by-chain batch (RW:1105-1116) from fixed-size payloads; weights come back by broadcast
        cap, P = 4, 5
        truth_theta = torch.arange(Rg * cap * P, dtype=torch.float32).reshape(Rg, cap, P) + 1.0      # what entry (slot, j) holds, whoever wrote it
        truth_L = -torch.arange(Rg * cap, dtype=torch.float64).reshape(Rg, cap) - 0.25 + 1e-9
        th = torch.full((Rg, cap, P), -777.0); tl = torch.full((Rg, cap), 555.0, dtype=torch.float64)
        cnt = torch.zeros(Rg, dtype=torch.int32); seen = torch.zeros(Rg, dtype=torch.int32); mine = torch.zeros((Rg, cap), dtype=torch.bool)

        def append(slots, n_rows):     # the accept kernel's append, for the slots this rank hosts
            for s_, n_ in zip(slots, n_rows):
                for _ in range(n_):
                    c = int(cnt[s_])
                    if c < cap:
                        th[s_, c] = truth_theta[s_, c]; tl[s_, c] = truth_L[s_, c]; cnt[s_] = c + 1
        # epoch 1 under the first binding (slot_of_row), then an exchange re-binds the chains, then epoch 2
        n1 = {2: 2, 0: 1, 4: 0, 6: 2, 1: 1, 7: 2, 3: 0, 5: 1}
        append(slot_of_row.tolist(), [n1[s_] for s_ in slot_of_row.tolist()])
        parallel.mark_owned(mine, seen, cnt, mask)
        parallel.sync_train_counts(cnt, seen, mask)
        assert cnt.tolist() == [n1[s_] for s_ in range(Rg)] and torch.equal(cnt, seen)
        mask2 = parallel.hosted_mask(new_local, Rg)
        n2 = {s_: (s_ % 3) for s_ in range(Rg)}
        append(new_local.tolist(), [n2[s_] for s_ in new_local.tolist()])
        parallel.mark_owned(mine, seen, cnt, mask2)
        parallel.sync_train_counts(cnt, seen, mask2)
        total = [min(cap, n1[s_] + n2[s_]) for s_ in range(Rg)]
        assert cnt.tolist() == total
        for skip_first, zero_row in ((False, True), (True, False), (False, False)):
            n_pad = R * (cap + 1)
            src, dest, zero_rows, n_out = parallel.training_row_plan(cnt, mine, n_pad, skip_first, zero_row)
            assert src.shape == (n_pad,) and dest.shape == (n_pad,)
            lo, zero = (1 if skip_first else 0), (1 if zero_row else 0)
            want_rows = []
            for s_ in range(Rg):
                if zero:
                    want_rows.append(torch.zeros(P + 1, dtype=torch.float64))
                for j_ in range(lo, total[s_]):
                    want_rows.append(torch.cat([truth_theta[s_, j_].double(), truth_L[s_, j_][None]]))
            want = torch.stack(want_rows) if want_rows else torch.zeros((0, P + 1), dtype=torch.float64)
            assert int(n_out) == want.shape[0]
            # the contract of sapt_train_rows_pack / sapt_train_rows_scatter (include/sapt.h), restated in torch for the CPU
            payload = torch.zeros((n_pad, P + 4), dtype=torch.float32)
            pi = payload.view(torch.int32)
            for i_ in range(n_pad):
                if src[i_] < 0:
                    pi[i_, P + 2] = -1
                    continue
                payload[i_, :P] = th.reshape(-1, P)[src[i_]]
                bits = tl.reshape(-1)[src[i_]].view(torch.int64)
                pi[i_, P] = ((bits & 0xffffffff) - (1 << 32) * ((bits >> 31) & 1)).to(torch.int32)
                pi[i_, P + 1] = (bits >> 32).to(torch.int32)
                pi[i_, P + 2] = dest[i_]
            # with the row count only the front of the payloads travels (here exactly max over ranks of the owned rows)
            allp_short = parallel.gather_payloads(payload, 0, world, n_owned=(src >= 0).sum(), round_rows=1)
            allp = parallel.gather_payloads(payload, 0, world)
            if rank == 0:
                assert allp.shape == (world * n_pad, P + 4)
                m_ = allp_short.shape[0] // world
                assert m_ <= n_pad and all(torch.equal(allp_short.view(torch.int32)[r_ * m_:(r_ + 1) * m_], allp.view(torch.int32)[r_ * n_pad:r_ * n_pad + m_]) for r_ in range(world))
                assert all(int((allp.view(torch.int32)[r_ * n_pad + m_:(r_ + 1) * n_pad, P + 2] >= 0).sum()) == 0 for r_ in range(world))
                batch = torch.full((Rg * (cap + zero), P + 1), 99.0, dtype=torch.float64)
                ai = allp.view(torch.int32)
                for i_ in range(allp.shape[0]):
                    d_ = int(ai[i_, P + 2])
                    if d_ >= 0:
                        batch[d_, :P] = allp[i_, :P].double()
                        lo_w, hi_w = int(ai[i_, P]) & 0xffffffff, int(ai[i_, P + 1]) & 0xffffffff
                        batch[d_, P] = torch.tensor([(hi_w << 32) | lo_w], dtype=torch.int64).view(torch.float64)[0] if (hi_w >> 31) == 0 else \
                            torch.tensor([((hi_w << 32) | lo_w) - (1 << 64)], dtype=torch.int64).view(torch.float64)[0]
                for z_ in zero_rows.tolist():
                    batch[z_] = 0.0
                assert torch.equal(batch[: int(n_out)], want), (skip_first, zero_row)
            else:
                assert allp is None and allp_short is None
        # every entry is owned by exactly one rank
        owned = mine.to(torch.int32)
        dist.all_reduce(owned)
        jj = torch.arange(cap)[None, :]
        assert torch.equal(owned, (jj < cnt[:, None]).to(torch.int32))
        # the retrained model goes back as one flat fp32 vector + [has_weights, minY, maxY]
        shapes = ((P, 3), (3,), (3, 2), (2,), (2, 1), (1,))
        wts = {k: torch.arange(int(np.prod(sh)), dtype=torch.float32).reshape(sh) + 10 * n_ for n_, (k, sh) in enumerate(zip(parallel.SURROGATE_WEIGHT_KEYS, shapes))}
        head, flat = parallel.broadcast_surrogate_device(wts if rank == 0 else None, shapes, -123.5, -1.0, 0, torch.device("cpu"))
        assert head.tolist() == [1.0, -123.5, -1.0]
        got = parallel.split_flat(flat, shapes)
        for k in parallel.SURROGATE_WEIGHT_KEYS:
            assert torch.equal(got[k], wts[k].reshape(-1))
        head, flat = parallel.broadcast_surrogate_device(None, shapes, 0.0, 0.0, 0, torch.device("cpu"))
        assert head[0] == 0
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
