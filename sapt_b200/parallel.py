"""Multi-GPU host logic (device-agnostic torch code, so it is testable with gloo on CPU).

Chains shard contiguously over ranks: rank g owns theta rows [g*R/G, (g+1)*R/G). The per-chain
scalars are indexed by GLOBAL chain id ("slot") and every rank keeps full-size arrays but only
updates the slots its rows host. Before an exchange the arrays are re-synchronised with ONE
all-reduce of a packed [K, R_global] fp64 buffer in which each rank zeroes the slots it does not
host (NCCL over NVLink on GPUs; the payload is <= 12 doubles per chain). Then every rank runs the
same deterministic sweep (counter-based uniforms), so the new binding is identical everywhere and
parameter vectors never cross GPUs. This replaces the reference's per-chain Queue.put/get of
P+4 float64 (surrogate_rw_common_interval.py:623,1095-1099)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

PACK_ROWS = 12  # lik_cur, lik_post, prior_cur, likl_wo_temp, last[4], mov[3], naccept


def shard_bounds(n_chains: int, rank: int, world: int) -> Tuple[int, int]:
    if n_chains % world != 0:
        raise ValueError(f"num_chains={n_chains} must be a multiple of the number of ranks ({world})")
    r = n_chains // world
    return rank * r, r


def hosted_mask(slot_of_row: torch.Tensor, n_slots: int) -> torch.Tensor:
    m = torch.zeros(n_slots, dtype=torch.float64, device=slot_of_row.device)
    m[slot_of_row.long()] = 1.0
    return m


def pack_state(lik_cur, lik_post, prior_cur, likl_wo_temp, last, mov, naccept) -> torch.Tensor:
    return torch.cat([lik_cur[None], lik_post[None], prior_cur[None], likl_wo_temp[None], last, mov.t(),
                      naccept.to(torch.float64)[None]], dim=0)


def unpack_state(pack, lik_cur, lik_post, prior_cur, likl_wo_temp, last, mov, naccept) -> None:
    lik_cur.copy_(pack[0]); lik_post.copy_(pack[1]); prior_cur.copy_(pack[2]); likl_wo_temp.copy_(pack[3])
    last.copy_(pack[4:8]); mov.copy_(pack[8:11].t()); naccept.copy_(pack[11].round().to(naccept.dtype))


def sync_hosted(pack: torch.Tensor, mask: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce(sum) of `pack` with non-hosted slots zeroed: afterwards every rank holds every
    chain's current scalars. NaN/inf in hosted slots survive (x + 0 == x); non-hosted entries are
    replaced by exact zeros first so stale NaNs cannot leak."""
    import torch.distributed as dist
    out = torch.where(mask[None, :] > 0, pack, torch.zeros_like(pack))
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def rebind_local(row_of_slot: torch.Tensor, row0: int, n_local: int) -> torch.Tensor:
    """slot_of_row for the rank-local rows, from the global row_of_slot (what the sweep kernel
    also refreshes in place on the device)."""
    ros = row_of_slot.long()
    sel = (ros >= row0) & (ros < row0 + n_local)
    out = torch.full((n_local,), -1, dtype=torch.int32, device=row_of_slot.device)
    out[(ros[sel] - row0)] = torch.nonzero(sel).squeeze(-1).to(torch.int32)
    return out


def merge_history(hist: torch.Tensor, group=None) -> torch.Tensor:
    """Histories are written only for hosted slots (zeros elsewhere): the merged history is the sum."""
    import torch.distributed as dist
    t = hist.clone() if hist.dtype != torch.uint8 else hist.to(torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


# ---- surrogate rendezvous: training rows stay where they were written -----------------------------------------------
# A chain's (theta', L) rows are appended to slot-indexed buffers on whichever rank hosts the chain at that step; the
# per-chain count travels with the chain (one [R_global] int32 all-reduce per exchange), so the new host continues at the
# right index and every rank knows which entries it wrote itself ("owned"). At the rendezvous each rank packs its owned
# rows with their final position in the chain-by-chain batch (RW:1105-1116) into a fixed-size payload, the payloads are
# gathered on the trainer rank, scattered into the batch, and only the retrained weights (<= 3.2 MB) come back by
# broadcast. Nothing here has a data-dependent shape except the length of the gather (see gather_payloads).

def mark_owned(train_mine: torch.Tensor, train_seen: torch.Tensor, train_count: torch.Tensor, mask: torch.Tensor) -> None:
    """Entries [seen, count) of the slots this rank hosts were written here since the counts were last synchronised."""
    cap = train_mine.shape[1]
    j = torch.arange(cap, device=train_mine.device)[None, :]
    new = (mask[:, None] > 0) & (j >= train_seen[:, None]) & (j < train_count[:, None])
    train_mine.logical_or_(new)


def sync_train_counts(train_count: torch.Tensor, train_seen: torch.Tensor, mask: torch.Tensor, group=None) -> None:
    """The count of a chain is current on the rank that hosts it (stale elsewhere): take it from there, everywhere."""
    import torch.distributed as dist
    cnt = torch.where(mask > 0, train_count, torch.zeros_like(train_count))
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    train_count.copy_(cnt)
    train_seen.copy_(cnt)


def training_row_plan(train_count: torch.Tensor, train_mine, n_pad: int, skip_first: bool, zero_row: bool, return_owned: bool = False):
    """Index vectors of the rendezvous, all with static shapes (device-side, no host synchronisation):
    src [n_pad] int32   : slot * cap + j of the i-th row this rank owns, chain-major, -1 for padding
    dest [n_pad] int32  : its row in the chain-by-chain batch (RW:1105-1116): chains in order, within a chain step order, after
                          the chain's all-zero first row when `zero_row` (RW:496,626); `skip_first` drops entry 0 (RV:683)
    zero_rows [R_global] int64 : batch row of every chain's zero row (empty when not shipped)
    n_out [] int64      : rows of the batch
    train_mine None = one rank owns everything it counts.
    return_owned: also return n_owned [] int64, the number of rows of this rank's payload that carry data (they come first).
    On a CUDA device all of this is ONE kernel (sapt_train_rows_plan); the torch code below is its restatement for the CPU tests
    (two dozen small kernels when run on the device: 0.4 ms per rendezvous at 8192 chains)."""
    Rg = train_count.shape[0]
    dev = train_count.device
    if dev.type == "cuda" and train_mine is not None:
        from . import _lib
        from .problem import _ptr, current_stream_ptr
        cap = train_mine.shape[1]
        cnt32 = train_count.to(torch.int32).contiguous()
        mine8 = train_mine.contiguous().view(torch.uint8)
        src = torch.empty(n_pad, dtype=torch.int32, device=dev); dest = torch.empty(n_pad, dtype=torch.int32, device=dev)
        zero_rows = torch.empty(Rg if zero_row else 0, dtype=torch.int64, device=dev)
        scal = torch.empty(2, dtype=torch.int64, device=dev)
        _lib.check(_lib.lib().sapt_train_rows_plan(_ptr(cnt32), _ptr(mine8), Rg, cap, n_pad, 1 if skip_first else 0, 1 if zero_row else 0, _ptr(src), _ptr(dest),
                                                   _ptr(zero_rows) if zero_row else None, scal[0:1].data_ptr(), scal[1:2].data_ptr(), current_stream_ptr()),
                   "sapt_train_rows_plan")
        return (src, dest, zero_rows, scal[0], scal[1]) if return_owned else (src, dest, zero_rows, scal[0])
    cap = train_mine.shape[1] if train_mine is not None else None
    lo = 1 if skip_first else 0
    zero = 1 if zero_row else 0
    cnt = train_count.to(torch.int64)
    if cap is None:
        raise ValueError("training_row_plan needs the [R_global, cap] ownership mask (use an all-true mask on one rank)")
    cnt = torch.clamp(cnt, max=cap)
    n_valid = torch.clamp(cnt - lo, min=0)
    per_chain = n_valid + zero
    start = torch.cumsum(per_chain, 0) - per_chain                      # first batch row of every chain
    n_out = per_chain.sum()
    j = torch.arange(cap, device=dev, dtype=torch.int64)[None, :]
    own = train_mine & (j >= lo) & (j < cnt[:, None])                   # [Rg, cap]
    own_i = own.to(torch.int64)
    owned_per_slot = own_i.sum(dim=1)
    local_start = torch.cumsum(owned_per_slot, 0) - owned_per_slot
    rank_in_slot = torch.cumsum(own_i, dim=1) - 1
    pos = torch.where(own, local_start[:, None] + rank_in_slot, torch.full_like(rank_in_slot, n_pad))   # non-owned -> dummy slot n_pad
    pos = torch.clamp(pos, max=n_pad)                                    # more owned rows than n_pad cannot happen (see engine); never write out of range
    flat = (torch.arange(Rg, device=dev, dtype=torch.int64)[:, None] * cap + j).reshape(-1)
    src = torch.full((n_pad + 1,), -1, dtype=torch.int64, device=dev)
    src.scatter_(0, pos.reshape(-1), flat)
    dst_val = (start[:, None] + zero + (j - lo)).expand(Rg, cap).reshape(-1)
    dest = torch.full((n_pad + 1,), -1, dtype=torch.int64, device=dev)
    dest.scatter_(0, pos.reshape(-1), dst_val)
    src, dest = src[:n_pad], dest[:n_pad]
    dest = torch.where(src >= 0, dest, torch.full_like(dest, -1))
    zero_rows = start if zero_row else start[:0]
    out = (src.to(torch.int32).contiguous(), dest.to(torch.int32).contiguous(), zero_rows.contiguous(), n_out)
    return out + ((src >= 0).sum(),) if return_owned else out


def gather_payloads(payload: torch.Tensor, trainer_rank: int, world: int, group=None, n_owned: torch.Tensor = None, round_rows: int = 256):
    """Payloads of all ranks -> one [world * m, W] tensor on the trainer rank (None elsewhere). The rows that carry data are packed at
    the front of a payload (training_row_plan); with `n_owned` (device scalar: how many) only the first m = max over ranks of n_owned,
    rounded up to `round_rows`, travel — one small all-reduce and ONE host read per rendezvous instead of moving the padding (8 x 8192 chains,
    surrogate_prob 0.5: 40 % of a 106 MB payload per rank). Without it m = all rows and the host never waits for the device."""
    import torch.distributed as dist
    if world == 1:
        return payload
    m = payload.shape[0]
    if n_owned is not None:
        t = n_owned.to(torch.int64).reshape(1).clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        m = min(m, ((int(t.item()) + round_rows - 1) // round_rows) * round_rows)
    part = payload[:m]
    rank = dist.get_rank(group)
    if m == 0:                      # every rank knows (m comes from the all-reduce): nothing to move
        return part if rank == trainer_rank else None
    if rank == trainer_rank:
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.gather(part, gather_list=parts, dst=trainer_rank, group=group)
        return torch.cat(parts, dim=0)
    dist.gather(part, gather_list=None, dst=trainer_rank, group=group)
    return None


SURROGATE_WEIGHT_KEYS = ("W1", "b1", "W2", "b2", "W3", "b3")


def _sizes(shapes):
    out = []
    for sh in shapes:
        n = 1
        for v in sh:
            n *= int(v)
        out.append(n)
    return out


def split_flat(flat: torch.Tensor, shapes) -> dict:
    out, o = {}, 0
    for k, n in zip(SURROGATE_WEIGHT_KEYS, _sizes(shapes)):
        out[k] = flat[o:o + n]
        o += n
    return out


def broadcast_surrogate_device(weights, shapes, min_y: float, max_y: float, trainer_rank: int, device, group=None):
    """Trainer rank -> everyone, as DEVICE tensors (the caller decides whether the host has to look): head = [has_weights,
    min_y, max_y] (fp64) and the flattened fp32 weights in the order W1, b1, W2, b2, W3, b3 (P h1 + h1 + h1 h2 + h2 + h2 + 1
    values: <= 3.2 MB at P = 3975 with topology 3). `weights` is the trainer's dict (or None) on the trainer rank and
    ignored elsewhere."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    head = torch.zeros(3, dtype=torch.float64, device=device)
    flat = torch.zeros(sum(_sizes(shapes)), dtype=torch.float32, device=device)
    if rank == trainer_rank:
        head[1], head[2] = float(min_y), float(max_y)
        if weights is not None:
            head[0] = 1.0
            flat.copy_(torch.cat([torch.as_tensor(weights[k]).to(torch.float32).reshape(-1).to(device) for k in SURROGATE_WEIGHT_KEYS]))
    dist.broadcast(head, src=trainer_rank, group=group)
    dist.broadcast(flat, src=trainer_rank, group=group)
    return head, flat


def sync_history_rows(hist_rows: torch.Tensor, group=None) -> None:
    """Rows of a per-chain history written since the last synchronisation: between two exchanges the binding is
    fixed, so every (row, chain) entry is non-zero on exactly one rank and a plain sum completes the rows everywhere
    (NaN markers survive: NaN + 0 = NaN)."""
    import torch.distributed as dist
    dist.all_reduce(hist_rows, op=dist.ReduceOp.SUM, group=group)
