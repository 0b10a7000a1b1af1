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


def sync_training_buffers(train_theta: torch.Tensor, train_L: torch.Tensor, train_count: torch.Tensor, mask: torch.Tensor, group=None) -> None:
    """Surrogate training rows are appended on whichever rank hosts a chain. In place: every chain's buffer
    ([R_global, cap, P] rows, [R_global, cap] likelihoods, [R_global] counts) is taken from its CURRENT host, so
    afterwards the buffers are complete and identical on all ranks. Must run before an exchange re-binds the chains
    (the rows of the interval were written under the old binding) and before the rows go to the trainer."""
    import torch.distributed as dist
    cnt = train_count.to(torch.float64) * mask
    train_theta.mul_(mask[:, None, None].to(train_theta.dtype))
    train_L.mul_(mask[:, None].to(train_L.dtype))
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(train_theta, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(train_L, op=dist.ReduceOp.SUM, group=group)
    train_count.copy_(cnt.round().to(train_count.dtype))


def sync_history_rows(hist_rows: torch.Tensor, group=None) -> None:
    """Rows of a per-chain history written since the last synchronisation: between two exchanges the binding is
    fixed, so every (row, chain) entry is non-zero on exactly one rank and a plain sum completes the rows everywhere
    (NaN markers survive: NaN + 0 = NaN)."""
    import torch.distributed as dist
    dist.all_reduce(hist_rows, op=dist.ReduceOp.SUM, group=group)
