"""Host logic for the one-process-per-GPU mt_ path: which rank decodes what, and how decoded shards are
assembled on one device.

mt_ streams are chains of independently decodable blocks (/root/reference/src/mt_rANS32x64_16w_decode.cpp:62-66:
every block header snapshots all N states), so ranks take contiguous block ranges and never exchange data while
decoding. torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used ONLY to assemble the decoded buffer on
one rank when a caller asks for it; raw and block_ streams are a single recurrence and do not shard at all.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

from . import capi


@dataclass(frozen=True)
class ShardPlan:
    rank: int
    world: int
    first_unit: int
    last_unit: int      # exclusive
    out_offset: int     # first decoded byte owned by this rank
    out_bytes: int
    in_offset: int      # first compressed byte this rank needs (16-byte aligned down)
    in_bytes: int


def plan_shards(blocks: Sequence[capi.Block], world: int) -> List[ShardPlan]:
    """Contiguous block ranges balanced on compressed + decoded bytes (hsr_mt_partition)."""
    first = capi.mt_partition(list(blocks), world)
    plans = []
    for r in range(world):
        a, b = first[r], first[r + 1]
        if a == b:
            plans.append(ShardPlan(r, world, a, b, 0, 0, 0, 0))
            continue
        out_lo = blocks[a].outOffset
        out_hi = blocks[b - 1].outOffset + blocks[b - 1].count
        in_lo = blocks[a].inOffset & ~15
        in_hi = blocks[b - 1].inEnd
        plans.append(ShardPlan(r, world, a, b, out_lo, out_hi - out_lo, in_lo, in_hi - in_lo))
    return plans


def assemble_on(dst: int, local, plans: Sequence[ShardPlan], total_bytes: int, group=None):
    """Gathers every rank's decoded shard (a 1-D uint8 tensor of plans[rank].out_bytes) into one tensor of
    total_bytes on rank `dst`. Variable-size point-to-point sends over NCCL/NVLink (or gloo); returns the
    assembled tensor on dst and None elsewhere."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    if rank != dst:
        if plans[rank].out_bytes:
            op = dist.P2POp(dist.isend, local[: plans[rank].out_bytes].contiguous(), dst, group)
            for req in dist.batch_isend_irecv([op]):
                req.wait()
        return None
    full = torch.empty(total_bytes, dtype=torch.uint8, device=local.device)
    ops = []  # every receive is posted at once (one NCCL group): the shards arrive side by side over NVLink
    for p in plans:
        if p.out_bytes == 0:
            continue
        view = full[p.out_offset: p.out_offset + p.out_bytes]
        if p.rank == dst:
            view.copy_(local[: p.out_bytes])
        else:
            ops.append(dist.P2POp(dist.irecv, view, p.rank, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return full
