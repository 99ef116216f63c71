"""CPU, world_size 2 over gloo: the N>1 host logic — shard plans tile the stream and assemble_on() rebuilds it."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import ROOT  # noqa: E402


def _worker(rank, world, port, stream_bytes, decoded_bytes, ok):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    pkg = entry.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stream = np.frombuffer(stream_bytes, np.uint8)
    decoded = np.frombuffer(decoded_bytes, np.uint8)
    blocks = pkg.mt_index(64, stream)
    plans = pkg.plan_shards(blocks, world)
    mine = plans[rank]
    # this rank "decodes" exactly its own output range (the checker's bytes stand in for the GPU here)
    local = torch.from_numpy(decoded[mine.out_offset: mine.out_offset + mine.out_bytes].copy())
    full = pkg.assemble_on(0, local, plans, decoded.size)
    good = True
    if rank == 0:
        good = full is not None and np.array_equal(full.numpy(), decoded)
    else:
        good = full is None
    # the compressed ranges each rank needs are contiguous, ordered and cover every block once
    good &= sum(p.last_unit - p.first_unit for p in plans) == len(blocks)
    good &= all(a.last_unit == b.first_unit for a, b in zip(plans, plans[1:]))
    good &= sum(p.out_bytes for p in plans) == decoded.size
    ok[rank] = bool(good)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_assemble(golden):
    stream = golden["stream/multi/2/64/15"]
    decoded = golden["in/multi"]
    world = 2
    ctx = mp.get_context("spawn")
    ok = ctx.Array("b", [0] * world)
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, stream.tobytes(), decoded.tobytes(), ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1, 1]


def test_plans_for_more_ranks_than_blocks(pkg, golden):
    stream = golden["stream/small/2/64/12"]  # a single block
    blocks = pkg.mt_index(64, stream)
    plans = pkg.plan_shards(blocks, 4)
    assert sum(p.out_bytes for p in plans) == golden["in/small"].size
    assert sum(1 for p in plans if p.out_bytes) == 1
