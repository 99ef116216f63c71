"""N-way host<->device copy ceiling of this box: every rank copies 852 MB host->device and 1 GB device->host AT THE SAME
TIME (pinned memory, two streams), all ranks at once — the traffic pattern of bench.py's e2e leg at N GPUs.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/pcie_probe_nway.py
Prints one JSON line per mode from rank 0: aggregate GB/s over both directions, per direction, and the slowest rank.
Modes: plain pinned buffers; buffers first-touched and pinned while the process is bound to the CPUs NVML reports as
local to the rank's GPU (NUMA placement).
"""
import json, os, time
import torch, torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
IN_B, OUT_B = 852_000_048, 1_000_000_000


def bind_local_cpus():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [64 * i + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception as e:  # noqa: BLE001
        return str(e)


def measure(mode):
    cpus = None
    if mode == "numa_local":
        cpus = bind_local_cpus()
    h_in = torch.empty(IN_B, dtype=torch.uint8)
    h_out = torch.empty(OUT_B, dtype=torch.uint8)
    h_in.fill_(1); h_out.fill_(2)  # first touch under the current CPU binding
    h_in = h_in.pin_memory(); h_out = h_out.pin_memory()
    d_in = torch.empty(IN_B, dtype=torch.uint8, device="cuda"); d_out = torch.empty(OUT_B, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def step(h2d=True, d2h=True):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)

    res = {}
    for label, kw in (("h2d_only", dict(d2h=False)), ("d2h_only", dict(h2d=False)), ("both", {})):
        step(**kw); torch.cuda.synchronize()
        times = []
        for _ in range(5):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            step(**kw)
            torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
        mine = min(times)
        if world > 1:
            t = torch.tensor([mine], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            slow = float(t.item())
        else:
            slow = mine
        nbytes = (IN_B if kw.get("h2d", True) else 0) + (OUT_B if kw.get("d2h", True) else 0)
        res[label] = {"slowest_rank_ms": round(slow * 1e3, 2), "aggregate_GBps": round(world * nbytes / slow / 1e9, 1),
                      "per_rank_GBps": round(nbytes / slow / 1e9, 1)}
    res["both"]["decoded_GBps_ceiling_if_each_rank_moved_a_whole_stream"] = round(world * OUT_B / (res["both"]["slowest_rank_ms"] * 1e-3) / 1e9, 1)
    if rank == 0:
        print(json.dumps({"mode": mode, "ranks": world, "h2d_bytes": IN_B, "d2h_bytes": OUT_B, "cpus_rank0": cpus if not isinstance(cpus, list) else f"{len(cpus)} cpus", **res}), flush=True)
    del h_in, h_out, d_in, d_out


for mode in ("plain", "numa_local"):
    measure(mode)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
