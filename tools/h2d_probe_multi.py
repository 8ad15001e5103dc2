#!/usr/bin/env python3
"""Aggregate pinned host->device bandwidth of N concurrent ranks (one per GPU), to find what limits the end-to-end arm at
N = 8 (e2e efficiency 0.80 while the resident arm scales at 0.99): every rank copies the bench's per-step
input volume (images + flattened map, 224 MB) in a loop while all other ranks do the same.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe_multi.py

Modes, run one after the other:
  default   pinned buffers allocated wherever the launcher put the process
  spread    the process binds itself to its own slice of the host cores BEFORE allocating and first-touching its pinned
            buffer (first-touch NUMA placement: the pages land on the node of that slice)
  two_bufs  like spread, two buffers on two streams per rank (two DMA queues)
Prints one JSON line on rank 0: per-rank and aggregate GB/s per mode + the host's NUMA layout as Linux reports it."""
import json
import os
import time

import torch
import torch.distributed as dist


def numa_layout():
    out = {}
    base = "/sys/devices/system/node"
    try:
        for n in sorted(os.listdir(base)):
            if n.startswith("node"):
                out[n] = open(os.path.join(base, n, "cpulist")).read().strip()
    except Exception:
        pass
    return out


def run_mode(nbytes, reps, two):
    dev = torch.cuda.current_device()
    bufs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2 if two else 1)]
    for b in bufs:
        b.fill_(1)                               # first touch by this (possibly re-bound) thread
    dst = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in bufs]
    streams = [torch.cuda.Stream() for _ in bufs]
    for _ in range(2):
        for b, d, s in zip(bufs, dst, streams):
            with torch.cuda.stream(s):
                d.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        for b, d, s in zip(bufs, dst, streams):
            with torch.cuda.stream(s):
                d.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = len(bufs) * nbytes * reps / dt / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(allr, t)
    dist.barrier()
    del bufs, dst
    return [float(x) for x in allr], dev


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes, reps = 224 << 20, 30
    res = {"world": world, "bytes_per_copy": nbytes, "numa": numa_layout(), "cores": os.cpu_count(),
           "affinity_at_start": sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))]}
    per, _ = run_mode(nbytes, reps, False)
    res["default"] = {"per_rank_GBs": per, "aggregate_GBs": sum(per)}
    cores = sorted(os.sched_getaffinity(0))
    share = max(1, len(cores) // world)
    mine = cores[rank * share:(rank + 1) * share] or cores
    os.sched_setaffinity(0, mine)
    per, _ = run_mode(nbytes, reps, False)
    res["spread"] = {"per_rank_GBs": per, "aggregate_GBs": sum(per), "cores_per_rank": share}
    per, _ = run_mode(nbytes // 2, reps, True)
    res["two_bufs"] = {"per_rank_GBs": per, "aggregate_GBs": sum(per)}
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
