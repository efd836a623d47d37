#!/usr/bin/env python
"""tools/pcie_bw.py -- what the HOST side of the end-to-end path can deliver: concurrent pinned-memory H2D + D2H copy bandwidth per GPU
when 1, 2, 4, 8 GPUs of the box copy at the same time (e2e scaling).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_bw.py [--mb 1024] [--secs 2]

Every rank binds to the CPUs of its GPU's NUMA node, first-touches its pinned buffers there, and then runs H2D and D2H copies on two
streams for `secs` seconds (the same traffic pattern as sb200_*_host: 2 parts in, 1 part out).  Rank 0 prints one JSON line with the
per-rank rates and the aggregate; run it for N = 1, 2, 4, 8 and compare the per-GPU rate with the N = 1 one: a drop that follows the
number of GPUs per socket names the root complex / host DRAM as the limit of bench.py's e2e figure, not the GPU path."""
import argparse
import json
import os
import time


def numa_of(dev):
    import torch

    try:
        p = torch.cuda.get_device_properties(dev)
        addr = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{addr}/numa_node").read())
        cpus = set()
        if node >= 0:
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        return node, cpus & os.sched_getaffinity(0)
    except Exception:
        return -1, set()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024, help="MiB per H2D buffer (D2H buffer is half of it)")
    ap.add_argument("--secs", type=float, default=2.0)
    ap.add_argument("--no-bind", action="store_true", help="do not bind to the GPU's NUMA node (shows the cross-socket penalty)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    node, cpus = numa_of(local)
    if cpus and not args.no_bind:
        os.sched_setaffinity(0, cpus)
    n_in = args.mb * (1 << 20) // 8
    h_in = torch.empty(n_in, dtype=torch.int64).pin_memory()
    h_out = torch.empty(n_in // 2, dtype=torch.int64).pin_memory()
    h_in.fill_(1)
    h_out.fill_(2)  # first touch on this rank's node
    d_in = torch.empty(n_in, dtype=torch.int64, device="cuda")
    d_out = torch.ones(n_in // 2, dtype=torch.int64, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def burst(k):
        for _ in range(k):
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)

    burst(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, reps = time.perf_counter(), 0
    while time.perf_counter() - t0 < args.secs:
        burst(4)
        torch.cuda.synchronize()
        reps += 4
    dt = time.perf_counter() - t0
    mine = torch.tensor([float(node), h_in.nbytes * reps / dt / 1e9, h_out.nbytes * reps / dt / 1e9], dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        per = [{"rank": r, "gpu_numa_node": int(x[0].item()), "h2d_GBps": round(x[1].item(), 2), "d2h_GBps": round(x[2].item(), 2)} for r, x in enumerate(allr)]
        print(json.dumps({"tool": "pcie_bw", "n_gpus": world, "bound_to_gpu_numa_node": not args.no_bind, "mb_per_h2d_copy": args.mb,
                          "per_rank": per, "h2d_GBps_total": round(sum(p["h2d_GBps"] for p in per), 1),
                          "d2h_GBps_total": round(sum(p["d2h_GBps"] for p in per), 1),
                          "h2d_GBps_per_gpu_min": min(p["h2d_GBps"] for p in per), "d2h_GBps_per_gpu_min": min(p["d2h_GBps"] for p in per)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
