#!/usr/bin/env python
"""Device-to-host copy ceiling with 1 / 2 / 4 / 8 ranks copying AT THE SAME TIME into pinned host memory -- the bound of the
fp32 host-buffer path (bench.py `e2e`: 1.88 GB of sensor images per step and GPU).

    python tools/pcie_bw.py                                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/pcie_bw.py

Every rank copies `--mb` MiB device -> pinned host `--reps` times between two barriers (CUDA events per rank, wall clock
over all ranks); rank 0 prints one JSON line: per-rank GB/s, the aggregate, and what that means for one bench step."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1792)           # 16384 envs x 7 scans x 4096 pixels x 4 B = 1792 MiB
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.mb << 20
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    res = {}
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        mine = n * a.reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([mine, wall], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(allr, t)
        else:
            allr = [t]
        per_rank = [float(x[0]) for x in allr]
        wall_max = max(float(x[1]) for x in allr)
        res[name] = {"per_rank_gbs": per_rank, "aggregate_gbs": world * n * a.reps / wall_max / 1e9,
                     "ms_per_bench_step_fp32": (1792 << 20) / (min(per_rank) * 1e9) * 1e3}
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"ranks": world, "mib_per_copy": a.mb, "reps": a.reps, **res,
                          "host": {"cpus": os.cpu_count()}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
