"""Platform probe (run under torchrun, one rank per GPU): plain pinned-memory cudaMemcpyAsync rates with ALL ranks
copying at once - the floor under bench.py's end-to-end leg, whose per-step traffic is 31.5 MB of observations down
and 2.6 MB of actions up per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n_down, n_up = 31_539_200, 2_621_440
    d_obs = torch.empty(n_down, dtype=torch.uint8, device="cuda")
    h_obs = torch.empty(n_down, dtype=torch.uint8).pin_memory()
    d_act = torch.empty(n_up, dtype=torch.uint8, device="cuda")
    h_act = torch.empty(n_up, dtype=torch.uint8).pin_memory()
    out = {}
    for name, fn, nbytes in (("d2h_obs", lambda: h_obs.copy_(d_obs, non_blocking=True), n_down),
                             ("h2d_actions", lambda: d_act.copy_(h_act, non_blocking=True), n_up),
                             ("both", lambda: (d_act.copy_(h_act, non_blocking=True), h_obs.copy_(d_obs, non_blocking=True)),
                              n_down + n_up)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        reps = 100
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[name] = {"GB_per_s": nbytes * reps / dt / 1e9, "ms_per_copy": dt / reps * 1e3}
        if world > 1:
            dist.barrier()
    res = [None] * world
    if world > 1:
        dist.all_gather_object(res, out)
    else:
        res = [out]
    if rank == 0:
        summary = {"world": world, "bytes_down": n_down, "bytes_up": n_up,
                   "per_rank_d2h_GBps": [round(r["d2h_obs"]["GB_per_s"], 1) for r in res],
                   "per_rank_h2d_GBps": [round(r["h2d_actions"]["GB_per_s"], 1) for r in res],
                   "per_rank_both_ms": [round(r["both"]["ms_per_copy"], 3) for r in res],
                   "sum_d2h_GBps": round(sum(r["d2h_obs"]["GB_per_s"] for r in res), 1),
                   "host": {"cpus": os.cpu_count()}}
        print(json.dumps(summary))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
