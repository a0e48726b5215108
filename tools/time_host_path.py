"""Tuning aid: wall-clock cost of the Python host path per env.step (dict API, tensor API, rollout) vs kernel time."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from swarmacb_isaaclab_b200.env import SwarmEnv  # noqa: E402


def wall(fn, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    return t_issue / n * 1e6, (time.perf_counter() - t0) / n * 1e6


def main():
    dev = "cuda:0"
    for name, E in (("homing_lily_4096", 256), ("homing_lily_4096", 4096), ("foraging_daisy_16384", 16384)):
        mission, mode, _, task, _ = bench.WORKLOADS[name]
        env = SwarmEnv(bench.make_cfg(mission, mode, E, dev))
        env.reset(seed=0)
        act = bench.gen_actions(torch, True, 1, E, dev)[0]
        adict = {a: act[:, i] for i, a in enumerate(env.possible_agents)}
        for _ in range(20):
            env.step(adict)
        i1, w1 = wall(lambda: env.step_tensor(act), 300)
        i2, w2 = wall(lambda: env.step(adict), 300)
        i3, w3 = wall(lambda: env.rollout(act, 5), 100)
        print(f"{name} E={E}: step_tensor issue {i1:.1f} us / wall {w1:.1f} us; step(dict) issue {i2:.1f} / wall {w2:.1f}; "
              f"rollout(5) issue {i3:.1f} / wall {w3:.1f}", flush=True)


if __name__ == "__main__":
    main()
