"""Tuning aid: ms per env.step of one workload over a list of env counts (wave-quantisation / tail study).

    python tools/sweep_envs.py foraging_daisy_16384 4736 9472 14208 16384 18944 65536
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from swarmacb_isaaclab_b200.env import SwarmEnv  # noqa: E402


def main():
    name = sys.argv[1]
    sizes = [int(a) for a in sys.argv[2:]]
    mission, mode, _, task, _ = bench.WORKLOADS[name]
    dev = "cuda:0"
    flush = None if os.environ.get("SWEEP_WARM_L2") else torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device=dev)
    for E in sizes:
        env = SwarmEnv(bench.make_cfg(mission, mode, E, dev))
        env.reset(seed=0)
        acts = bench.gen_actions(torch, bool(env.params.discrete_actions), 32, E, dev)
        ms = bench.time_steps(torch, env, acts, 100, 30, flush)
        ms.sort()
        med = ms[len(ms) // 2]
        print(f"{name} E={E:6d} median {med*1e3:8.2f} us  min {ms[0]*1e3:8.2f} us  "
              f"{E*20/med/1e6:8.1f} M agent-steps/s  {med*1e3/E*1e3:7.3f} ns/env", flush=True)
        del env, acts


if __name__ == "__main__":
    main()
