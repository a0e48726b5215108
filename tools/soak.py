"""One-off soak: long free-running rollouts (several episode roll-overs) run twice from the same seed must end in
bit-identical state (a race in the block-wide exchanges would show up as nondeterminism), stay finite, inside the
arena and without deep overlaps.   python tools/soak.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from swarmacb_isaaclab_b200.env import SwarmEnv  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N = 20
for name in bench.WORKLOADS:
    mission, mode, E, task, _ = bench.WORKLOADS[name]
    finals = []
    for rep in range(2):
        env = SwarmEnv(bench.make_cfg(mission, mode, E, "cuda:0"))
        env.reset(seed=11)
        g = torch.Generator(device="cuda:0").manual_seed(3)
        discrete = bool(env.params.discrete_actions)
        total = torch.zeros(E, device="cuda:0", dtype=torch.float64)
        resets = 0
        for t in range(T):
            if t % 5 == 0:
                act = (torch.randint(0, 6, (E, N, 1), generator=g, device="cuda:0") if discrete
                       else torch.rand(E, N, 2, generator=g, device="cuda:0") * 2 - 1)
            obs, rew, to = env.step_tensor(act)
            total += rew
            if t % 500 == 499:
                resets += int(to.sum())
        torch.cuda.synchronize()
        st = env.dump_state()
        finals.append((st, obs.cpu().numpy().copy(), total.cpu().numpy()))
    a, b = finals
    same = all(np.array_equal(a[0][k], b[0][k]) for k in a[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    pos = a[0]["pos"]
    d = np.linalg.norm(pos[:, :, None] - pos[:, None], axis=-1) + np.eye(N) * 10
    print(f"{name:26s} steps={T} identical={same} finite={np.isfinite(pos).all() and np.isfinite(a[1]).all()} "
          f"max_r={np.linalg.norm(pos, axis=-1).max():.4f} min_pair={d.min():.4f} "
          f"episodes={int(a[0]['episode_length_buf'].max())} mean_return/step={a[2].mean() / T:.4f}", flush=True)
    assert same
