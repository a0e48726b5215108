"""Tuning aid: fused decision-period rollout (swarm_rollout, one launch) vs the same number of single env.steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from swarmacb_isaaclab_b200.env import SwarmEnv  # noqa: E402


def timed(fn, flush):
    flush.add_(1.0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    dev = "cuda:0"
    flush = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device=dev)
    for name in sys.argv[2:] or list(bench.WORKLOADS):
        mission, mode, E, task, _ = bench.WORKLOADS[name]
        # twin environments on the same seed: A takes T single steps, B one rollout, per decision (same states)
        A, B = SwarmEnv(bench.make_cfg(mission, mode, E, dev)), SwarmEnv(bench.make_cfg(mission, mode, E, dev))
        A.reset(seed=0)
        B.reset(seed=0)
        acts = bench.gen_actions(torch, bool(A.params.discrete_actions), 40, E, dev)
        single, fused = [], []
        for d in range(40):
            act = acts[d]
            single.append(timed(lambda: [A.step_tensor(act) for _ in range(T)], flush))
            fused.append(timed(lambda: B.rollout(act, T), flush))
        assert torch.equal(A.agent_pos, B.agent_pos)
        single, fused = sorted(single[8:]), sorted(fused[8:])
        s, f = single[len(single) // 2], fused[len(fused) // 2]
        print(f"{name:26s} T={T}: {T} launches {s*1e3:8.1f} us  rollout {f*1e3:8.1f} us  "
              f"x{s/f:.2f}  {E*20*T/f/1e6:.2f} G agent-steps/s", flush=True)


if __name__ == "__main__":
    main()
