"""Measurement aid: where does a launch's time go between its blocks?  Needs the -DSWARM_BLOCK_TIMES variant:

    python tools/build_variants.py times=SWARM_BLOCK_TIMES
    SWARM_LIB_OVERRIDE=$PWD/swarmacb-isaaclab_b200/variants/lib_times.so python tools/block_times.py homing_lily_4096 [warm-up steps]

Every block records %globaltimer at entry and when its last warp leaves, and its SM.  Printed: when blocks start
(the launch ramp), how long they live, when each SM goes idle, and how unevenly the work is spread - the numbers
behind "an SM idles x % of the launch" in profiles/README.md.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from swarmacb_isaaclab_b200 import _lib  # noqa: E402
from swarmacb_isaaclab_b200.env import SwarmEnv  # noqa: E402


def pct(a, qs=(0, 10, 50, 90, 99, 100)):
    return "  ".join(f"p{q}={np.percentile(a, q) / 1e3:6.2f}" for q in qs)


def main():
    name = sys.argv[1]
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    mission, mode, E, _, _ = bench.WORKLOADS[name]
    dev = "cuda:0"
    lib = _lib.load()
    fn = lib.swarm_debug_block_times
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    env = SwarmEnv(bench.make_cfg(mission, mode, E, dev))
    env.reset(seed=0)
    acts = bench.gen_actions(torch, bool(env.params.discrete_actions), 32, E, dev)
    flush = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device=dev)
    n_blocks = (E + 7) // 8
    for i in range(warm):
        env.step_tensor(acts[i % 32])
    rows = []
    for rep in range(5):
        flush.add_(1.0)
        torch.cuda.synchronize()
        env.step_tensor(acts[(warm + rep) % 32])
        torch.cuda.synchronize()
        buf = np.zeros(6 * n_blocks, np.uint64)
        assert fn(buf.ctypes.data, n_blocks) == 0
        t = buf.reshape(n_blocks, 6).astype(np.int64)
        t0 = t[:, 0].min()
        start, end, sm = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] & 0xFFFF
        rounds, rebuilds, pair_passes = (t[:, 2] >> 32) & 0xFFFF, (t[:, 2] >> 48) & 0xFF, (t[:, 2] >> 56) & 0xFF
        dur = end - start
        total = end.max()
        sm_end = np.array([end[sm == s].max() for s in np.unique(sm)])
        sm_start = np.array([start[sm == s].min() for s in np.unique(sm)])
        per_sm = np.bincount(sm)
        per_sm = per_sm[per_sm > 0]
        # busy time of an SM = union of its blocks' residence intervals
        busy = []
        for s in np.unique(sm):
            iv = sorted(zip(start[sm == s], end[sm == s]))
            b, cur_a, cur_b = 0, iv[0][0], iv[0][1]
            for a, e in iv[1:]:
                if a > cur_b:
                    b += cur_b - cur_a
                    cur_a, cur_b = a, e
                else:
                    cur_b = max(cur_b, e)
            busy.append(b + cur_b - cur_a)
        busy = np.array(busy)
        rows.append((total, busy.mean() / total))
        if rep == 4:
            print(f"{name}: {n_blocks} blocks on {len(per_sm)} SMs ({per_sm.min()}..{per_sm.max()} per SM), "
                  f"first entry -> last exit {total / 1e3:.2f} us  (globaltimer resolution "
                  f"{np.diff(np.unique(t[:, :2])).min()} ns)")
            print(f"  block entry   (us after the first): {pct(start)}")
            print(f"  block exit                        : {pct(end)}")
            print(f"  block residence                   : {pct(dur)}")
            staged, pre_sense, post_sense = t[:, 3] - t[:, 0], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4]
            print(f"  phase: state loads + staging      : {pct(staged)}")
            print(f"  phase: decode, motion, collisions : {pct(pre_sense)}")
            print(f"  phase: sensor suite (warp 0)      : {pct(post_sense)}")
            print(f"  phase: stores + block's last warp : {pct(t[:, 1] - t[:, 5])}")
            slow = np.argsort(dur)[-max(1, n_blocks // 20):]
            print(f"  slowest 5 % of the blocks         : collisions {pre_sense[slow].mean() / 1e3:.2f} us (all {pre_sense.mean() / 1e3:.2f}), "
                  f"sensors {post_sense[slow].mean() / 1e3:.2f} us (all {post_sense.mean() / 1e3:.2f})")
            print("  solver rounds -> blocks, mean collision-phase us, mean rebuilds, mean pair passes:")
            for r in np.unique(rounds):
                m = rounds == r
                print(f"    {int(r)} rounds: {int(m.sum()):5d} blocks  {pre_sense[m].mean() / 1e3:6.2f} us  "
                      f"{rebuilds[m].mean():.2f} rebuilds  {pair_passes[m].mean():.2f} pair passes")
            print(f"  SM first entry                    : {pct(sm_start)}")
            print(f"  SM last exit                      : {pct(sm_end)}")
            print(f"  SM busy share of the launch       : mean {busy.mean() / total:.3f}  min {busy.min() / total:.3f}")
            by_count = {int(c): float(np.mean([e for e, k in zip(sm_end, per_sm) if k == c])) / 1e3 for c in np.unique(per_sm)}
            print(f"  mean SM last exit by blocks per SM: {by_count}")
            order = np.argsort(start)
            second = start[order][len(per_sm) * int(per_sm.max()):] if n_blocks > 148 * 7 else np.array([])
            if second.size:
                print(f"  blocks entering after the first wave: {second.size}, entry {pct(second)}")
    print("  launches: " + "  ".join(f"{a / 1e3:.2f} us / busy {b:.3f}" for a, b in rows))


if __name__ == "__main__":
    main()
