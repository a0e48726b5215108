import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import fixtures
from oracle import oracle
from swarmacb_isaaclab_b200.env import SwarmEnv
from swarmacb_isaaclab_b200.params import N
sys.path.insert(0, "tests")
from test_gpu_oracle import _cluster
for mission, mode in [("hom", "lily"), ("shl", "daisy"), ("dgt", "dandelion"), ("for", "daisy"), ("xor", "oc2")]:
    E = 48
    cfg = fixtures.make_cfg(mission, mode, E, device="cuda:0")
    env = SwarmEnv(cfg); p = env.params
    rng = np.random.default_rng(7)
    host = oracle.new_state(E)
    noise = dict(rab_u=rng.random((E, N, N), dtype=np.float32), spawn_u=_cluster(rng, E), yaw_u=rng.random((E, N), dtype=np.float32))
    env.inject_noise(**noise); env.reset(); oracle.reset(p, host, **noise)
    torch.cuda.synchronize()
    dev = env.dump_state()
    d = np.abs(dev["pos"] - host["pos"])
    bad = np.argwhere(d > 0)
    print(mission, mode, "reset pos mismatches:", len(bad), "max", d.max(), "yaw eq", np.array_equal(dev["yaw"], host["yaw"]))
    if len(bad):
        e, i, c = bad[0]
        print("  first", bad[0], dev["pos"][e, i], host["pos"][e, i], "envs:", sorted(set(bad[:, 0].tolist()))[:10])
    # one step
    act = rng.integers(0, 6, (E, N), dtype=np.int64) if p.discrete_actions else (rng.random((E, N, 2), dtype=np.float32) * 2 - 1).astype(np.float32)
    noise = dict(rab_u=rng.random((E, N, N), dtype=np.float32), turn_dur=rng.integers(1, 5, (E, N, 3)).astype(np.int32))
    for k in ("cached_left", "beh_cache", "fsm", "prev_ground"):
        print("   ", k, "equal after reset:", np.array_equal(dev[k], host[k]), np.abs(dev[k].astype(np.float64) - host[k]).max())
    env.inject_noise(**noise)
    env.step_tensor(torch.as_tensor(act, device="cuda:0")); oracle.step(p, host, act, **noise)
    torch.cuda.synchronize(); dev = env.dump_state()
    d = np.abs(dev["pos"] - host["pos"]); print("   after step: pos mismatches", int((d > 0).sum()), d.max(), "wheels eq", np.array_equal(dev["cached_left"], host["cached_left"]))
