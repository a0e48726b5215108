import sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
from swarmacb_isaaclab_b200.env import SwarmEnv
E = int(sys.argv[1])
env = SwarmEnv(bench.make_cfg("hom", "lily", E, "cuda:0"))
env.reset(seed=0)
acts = bench.gen_actions(torch, True, 8, E, "cuda:0")
for t in range(12):
    env.step_tensor(acts[t % 8])
torch.cuda.synchronize()
