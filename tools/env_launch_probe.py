"""A few environment steps for a launch-list capture (experiment)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pednstream_b200.rl import BatchedPedNetEnv
env = BatchedPedNetEnv("45_intersections", replicas=1024, obs_mode="option3", seed=1, device="cuda:0")
a = torch.rand((1024, env.n_act), device=env.device) * 4.0
for _ in range(330):
    env.step(a)
torch.cuda.synchronize()
