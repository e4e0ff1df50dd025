"""Scenario groups: us per environment step of G groups x R/G replicas.  usage: grouped_probe.py R G"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pednstream_b200.rl import GroupedPedNetEnv

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
env = GroupedPedNetEnv("45_intersections", replicas=R, groups=G, obs_mode="option3", seed=1000, device="cuda:0",
                       randomize="device")
torch.manual_seed(0)
a = torch.rand((8, R, env.n_act), device="cuda") * 4.0
for k in range(10):
    env.step(a[k & 7])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(40):
    env.step(a[k & 7])
e1.record()
torch.cuda.synchronize()
env.check_errors()
us = 1e3 * e0.elapsed_time(e1) / 40
print(json.dumps({"replicas": R, "groups": G, "us_per_env_step": us, "env_steps_per_s": R / us * 1e6}))
