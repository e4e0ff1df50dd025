"""Batched-environment probe: ms per env step in windows of an episode (CUDA events), for tuning runs and
launch-list / ncu captures.  usage: env_probe.py REPLICAS [windows "a-b,c-d"] [dataset]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pednstream_b200.rl import BatchedPedNetEnv

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
windows = sys.argv[2] if len(sys.argv) > 2 else "10-60,300-350,600-650"
dataset = sys.argv[3] if len(sys.argv) > 3 else "45_intersections"
wins = [tuple(int(x) for x in w.split("-")) for w in windows.split(",")]
env = BatchedPedNetEnv(dataset, replicas=R, obs_mode="option3", seed=1, device="cuda:0")
gen = torch.Generator(device=env.device)
gen.manual_seed(0)
a = torch.rand((8, R, env.n_act), generator=gen, device=env.device) * 4.0
out = {"replicas": R, "links": env.engine.L, "launches_per_step": env.launches_per_step(), "windows": {}}
t = 1
for lo, hi in wins:
    while t < lo:
        env.step(a[t & 7]); t += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while t < hi:
        env.step(a[t & 7]); t += 1
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (hi - lo)
    out["windows"][f"{lo}-{hi}"] = {"us_per_env_step": round(us, 2), "env_steps_per_s": round(R / us * 1e6),
                                   "link_steps_per_s": round(R * env.engine.L / us * 1e6),
                                   "roofline_frac_176B": round(R * env.engine.L * 176 / (us * 1e-6) / 6455.6e9, 4)}
env.engine.check_errors()
num = env.engine.history("num_pedestrians")[t - 1]
out["mean_pedestrians_per_link"] = float(num.mean())
out["max_pedestrians_on_a_link"] = float(num.max())
print(json.dumps(out))
