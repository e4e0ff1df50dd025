"""Does a CUDA graph help?  Captures the launch chain of a 20-step `pns_step` call (and of 20 environment steps)
once and replays it, against issuing the same launches directly.  A captured graph is only valid for the steps it
was captured for (the kernels' parameters carry the rows of their step), so this measures what a graph could save
at best -- the replay re-runs steps t0 .. t0+19 on the same state."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pednstream_b200.engine import Engine          # noqa: E402
from pednstream_b200.grid import build_grid_plan   # noqa: E402


def events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def time_it(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = events()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
K = 20
plan, gate, tf, demand = build_grid_plan(512, 200)
eng = Engine(plan, replicas=1, rng="philox", seed=1, device="cuda:0")
eng.initialise(gate, np.zeros(len(gate), np.int32), tf, demand, None, None)
eng.run(1, 5)
torch.cuda.synchronize()
direct = time_it(lambda: eng.run(6, K), 10)
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
try:
    with torch.cuda.stream(side):
        eng.run(6, K)                          # warm-up on the capture stream
        with torch.cuda.graph(g, stream=side):
            eng.run(6, K)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    replay = time_it(g.replay, 10)
    out["lattice512"] = {"us_per_step_direct": 1e3 * direct / K, "us_per_step_graph_replay": 1e3 * replay / K}
except Exception as exc:                       # noqa: BLE001
    out["lattice512"] = {"us_per_step_direct": 1e3 * direct / K, "capture_failed": repr(exc)[:300]}
del eng
torch.cuda.empty_cache()

from pednstream_b200.rl import BatchedPedNetEnv    # noqa: E402
for R in (1024, 8192):
    env = BatchedPedNetEnv("45_intersections", replicas=R, obs_mode="option3", seed=1, device="cuda:0")
    acts = torch.rand((K, R, env.n_act), device="cuda") * 4.0
    obs = torch.empty((K, R, env.n_obs), dtype=torch.float32, device="cuda")
    rew = torch.empty((K, R), dtype=torch.float32, device="cuda")
    env.rollout(acts[:5], obs[:5], rew[:5])

    def steps():
        env.sim_step = 6
        for k in range(K):
            env.step(acts[k], obs[k], rew[k])

    def roll():
        env.sim_step = 6
        env.rollout(acts, obs, rew)

    per_step = time_it(steps, 5)
    native = time_it(roll, 5)
    rec = {"us_per_env_step_python_loop": 1e3 * per_step / K, "us_per_env_step_native_rollout": 1e3 * native / K}
    try:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            roll()
            with torch.cuda.graph(g, stream=side):
                roll()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        rec["us_per_env_step_graph_replay"] = 1e3 * time_it(g.replay, 5) / K
    except Exception as exc:                   # noqa: BLE001
        rec["capture_failed"] = repr(exc)[:300]
    out[f"env45_{R}"] = rec
    del env, g
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
