"""Timing of the 'optimal' node model (one linear program per regular node, replica and step) on the GPU."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pednstream_b200.engine import Engine          # noqa: E402
from pednstream_b200.grid import build_grid_plan   # noqa: E402


def timed(eng, t0, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run(t0, n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


torch.manual_seed(0)
out = {}
only_env = len(sys.argv) > 1 and sys.argv[1] == "env"          # `env`: the batched part only (ncu captures)
for model in ("classic", "optimal"):
    for size in (() if only_env else (128, 512)):
        S = 260
        plan, gate, tf, demand = build_grid_plan(size, S, node_model=model)
        eng = Engine(plan, replicas=1, rng="philox", seed=1, device="cuda:0")
        eng.initialise(gate, np.zeros(len(gate), np.int32), tf, demand, None, None)
        eng.run(1, 200)
        ms = timed(eng, 201, 40)
        eng.check_errors()
        out[f"lattice{size}_{model}"] = {"ms_per_step": ms, "lp_nodes": int(eng.net.n_lp_nodes),
                                         "us_per_program": 1e3 * ms / max(1, eng.net.n_lp_nodes)}
        del eng
        torch.cuda.empty_cache()

from pednstream_b200.rl import BatchedPedNetEnv    # noqa: E402
for model in ("classic", "optimal"):
    env = BatchedPedNetEnv("45_intersections", replicas=1024, obs_mode="option3", seed=1,
                           params={"assign_flows_type": model})
    env.reset()
    a = torch.rand((40, 1024, env.n_act), device="cuda") * 2 - 1
    env.rollout(a[:20])
    torch.cuda.synchronize()
    t0 = time.time()
    env.rollout(a[20:])
    torch.cuda.synchronize()
    out[f"env45_1024_{model}"] = {"us_per_env_step": (time.time() - t0) / 20 * 1e6,
                                  "lp_nodes": int(env.engine.net.n_lp_nodes)}
print(json.dumps(out, indent=1))
