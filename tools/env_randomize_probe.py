"""Throughput and reset cost of the batched environment with per-replica randomised scenarios (experiment)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pednstream_b200.rl import BatchedPedNetEnv
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
out = {}
for rnd in (False, True):
    t0 = time.perf_counter()
    env = BatchedPedNetEnv("45_intersections", replicas=R, obs_mode="option3", seed=1, randomize=rnd, device="cuda:0")
    torch.cuda.synchronize()
    build = time.perf_counter() - t0
    a = torch.rand((R, env.n_act), device=env.device) * 4.0
    for _ in range(10):
        env.step(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        env.step(a)
    e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter(); env.reset(); torch.cuda.synchronize(); reset = time.perf_counter() - t0
    env.engine.check_errors()
    out["randomize" if rnd else "plain"] = {"env_steps_per_s": R * 200 / (e0.elapsed_time(e1) * 1e-3), "construct_s": build,
                                            "reset_s": reset, "classes": int(env.engine.net.n_classes)}
print(json.dumps(out))
