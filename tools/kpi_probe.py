"""Time of the device KPI pass for a finished batched episode (experiment)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pednstream_b200.rl import BatchedPedNetEnv
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
env = BatchedPedNetEnv("45_intersections", replicas=R, obs_mode="option3", seed=1, device="cuda:0")
a = torch.rand((R, env.n_act), device=env.device) * 4.0
for _ in range(env.simulation_steps):
    env.step(a)
torch.cuda.synchronize()
env.kpis()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    k = env.kpis()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
S, L = env.simulation_steps, env.engine.L
read = 12.0 * (S + 1) * L * R
print(json.dumps({"replicas": R, "kpi_ms": ms, "history_bytes_read": read, "GBps": read / (ms * 1e-3) / 1e9,
                  "episodes_per_s": R / (ms * 1e-3), "mean_throughput": float((k[:, 1] / k[:, 0]).mean())}))
