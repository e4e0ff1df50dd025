"""Per-chunk step time of the 512x512 lattice as the network fills up (experiment)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pednstream_b200.engine import Engine
from pednstream_b200.grid import build_grid_plan
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S = 1100
np.random.seed(0)
plan, gate, tf, demand = build_grid_plan(size, S, locality_order=True)
eng = Engine(plan, replicas=1, rng="philox", seed=0, device="cuda:0")
eng.initialise(gate, None, tf, demand, None)
t = 1
eng.run(t, 10); t += 10
torch.cuda.synchronize()
for chunk in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run(t, 100); e1.record(); torch.cuda.synchronize()
    t += 100
    num = eng.history("num_pedestrians")[t - 1, :, 0]
    snd = eng.history("sending_flow")[t - 2, :plan["n_links"], 0]
    print(json.dumps({"t": t - 1, "us_per_step": round(e0.elapsed_time(e1) * 10, 2), "peds": float(num.sum()),
                      "links_occupied": float((num > 0).float().mean()), "links_sending": float((snd > 0).float().mean()),
                      "max_num": float(num.max())}), flush=True)
eng.check_errors()
