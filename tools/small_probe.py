import sys, numpy as np, torch
sys.path.insert(0,'.')
from pednstream_b200 import NetworkEnvGenerator
for name in ("long_corridor","nine_intersections","45_intersections","melbourne"):
    np.random.seed(0)
    net = NetworkEnvGenerator("data").create_network(name, verbose=False, rng="philox", seed=3, device="cuda:0")
    eng = net.engine
    net.network_loading(1)
    eng.run(2, 50)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run(52, 200); e1.record(); torch.cuda.synchronize()
    ms, cnt = eng.run_profiled(252, 100)
    print(name, "us/step chained", 1e3*e0.elapsed_time(e1)/200, "per kernel us", [round(1e3*m/c,2) if c else None for m,c in zip(ms,cnt)], "rows", eng.net.n_rows, "dyn", eng.net.n_dyn_rows, "groups", eng.net.n_groups, "opts", eng.net.n_opts, "terms", eng.net.n_terms)
