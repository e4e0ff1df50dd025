"""Timing of pns_lp_solve (one warp per program) on synthetic congested programs.  usage: lp_probe.py [n]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from pednstream_b200 import _native                         # noqa: E402
from test_optimal_node_model import synthetic_programs      # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 46080
lib = _native.load()
out = {}
for m in (3, 4, 5, 8):
    s, r, phi = synthetic_programs(m, n, seed=3)
    E = m * (m - 1)
    ds, dr, dp = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (s, r, phi))
    dx = torch.zeros((n, E), dtype=torch.float64, device="cuda")
    do = torch.zeros((n,), dtype=torch.float64, device="cuda")
    di = torch.zeros((n,), dtype=torch.int32, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())                  # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        assert lib.pns_lp_solve(m, n, P(ds), P(dr), P(dp), 0.01, P(dx), P(do), P(di), st) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.pns_lp_solve(m, n, P(ds), P(dr), P(dp), 0.01, P(dx), P(do), P(di), st)
    e1.record()
    torch.cuda.synchronize()
    piv = (di.cpu().numpy() & 0xfffffff)
    out[f"m{m}"] = {"us_per_launch": 1e3 * e0.elapsed_time(e1) / 5, "ns_per_program": 1e6 * e0.elapsed_time(e1) / 5 / n,
                    "pivots_mean": float(piv.mean()), "pivots_max": int(piv.max())}
print(json.dumps(out))
