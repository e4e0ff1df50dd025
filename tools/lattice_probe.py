"""Lattice probe for profiler captures: runs the 512x512 lattice to step T0 untimed, then N more steps.
usage: lattice_probe.py T0 N [origin_stride]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pednstream_b200.engine import Engine
from pednstream_b200.grid import build_grid_plan, default_origins

T0, N = int(sys.argv[1]), int(sys.argv[2])
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 64
S = T0 + N + 4
plan, gate, tf, demand = build_grid_plan(512, S, locality_order=True, origins=default_origins(512, stride))
eng = Engine(plan, replicas=1, rng="philox", seed=0, device="cuda:0")
eng.initialise(gate, None, tf, demand, None)
eng.run(1, T0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.run(T0 + 1, N)
e1.record()
torch.cuda.synchronize()
print("us per step", 1e3 * e0.elapsed_time(e1) / N)
eng.check_errors()
print("pedestrians on links", float(eng.history("num_pedestrians")[T0 + N].sum()))
