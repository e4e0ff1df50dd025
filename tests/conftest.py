import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "emu")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the live reference tree at /root/reference")


@pytest.fixture(scope="session")
def emu_lib():
    """Host-emulation build of the kernels (tests/emu) -- CPU-side unit tests of kernel logic."""
    import build_emu
    from pednstream_b200 import _native
    return _native._declare(ctypes.CDLL(build_emu.build()))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def row_digests(a):
    import hashlib
    out = np.empty(a.shape[0], dtype=np.uint64)
    for t in range(a.shape[0]):
        out[t] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a[t]).tobytes()).digest()[:8],
                               dtype=np.uint64)[0]
    return out


LINK_FIELDS = ("inflow", "outflow", "cumulative_inflow", "cumulative_outflow", "sending_flow",
               "receiving_flow", "back_gate_width_data", "num_pedestrians", "density", "speed",
               "travel_time", "avg_travel_time", "link_flow")


def assert_matches_golden(gold, fields_by_name, steps, n_links):
    """Compare [S+1, >=L] arrays with a golden fixture on the rows that are final after `steps`
    steps: every row < steps, plus row `steps` of the fields a step writes at its own index."""
    for f in LINK_FIELDS:
        a = np.ascontiguousarray(fields_by_name[f][:, :n_links])
        got = row_digests(a[: steps + 1])
        want = gold["rows_" + f][: steps + 1]
        upto = steps if f in ("sending_flow", "receiving_flow") else steps + 1
        bad = np.nonzero(got[:upto] != want[:upto])[0]
        if len(bad):
            t = int(bad[0])
            cols = gold["sample_links"]
            raise AssertionError(
                f"field {f}: first differing row t={t} (of {upto}); sampled links {cols.tolist()}: "
                f"got {a[t, cols].tolist()} want {gold['sample_' + f][t].tolist()}")


def make_network(case, **kw):
    """Build the product-side network of a golden case (same seed protocol as gen_golden.py)."""
    import copy
    from oracle.gen_golden import CASES, LONG_CORRIDOR_EXAMPLE
    from pednstream_b200 import Network, NetworkEnvGenerator
    c = CASES[case]
    np.random.seed(c["seed"])
    if "inline" in c:
        spec = copy.deepcopy(LONG_CORRIDOR_EXAMPLE)
        return Network(np.array(spec["adjacency"]), spec["params"], origin_nodes=spec["origin_nodes"],
                       verbose=False, **kw)
    g = NetworkEnvGenerator()
    if c.get("steps_override") or c.get("default_link") or c.get("params"):
        g.network_data = g.load_network_data(c["dataset"])
        if c.get("steps_override"):
            g.config["params"]["simulation_steps"] = c["steps_override"]
        if c.get("default_link"):
            g.config["params"]["default_link"].update(c["default_link"])
        if c.get("params"):
            g.config["params"].update(c["params"])
    net = g.create_network(c["dataset"], verbose=False, **({} if c.get("randomize") is not None else kw))
    if c.get("randomize") is not None:
        net = g.randomize_network(c["dataset"], seed=c["randomize"], verbose=False, **kw)
    return net
