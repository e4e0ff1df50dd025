"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the *live, unmodified* reference.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden [case ...]

The reference has no golden vectors of its own, so parity is pinned on
trajectories produced here by importing the reference (oracle/ref_harness.py) with
`np.random.seed(seed)` before network construction.  Per case the fixture stores, for each of the
13 per-link history fields, one 64-bit digest per time row (sha256 of the row bytes over the
physical links, reference `network.links` order) plus the full series of a few links and the
origin demand, which is enough to localise a divergence to (field, first bad row).
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (dataset | None for inline params, steps simulated, steps override, seed)
CASES = {
    "long_corridor_example": dict(inline="long_corridor_example", run=499, seed=0),
    "long_corridor": dict(dataset="long_corridor", run=599, seed=0),
    "nine_intersections": dict(dataset="nine_intersections", run=499, seed=0),
    "45_intersections": dict(dataset="45_intersections", run=699, seed=0),
    "butterfly_scA": dict(dataset="butterfly_scA", run=599, seed=0),
    "small_network": dict(dataset="small_network", run=499, seed=0),
    "one_intersection_v0": dict(dataset="one_intersection_v0", run=599, seed=0),
    "od_flow_example": dict(dataset="od_flow_example", run=499, seed=0),
    "delft": dict(dataset="delft", run=120, seed=0),
    "melbourne_2000": dict(dataset="melbourne", run=1999, steps_override=2000, seed=0),
    # domain randomisation (reference env_loader.py:160-424): create the scenario, then
    # NetworkEnvGenerator.randomize_network(dataset, seed=randomize) and run the perturbed network
    "45_intersections_rand7": dict(dataset="45_intersections", run=400, seed=0, randomize=7),
    "nine_intersections_rand3": dict(dataset="nine_intersections", run=300, seed=0, randomize=3),
    "delft_rand11": dict(dataset="delft", run=80, seed=0, randomize=11),
    # the third fundamental diagram (reference src/utils/functions.py:124-128): no shipped scenario selects it, so
    # the scenario's default_link block is overridden before the network is built
    "nine_intersections_smulders": dict(dataset="nine_intersections", run=499, seed=0,
                                        default_link={"fd_type": "smulders"}),
    "45_intersections_smulders": dict(dataset="45_intersections", run=400, seed=2,
                                      default_link={"fd_type": "smulders", "speed_noise_std": 0.2}),
    # the 'optimal' node model (reference src/LTM/node.py:249-271: one scipy.optimize.linprog per regular node and
    # step): no shipped scenario selects it, so the simulation block's assign_flows_type is overridden
    "nine_intersections_optimal": dict(dataset="nine_intersections", run=300, seed=0,
                                       params={"assign_flows_type": "optimal"}),
    "45_intersections_optimal": dict(dataset="45_intersections", run=150, seed=1,
                                     params={"assign_flows_type": "optimal"}),
}
LP_CASES = ("nine_intersections_optimal", "45_intersections_optimal")
LP_PER_SHAPE = 400          # programs kept per slot count in tests/golden/lp_programs.npz

# parameters of reference examples/long_corridor.py:25-63 (scenario 1), restated as data
LONG_CORRIDOR_EXAMPLE = dict(
    adjacency=[[0, 1, 0, 0, 0, 0], [1, 0, 1, 0, 0, 0], [0, 1, 0, 1, 0, 0],
               [0, 0, 1, 0, 1, 0], [0, 0, 0, 1, 0, 1], [0, 0, 0, 0, 1, 0]],
    params={"unit_time": 10, "simulation_steps": 600,
            "default_link": {"length": 100, "width": 2, "free_flow_speed": 1.1, "k_critical": 2,
                             "k_jam": 6, "fd_type": "yperman", "bi_factor": 1, "controller_type": "gate"},
            "demand": {"origin_0": {"peak_lambda": 25, "base_lambda": 5},
                       "origin_5": {"peak_lambda": 25, "base_lambda": 5}}},
    origin_nodes=[5, 0])


def row_digests(a: np.ndarray) -> np.ndarray:
    out = np.empty(a.shape[0], dtype=np.uint64)
    for t in range(a.shape[0]):
        out[t] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a[t]).tobytes()).digest()[:8], dtype=np.uint64)[0]
    return out


def build_reference_network(case):
    import copy
    np.random.seed(case["seed"])
    if "inline" in case:
        Network, _ = rh.import_reference()
        spec = copy.deepcopy(LONG_CORRIDOR_EXAMPLE)
        net = Network(np.array(spec["adjacency"]), spec["params"], origin_nodes=spec["origin_nodes"])
        import logging
        net.logger.setLevel(logging.ERROR)
        return net
    net, gen = rh.create_network(case["dataset"], steps_override=case.get("steps_override"),
                                 default_link=case.get("default_link"), params=case.get("params"))
    if case.get("randomize") is not None:
        import logging
        net = gen.randomize_network(case["dataset"], seed=case["randomize"])
        if net.logger is not None:
            net.logger.setLevel(logging.ERROR)
    return net


def generate(name):
    case = CASES[name]
    net = build_reference_network(case)
    t0 = time.time()
    for t in range(1, case["run"] + 1):
        net.network_loading(t)
    wall = time.time() - t0
    arrays = rh.collect_link_arrays(net)
    L = arrays["inflow"].shape[1]
    keep = sorted(set([0, 1, L // 2, L - 1]))
    out = {"steps_run": np.int64(case["run"]), "seed": np.int64(case["seed"]),
           "sim_steps": np.int64(net.simulation_steps), "n_links": np.int64(L),
           "link_keys": arrays["link_keys"], "sample_links": np.asarray(keep, dtype=np.int64),
           "reference_seconds": np.float64(wall),
           "origin_nodes": np.asarray(net.origin_nodes, dtype=np.int64),
           "destination_nodes": np.asarray(net.destination_nodes, dtype=np.int64)}
    for f in rh.LINK_FIELDS:
        out["rows_" + f] = row_digests(arrays[f])
        out["sample_" + f] = arrays[f][:, keep]
    dem_nodes = [n for n in net.nodes.values() if n.demand is not None]
    out["demand_nodes"] = np.asarray([n.node_id for n in dem_nodes], dtype=np.int64)
    for n in dem_nodes:
        out[f"demand_{n.node_id}"] = np.asarray(n.demand)
    out["total_in"] = arrays["cumulative_inflow"][case["run"]].sum()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {L} links, {case['run']} steps, reference {wall:.1f}s -> {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)", flush=True)


def generate_lp_programs():
    """Programs that the reference's 'optimal' node model hands to scipy.optimize.linprog while it runs LP_CASES,
    with the solutions linprog returned (node.py:262): per slot count m the right-hand sides (s, r), the turning
    fractions, x and the objective.  Every second program kept comes from a step where a receiving flow binds."""
    import sys
    rng = np.random.RandomState(0)
    kept = {}
    for name in LP_CASES:
        case = CASES[name]
        net = build_reference_network(case)
        recs = []
        with rh.reference_modules():
            nodemod = sys.modules["src.LTM.node"]
            solver = nodemod.linprog

            def spy(c, **kw):
                res = solver(c, **kw)
                if res.success:
                    recs.append((kw["b_ub"].copy(), kw["A_eq"].copy(), res.x.copy(), float(res.fun)))
                return res
            nodemod.linprog = spy
            try:
                for t in range(1, case["run"] + 1):
                    net.network_loading(t)
            finally:
                nodemod.linprog = solver
        for b, A_eq, x, fun in recs:
            m = len(b) // 2
            E = m * (m - 1)
            phi = A_eq[np.arange(E), np.arange(E)] + 1                       # node.py:131: the diagonal holds phi - 1
            xe = x[:E].reshape(m, m - 1)
            inflow = np.zeros(m)
            for i in range(m):
                for k in range(m - 1):
                    inflow[k if k < i else k + 1] += xe[i, k]
            binding = bool(np.any(inflow >= b[m:] - 1e-9)) and b[:m].sum() > 0
            kept.setdefault((m, binding), []).append((b[:m], b[m:], phi, x[:E], fun))
    out = {}
    for m in sorted({k[0] for k in kept}):
        rows = []
        for binding in (True, False):
            pool = kept.get((m, binding), [])
            take = rng.permutation(len(pool))[:LP_PER_SHAPE // 2]
            rows += [pool[i] for i in take]
        out[f"s_{m}"] = np.array([r[0] for r in rows])
        out[f"r_{m}"] = np.array([r[1] for r in rows])
        out[f"phi_{m}"] = np.array([r[2] for r in rows])
        out[f"x_{m}"] = np.array([r[3] for r in rows])
        out[f"objective_{m}"] = np.array([r[4] for r in rows])
        print(f"lp_programs: m={m}: {len(rows)} programs "
              f"({len(kept.get((m, True), []))} binding / {len(kept.get((m, False), []))} free recorded)", flush=True)
    path = os.path.join(GOLDEN_DIR, "lp_programs.npz")
    np.savez_compressed(path, **out)
    print(f"-> {path} ({os.path.getsize(path) / 1024:.0f} KiB)", flush=True)


# ---- control-environment fixtures (reference rl/pz_pednet_env.py with scripted actions) -----------
ENV_CASES = {
    "env_nine_intersections": dict(dataset="nine_intersections", obs_mode="option3", normalize_obs=False,
                                   seed=123, steps=300),
    "env_45_intersections": dict(dataset="45_intersections", obs_mode="option3", normalize_obs=False,
                                 seed=123, steps=250),
    "env_long_corridor": dict(dataset="long_corridor", obs_mode="option1", normalize_obs=False,
                              seed=5, steps=300),
    "env_nine_intersections_opt2n": dict(dataset="nine_intersections", obs_mode="option2", normalize_obs=True,
                                         seed=9, steps=120),
    "env_butterfly_opt5": dict(dataset="butterfly_scA", obs_mode="option5", normalize_obs=False, seed=3, steps=200),
    # option4 = [get_density / k_jam, back gate width] per controlled link (reference rl/builders.py:147-152)
    "env_nine_intersections_opt4": dict(dataset="nine_intersections", obs_mode="option4", normalize_obs=False,
                                        seed=21, steps=250),
    "env_45_intersections_opt4": dict(dataset="45_intersections", obs_mode="option4", normalize_obs=False,
                                      seed=4, steps=150),
    # action_gap = 3: three simulation steps per decision, rewards summed over them (rl/pz_pednet_env.py:224-253)
    "env_45_intersections_gap3": dict(dataset="45_intersections", obs_mode="option3", normalize_obs=False,
                                      seed=6, steps=90, action_gap=3),
}


def scripted_actions(env, steps, seed=7):
    """[steps, n_act] float32: uniform over each agent's action box from a private RandomState
    (the environment's own global stream is left alone)."""
    rs = np.random.RandomState(seed)
    rows = []
    for _ in range(steps):
        row = []
        for a in env.possible_agents:
            sp = env.action_space(a)
            row.append(rs.uniform(sp.low, sp.high).astype(np.float32))
        rows.append(np.concatenate(row) if row else np.zeros(0, dtype=np.float32))
    return np.stack(rows)


def generate_env(name):
    case = ENV_CASES[name]
    env = rh.make_reference_env(case["dataset"], obs_mode=case["obs_mode"], normalize_obs=case["normalize_obs"],
                                seed=case["seed"], **({"action_gap": case["action_gap"]} if "action_gap" in case else {}))
    obs0, _ = env.reset()
    steps = case["steps"]
    actions = scripted_actions(env, steps)
    agents = list(env.possible_agents)
    widths = [env.action_space(a).shape[0] for a in agents]
    obs_rows, rew_rows, done_rows = [], [], []
    for k in range(steps):
        off, act = 0, {}
        for a, w in zip(agents, widths):
            act[a] = actions[k, off:off + w]
            off += w
        obs, rew, term, trunc, info = env.step(act)
        obs_rows.append(np.concatenate([obs[a] for a in agents]))
        rew_rows.append(np.array([float(rew.get(a, 0.0)) for a in agents], dtype=np.float64))
        done_rows.append(bool(term[agents[0]]))
    arrays = rh.collect_link_arrays(env.network)
    out = {"steps_run": np.int64(steps), "seed": np.int64(case["seed"]), "actions": actions,
           "agents": np.array(agents), "action_widths": np.array(widths, dtype=np.int64),
           "obs0": np.concatenate([obs0[a] for a in agents]), "obs": np.stack(obs_rows),
           "rewards": np.stack(rew_rows), "done": np.array(done_rows),
           "n_links": np.int64(arrays["inflow"].shape[1]), "sample_links": np.asarray([0, 1], dtype=np.int64)}
    for f in rh.LINK_FIELDS:
        out["rows_" + f] = row_digests(arrays[f])
        out["sample_" + f] = arrays[f][:, [0, 1]]
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(agents)} agents {agents}, {steps} env steps -> {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)", flush=True)


# ---- turning fractions per step (reference node.turning_fractions after every network_loading) -------------
TF_CASES = {"tf_nine_intersections": dict(dataset="nine_intersections", run=220, seed=0),
            "tf_45_intersections": dict(dataset="45_intersections", run=160, seed=1)}


def generate_tf(name):
    """Every routed node's turning fractions after each step, recorded from the live reference
    (path_finder.update_turning_fractions + check_fractions, path_finder.py:591-715) -- idle steps included."""
    case = TF_CASES[name]
    net = build_reference_network(case)
    routed = [n for n in net.nodes.values()
              if net.path_finder is not None and n.node_id in net.path_finder.nodes_in_paths and n.source_num > 2]
    series = {n.node_id: [] for n in routed}
    for t in range(1, case["run"] + 1):
        net.network_loading(t)
        for n in routed:
            series[n.node_id].append(np.array(n.turning_fractions, dtype=np.float64).copy())
    out = {"steps_run": np.int64(case["run"]), "seed": np.int64(case["seed"]),
           "routed_nodes": np.asarray([n.node_id for n in routed], dtype=np.int64)}
    for k, v in series.items():
        out[f"tf_{k}"] = np.stack(v)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(routed)} routed nodes, {case['run']} steps -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)",
          flush=True)


# ---- on-device (Philox) draw mode on jammed lattices: digests of the oracle run with the Python restatement of
# the device samplers.  No reference involved (the reference has no counter-based mode): these pin the
# single-replica lattice kernel path (k_link_lane, one parameter class, host-resolved lag rows, launch order)
# against the CPU restatement at sizes the unit tests do not reach.
LATTICE_CASES = {"lattice32_philox": dict(size=32, stride=2, run=320, seed=5, base=30, peak=50),
                 "lattice64_philox": dict(size=64, stride=8, run=300, seed=7, base=40, peak=70)}


def lattice_network(case):
    from pednstream_b200 import Network
    from pednstream_b200.grid import DEFAULT_LINK, default_origins, grid_adjacency
    size = case["size"]
    origins = default_origins(size, stride=case["stride"])
    params = {"unit_time": 10, "simulation_steps": case["run"] + 30, "default_link": dict(DEFAULT_LINK),
              "demand": {f"origin_{o}": {"peak_lambda": case["peak"], "base_lambda": case["base"]} for o in origins}}
    np.random.seed(case["seed"])
    return Network(grid_adjacency(size), params, origin_nodes=list(origins), verbose=False), origins


def generate_lattice(name):
    from oracle.ltm_oracle import F32_FIELDS, F64_FIELDS, LtmOracle
    from oracle.philox import PhiloxDraws
    case = LATTICE_CASES[name]
    net, origins = lattice_network(case)
    t0 = time.time()
    h = LtmOracle(net, draws=PhiloxDraws(seed=case["seed"])).run(case["run"])
    L = len(net.links)
    out = {"steps_run": np.int64(case["run"]), "seed": np.int64(case["seed"]), "n_links": np.int64(L),
           "sim_steps": np.int64(net.simulation_steps), "origins": np.asarray(origins, dtype=np.int64),
           "max_pedestrians": np.float64(h["num_pedestrians"][: case["run"] + 1].max()),
           "jammed_link_steps": np.int64((h["num_pedestrians"][: case["run"] + 1] >= 1100).sum())}
    for f in F64_FIELDS[:7] + F32_FIELDS:
        out["rows_" + f] = row_digests(np.ascontiguousarray(h[f][: case["run"] + 1, :L]))
    dem = net.plan["demand_nodes"]
    out["demand"] = np.stack([np.asarray(n.demand, dtype=np.float64)[: net.simulation_steps] for n in dem], axis=1)
    out["demand_node_ids"] = np.asarray([n.node_id for n in dem], dtype=np.int64)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {L} links, {case['run']} steps, oracle {time.time() - t0:.0f}s, max pedestrians on a link "
          f"{out['max_pedestrians']:.0f}, jammed link-steps {int(out['jammed_link_steps'])} -> {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)", flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES) + list(ENV_CASES) + list(TF_CASES) + list(LATTICE_CASES) + ["lp_programs"]):
        if nm == "lp_programs":
            generate_lp_programs()
        elif nm in ENV_CASES:
            generate_env(nm)
        elif nm in TF_CASES:
            generate_tf(nm)
        elif nm in LATTICE_CASES:
            generate_lattice(nm)
        else:
            generate(nm)
