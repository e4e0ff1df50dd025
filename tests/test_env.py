"""Control environments: the single-network facade vs fixtures recorded from the reference's own
PedNetParallelEnv (scripted actions), and the batched device environment vs the facade."""
import numpy as np
import pytest
import torch

from conftest import assert_matches_golden, load_golden
from oracle.ltm_oracle import F32_FIELDS, F64_FIELDS
from pednstream_b200.rl import BatchedPedNetEnv, PedNetParallelEnv

ENV_CASES = {"env_nine_intersections": ("nine_intersections", "option3", False),
             "env_45_intersections": ("45_intersections", "option3", False),
             "env_long_corridor": ("long_corridor", "option1", False),
             "env_nine_intersections_opt2n": ("nine_intersections", "option2", True),
             "env_butterfly_opt5": ("butterfly_scA", "option5", False),
             "env_nine_intersections_opt4": ("nine_intersections", "option4", False),
             "env_45_intersections_opt4": ("45_intersections", "option4", False),
             "env_45_intersections_gap3": ("45_intersections", "option3", False, 3)}


def run_env_case(name, **engine_kw):
    dataset, obs_mode, norm, *gap = ENV_CASES[name]
    gap = gap[0] if gap else 1
    gold = load_golden(name)
    env = PedNetParallelEnv(dataset, obs_mode=obs_mode, normalize_obs=norm, seed=int(gold["seed"]), action_gap=gap,
                            **engine_kw)
    agents = list(env.possible_agents)
    assert agents == gold["agents"].tolist()
    widths = gold["action_widths"].tolist()
    assert [env.action_space(a).shape[0] for a in agents] == widths
    obs0, _ = env.reset()
    assert np.array_equal(np.concatenate([obs0[a] for a in agents]), gold["obs0"])
    steps = int(gold["steps_run"])
    for k in range(steps):
        off, act = 0, {}
        for a, w in zip(agents, widths):
            act[a] = gold["actions"][k, off:off + w]
            off += w
        obs, rew, term, trunc, info = env.step(act)
        got = np.concatenate([obs[a] for a in agents])
        assert got.dtype == np.float32
        assert np.array_equal(got, gold["obs"][k]), f"observation differs at env step {k + 1}"
        r = np.array([float(rew.get(a, 0.0)) for a in agents])
        assert np.array_equal(r, gold["rewards"][k]), f"reward differs at env step {k + 1}: {r} vs {gold['rewards'][k]}"
        assert bool(term[agents[0]]) == bool(gold["done"][k]) and not any(trunc.values())
        assert info[agents[0]]["step"] == gap * (k + 1)
    fields = {f: env.network._store.field(f) for f in F64_FIELDS[:7] + F32_FIELDS}
    assert_matches_golden(gold, fields, gap * steps, int(gold["n_links"]))
    return env


@pytest.mark.parametrize("name", list(ENV_CASES))
def test_env_matches_reference_env_emulated(name, emu_lib):
    run_env_case(name, _lib=emu_lib, _emulation=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(ENV_CASES))
def test_env_matches_reference_env_cuda(name):
    run_env_case(name, device="cuda:0")


def test_env_api_surface(emu_lib):
    env = PedNetParallelEnv("nine_intersections", obs_mode="option3", seed=1, _lib=emu_lib, _emulation=True)
    assert env.possible_agents == ["gate_3", "gate_4", "gate_7"] and env.agents == env.possible_agents
    assert env.observation_space("gate_4").shape == (20,) and env.action_space("gate_4").shape == (4,)
    assert float(env.action_space("gate_4").high[0]) == 4.0
    with pytest.raises(ValueError):
        env.step({"gate_99": np.zeros(3, dtype=np.float32)})
    with pytest.raises(ValueError):
        env.observation_space("nope")
    with pytest.raises(ValueError):
        PedNetParallelEnv("nine_intersections", obs_mode="bogus", _lib=emu_lib, _emulation=True)


def test_env_randomized_reset(emu_lib):
    """reset(options={'randomize': True}) (pz_pednet_env.py:163-165) builds a perturbed scenario through
    randomize_network; with the numpy stream seeded the same way it is reproducible, and it differs from
    the plain scenario."""
    def episode(randomize, steps=25):
        env = PedNetParallelEnv("45_intersections", obs_mode="option3", seed=4, _lib=emu_lib, _emulation=True)
        np.random.seed(77)
        obs, _ = env.reset(options={"randomize": randomize})
        agents = env.possible_agents
        for _ in range(steps):
            obs, rew, *_ = env.step({a: np.full(env.action_space(a).shape, 2.0, dtype=np.float32) for a in agents})
        return env, np.concatenate([obs[a] for a in agents])
    a, oa = episode(True)
    b, ob = episode(True)
    c, oc = episode(False)
    assert np.array_equal(oa, ob)
    assert a.network.origin_nodes == b.network.origin_nodes
    changed = [k for k in a.network.links if a.network.links[k].k_jam != c.network.links[k].k_jam
               or a.network.links[k].free_flow_speed != c.network.links[k].free_flow_speed]
    assert changed, "randomisation must perturb some links"
    assert not np.array_equal(oa, oc)


# ---------------------------------------------------------------------------------- batched
def _batched_vs_facade(dataset, obs_mode, norm, steps, R, picks, lib=None, emulation=False, device=None):
    """Replica r of the batched environment must equal a single-network facade run that uses the
    same demand, the same actions and the Philox key of replica r."""
    kw = dict(_lib=lib, _emulation=emulation) if emulation else dict(device=device)
    benv = BatchedPedNetEnv(dataset, replicas=R, obs_mode=obs_mode, normalize_obs=norm, seed=21, **kw)
    dev = benv.device
    rs = np.random.RandomState(4)
    lo = benv._env_t["act_lo"].cpu().numpy()
    hi = benv._env_t["act_hi"].cpu().numpy()
    acts = rs.uniform(lo - 0.5, hi + 0.5, size=(steps, R, benv.n_act)).astype(np.float32)
    demand = benv.engine.demand.cpu().numpy().reshape(benv.simulation_steps + 1, -1, R)
    obs_b = [benv.obs.cpu().numpy().copy()]
    rew_b, done_b = [], []
    for k in range(steps):
        o, r, d, _ = benv.step(torch.from_numpy(acts[k]).to(dev))
        obs_b.append(o.cpu().numpy().copy())
        rew_b.append(r.cpu().numpy().copy())
        done_b.append(d)
    benv.engine.check_errors()
    for rep in picks:
        fkw = dict(_lib=lib, _emulation=True) if emulation else dict(device=device)
        env = PedNetParallelEnv(dataset, obs_mode=obs_mode, normalize_obs=norm, seed=21, rng="philox", **fkw)
        env.reset()
        for row, node in enumerate(env.network.plan["demand_nodes"]):
            node.demand = demand[:, row, rep].copy()
        eng = env.network.engine
        eng.io.seed = benv.engine.io.seed
        eng.io.replica_base = rep
        agents = env.possible_agents
        first = np.concatenate([env._observe()[a] for a in agents])
        assert np.array_equal(first, obs_b[0][rep])
        for k in range(steps):
            act = {a: acts[k, rep, benv.action_slices[a]] for a in agents}
            obs, rew, term, _, _ = env.step(act)
            assert np.array_equal(np.concatenate([obs[a] for a in agents]), obs_b[k + 1][rep]), (rep, k)
            assert np.float32(rew.get(agents[0], 0.0)) == rew_b[k][rep], (rep, k)
        for f in F64_FIELDS[:7] + F32_FIELDS:
            want = env.network._store.field(f)
            got = benv.engine.history(f)[:, :, rep].cpu().numpy()
            assert np.array_equal(want[: steps + 1], got[: steps + 1]), (f, rep)


def _check_device_scenarios_against_restatement(benv, picks):
    """randomize='device': what the kernel drew (read back through `scenario`) equals the Python restatement of
    k_scenario_draw (oracle/philox.py device_scenario), value for value."""
    from oracle import philox as ph
    net = benv.network
    corridors = benv._corridors()
    base = [(net.links[c].k_critical, net.links[c].k_jam, net.links[c].free_flow_speed) for c in corridors]
    rows = net.plan["demand_nodes"]
    origin_rows = [k for k, n in enumerate(rows) if n.node_id in net.origin_nodes]
    od_keys = list(net.plan["od_keys"])
    for rep in picks:
        over, weights, demand = ph.device_scenario(benv.device_scenario_seed(benv.episode - 1), benv.replica_base + rep,
                                                   len(corridors), int(len(corridors) * 0.2), base, len(od_keys), origin_rows)
        sc = benv.scenario(rep)
        want_over = {f"{corridors[c][0]}_{corridors[c][1]}": {"k_critical": v[0], "k_jam": v[1], "free_flow_speed": v[2]}
                     for c, v in over.items()}
        assert sc["link_params_overrides"] == want_over
        assert len(want_over) == int(len(corridors) * 0.2)
        for j, key in enumerate(od_keys):
            assert np.array_equal(sc["od_flows"][key], np.full(benv.simulation_steps + 1, weights[j]))
        names = {0: "gaussian_peaks", 1: "constant", 2: "sudden_demand"}
        for row, (pat, b, p) in demand.items():
            got = sc["demand_params_overrides"][f"origin_{rows[row].node_id}"]
            assert (got["pattern"], got["base_lambda"], got["peak_lambda"]) == (names[pat], b, p)
            assert 2.0 <= b < 10.0 and b + 5.0 <= p < 30.0 + 1e-9


def _randomized_batched_vs_facade(dataset, steps, R, picks, lib=None, emulation=False, device=None, mode=True):
    """randomize=True / 'device': replica r must equal the facade network built from `scenario(r)` through
    create_network's override arguments (what randomize_network does, minus the OD-node perturbation)."""
    kw = dict(_lib=lib, _emulation=emulation) if emulation else dict(device=device)
    benv = BatchedPedNetEnv(dataset, replicas=R, obs_mode="option3", seed=5, randomize=mode, **kw)
    if mode == "device":
        _check_device_scenarios_against_restatement(benv, picks)
    dev = benv.device
    rs = np.random.RandomState(9)
    acts = rs.uniform(0.0, 4.0, size=(steps, R, benv.n_act)).astype(np.float32)
    demand = benv.engine.demand.cpu().numpy().reshape(benv.simulation_steps + 1, -1, R)
    obs_b, rew_b = [benv.obs.cpu().numpy().copy()], []
    for k in range(steps):
        o, r, _, _ = benv.step(torch.from_numpy(acts[k]).to(dev))
        obs_b.append(o.cpu().numpy().copy())
        rew_b.append(r.cpu().numpy().copy())
    benv.engine.check_errors()
    assert benv.engine.net.n_classes > 1 and benv.engine.net.per_replica_scenario == 1
    for rep in picks:
        sc = benv.scenario(rep)
        assert sc["link_params_overrides"], "the scenario must perturb some corridor"
        fkw = dict(_lib=lib, _emulation=True) if emulation else dict(device=device)
        env = PedNetParallelEnv(dataset, obs_mode="option3", seed=5, rng="philox", **fkw)
        env.network = env.env_generator.create_network(dataset, verbose=False, rng="philox", **sc, **fkw)
        env._bind_network()
        for row, node in enumerate(env.network.plan["demand_nodes"]):
            node.demand = demand[:, row, rep].copy()
        eng = env.network.engine
        eng.io.seed = benv.engine.io.seed
        eng.io.replica_base = rep
        agents = env.possible_agents
        assert np.array_equal(np.concatenate([env._observe()[a] for a in agents]), obs_b[0][rep])
        for k in range(steps):
            obs, rew, *_ = env.step({a: acts[k, rep, benv.action_slices[a]] for a in agents})
            assert np.array_equal(np.concatenate([obs[a] for a in agents]), obs_b[k + 1][rep]), (rep, k)
            assert np.float32(rew.get(agents[0], 0.0)) == rew_b[k][rep], (rep, k)
        for f in F64_FIELDS[:7] + F32_FIELDS:
            want = env.network._store.field(f)
            got = benv.engine.history(f)[:, :, rep].cpu().numpy()
            assert np.array_equal(want[: steps + 1], got[: steps + 1]), (f, rep)
    # replicas differ from each other, and a new episode draws new scenarios
    a0, a1 = benv.scenario(0), benv.scenario(1)
    assert a0["link_params_overrides"] != a1["link_params_overrides"]
    before = benv.scenario(0)
    benv.reset()
    assert benv.scenario(0)["link_params_overrides"] != before["link_params_overrides"]


@pytest.mark.parametrize("mode", [True, "device"])
def test_batched_env_randomized_scenarios_emulated(emu_lib, mode):
    _randomized_batched_vs_facade("45_intersections", 60, 3, (0, 2), lib=emu_lib, emulation=True, mode=mode)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [True, "device"])
def test_batched_env_randomized_scenarios_cuda(mode):
    _randomized_batched_vs_facade("45_intersections", 150, 48, (0, 17, 47), device="cuda:0", mode=mode)


def _grouped_vs_randomize_network(dataset, steps, R, G, lib=None, emulation=False, device=None):
    """Scenario groups (per-group OD nodes): the first replica of every group must equal the network that
    `randomize_network(dataset, seed)` builds -- OD nodes, bottlenecks, OD weights and demand pattern, the whole of
    reference env_loader.py:160-181 -- and the groups' OD node sets must differ from each other."""
    from pednstream_b200.rl import GroupedPedNetEnv
    kw = dict(_lib=lib, _emulation=emulation) if emulation else dict(device=device)
    genv = GroupedPedNetEnv(dataset, replicas=R, groups=G, obs_mode="option3", seed=5, randomize="host", **kw)
    dev = genv.device
    rs = np.random.RandomState(3)
    acts = rs.uniform(0.0, 4.0, size=(steps, R, genv.n_act)).astype(np.float32)
    obs_b, rew_b = [genv.obs.cpu().numpy().copy()], []
    for k in range(steps):
        o, r, _, _ = genv.step(torch.from_numpy(acts[k]).to(dev))
        if not emulation:
            torch.cuda.synchronize()
        obs_b.append(o.cpu().numpy().copy())
        rew_b.append(r.cpu().numpy().copy())
    genv.check_errors()
    assert len({str(od) for od in genv.od_nodes}) > 1, "the groups must not all share one OD node set"
    base_od, exact = None, 0
    for g, (benv, sl) in enumerate(zip(genv.envs, genv.slices)):
        rep = sl.start                                    # global index of the group's first replica
        seed_g = benv.scenario_seed(0, 0)
        fkw = dict(_lib=lib, _emulation=True) if emulation else dict(device=device)
        env = PedNetParallelEnv(dataset, obs_mode="option3", seed=5, rng="philox", **fkw)
        if base_od is None:
            base_od = (list(env.env_generator.config["origin_nodes"]), list(env.env_generator.config["destination_nodes"]))
        state = np.random.get_state()
        if benv.od_nodes_seed == seed_g:
            exact += 1
            env.network = env.env_generator.randomize_network(dataset, seed=seed_g, verbose=False, rng="philox", **fkw)
        else:       # the perturbation of seed_g has no OD pair (the reference cannot build it): OD nodes of the next seed
            env.env_generator.generate_random_od_nodes(benv.od_nodes_seed)
            env.network = env.env_generator.create_network(dataset, verbose=False, rng="philox", **benv.scenario(0), **fkw)
        np.random.set_state(state)
        assert env.env_generator.config["origin_nodes"] == benv.od_nodes["origin_nodes"]
        assert env.env_generator.config["destination_nodes"] == benv.od_nodes["destination_nodes"]
        env._bind_network()
        demand = benv.engine.demand.cpu().numpy().reshape(benv.simulation_steps + 1, -1, benv.R)
        for row, node in enumerate(env.network.plan["demand_nodes"]):
            node.demand = demand[:, row, 0].copy()
        eng = env.network.engine
        eng.io.seed = benv.engine.io.seed
        eng.io.replica_base = benv.replica_base
        agents = env.possible_agents
        assert np.array_equal(np.concatenate([env._observe()[a] for a in agents]), obs_b[0][rep])
        for k in range(steps):
            obs, rew, *_ = env.step({a: acts[k, rep, genv.action_slices[a]] for a in agents})
            assert np.array_equal(np.concatenate([obs[a] for a in agents]), obs_b[k + 1][rep]), (g, k)
            assert np.float32(rew.get(agents[0], 0.0)) == rew_b[k][rep], (g, k)
        for f in F64_FIELDS[:7] + F32_FIELDS:
            want = env.network._store.field(f)
            got = benv.engine.history(f)[:, :, 0].cpu().numpy()
            assert np.array_equal(want[: steps + 1], got[: steps + 1]), (f, g)
    assert any((od["origin_nodes"], od["destination_nodes"]) != base_od for od in genv.od_nodes)
    assert exact >= G - 1, "most groups must be exactly randomize_network(seed)"
    assert genv.kpis().shape[0] == R


def test_grouped_env_od_node_perturbation_emulated(emu_lib):
    _grouped_vs_randomize_network("45_intersections", 50, 5, 3, lib=emu_lib, emulation=True)


def test_grouped_env_action_gap_and_rollout_emulated(emu_lib):
    """Groups with action_gap 2: a rollout equals the same decisions through step(), and the groups' replicas equal
    a stand-alone batched environment of that group (same seeds, same OD nodes)."""
    from pednstream_b200.rl import GroupedPedNetEnv
    kw = dict(_lib=emu_lib, _emulation=True)
    a = GroupedPedNetEnv("45_intersections", replicas=4, groups=2, seed=5, randomize="host", action_gap=2, **kw)
    b = GroupedPedNetEnv("45_intersections", replicas=4, groups=2, seed=5, randomize="host", action_gap=2, **kw)
    acts = torch.from_numpy(np.random.RandomState(1).uniform(0, 4, (8, 4, a.n_act)).astype(np.float32))
    obs, rew, _ = a.rollout(acts)
    for k in range(8):
        o, r, _, info = b.step(acts[k])
        assert info["step"] == 2 * (k + 1)
        assert torch.equal(o, obs[k]) and torch.equal(r, rew[k])
    g1 = a.envs[1]
    solo = BatchedPedNetEnv("45_intersections", replicas=g1.R, obs_mode="option3", seed=5, replica_base=g1.replica_base,
                            randomize="host", od_nodes_seed=g1.od_nodes_seed, action_gap=2, **kw)
    for k in range(8):
        o, r, _, _ = solo.step(acts[k][a.slices[1]])
        assert torch.equal(o, obs[k][a.slices[1]]) and torch.equal(r, rew[k][a.slices[1]])


@pytest.mark.gpu
def test_grouped_env_od_node_perturbation_cuda():
    _grouped_vs_randomize_network("45_intersections", 150, 40, 5, device="cuda:0")


def _batched_gap_vs_facade(gap, decisions, R, picks, lib=None, emulation=False, device=None):
    """BatchedPedNetEnv(action_gap=gap): replica r against the single-network environment with the same gap (which
    is pinned to the reference's own env by the env_45_intersections_gap3 fixture)."""
    kw = dict(_lib=lib, _emulation=emulation) if emulation else dict(device=device)
    benv = BatchedPedNetEnv("45_intersections", replicas=R, obs_mode="option3", seed=5, action_gap=gap, **kw)
    dev = benv.device
    acts = np.random.RandomState(2).uniform(0.0, 4.0, size=(decisions, R, benv.n_act)).astype(np.float32)
    demand = benv.engine.demand.cpu().numpy().reshape(benv.simulation_steps + 1, -1, R)
    obs_b, rew_b = [], []
    half = decisions // 2
    for k in range(half):                                   # first half through step(), second through rollout()
        o, r, _, info = benv.step(torch.from_numpy(acts[k]).to(dev))
        assert info["step"] == gap * (k + 1)
        obs_b.append(o.cpu().numpy().copy()); rew_b.append(r.cpu().numpy().copy())
    o, r, _ = benv.rollout(torch.from_numpy(acts[half:]).to(dev))
    obs_b += list(o.cpu().numpy()); rew_b += list(r.cpu().numpy())
    benv.engine.check_errors()
    assert benv.sim_step == gap * decisions + 1
    cum = benv.cumulative_reward.cpu().numpy()
    for rep in picks:
        fkw = dict(_lib=lib, _emulation=True) if emulation else dict(device=device)
        env = PedNetParallelEnv("45_intersections", obs_mode="option3", seed=5, rng="philox", action_gap=gap, **fkw)
        for row, node in enumerate(env.network.plan["demand_nodes"]):
            node.demand = demand[:, row, rep].copy()
        eng = env.network.engine
        eng.io.seed = benv.engine.io.seed
        eng.io.replica_base = rep
        agents = env.possible_agents
        for k in range(decisions):
            obs, rew, *_ = env.step({a: acts[k, rep, benv.action_slices[a]] for a in agents})
            assert np.array_equal(np.concatenate([obs[a] for a in agents]), obs_b[k][rep]), (rep, k)
            assert np.float32(rew.get(agents[0], 0.0)) == rew_b[k][rep], (rep, k)
        assert np.float32(env._cumulative_rewards[agents[0]]) == cum[rep]
        for f in F64_FIELDS[:7] + F32_FIELDS:
            want = env.network._store.field(f)
            got = benv.engine.history(f)[:, :, rep].cpu().numpy()
            assert np.array_equal(want[: gap * decisions + 1], got[: gap * decisions + 1]), (f, rep)


def test_batched_env_action_gap_emulated(emu_lib):
    _batched_gap_vs_facade(3, 20, 3, (0, 2), lib=emu_lib, emulation=True)


@pytest.mark.gpu
def test_batched_env_action_gap_cuda():
    _batched_gap_vs_facade(3, 60, 40, (0, 21, 39), device="cuda:0")


def test_batched_env_matches_facade_emulated(emu_lib):
    _batched_vs_facade("nine_intersections", "option3", False, 40, 3, (0, 2), lib=emu_lib, emulation=True)


def test_batched_env_separator_emulated(emu_lib):
    _batched_vs_facade("long_corridor", "option1", False, 60, 2, (1,), lib=emu_lib, emulation=True)


@pytest.mark.gpu
@pytest.mark.parametrize("dataset,obs_mode,norm,steps,R,picks", [
    ("45_intersections", "option3", False, 80, 64, (0, 33, 63)),
    ("nine_intersections", "option2", True, 60, 40, (7,)),
    ("long_corridor", "option1", False, 80, 33, (32,)),
    ("butterfly_scA", "option5", False, 90, 70, (3, 69)),
    ("nine_intersections", "option4", False, 70, 65, (0, 64))])
def test_batched_env_matches_facade_cuda(dataset, obs_mode, norm, steps, R, picks):
    _batched_vs_facade(dataset, obs_mode, norm, steps, R, picks, device="cuda:0")


@pytest.mark.gpu
def test_batched_env_full_episode_and_reset():
    env = BatchedPedNetEnv("45_intersections", replicas=256, obs_mode="option3", seed=2, device="cuda:0")
    assert env.n_act == 4 and env.n_obs == 20
    a = torch.full((256, 4), 2.0, dtype=torch.float32, device=env.device)
    done = False
    n = 0
    while not done:
        obs, rew, done, info = env.step(a)
        n += 1
    assert n == env.simulation_steps == 700
    env.engine.check_errors()
    assert torch.isfinite(obs).all() and torch.isfinite(env.cumulative_reward).all()
    assert float(env.cumulative_reward.std()) > 0          # replicas differ (independent demand / draws)
    first = env.reset()
    assert env.sim_step == 1 and float(first[:, :4].abs().sum()) == 0.0 and float(first[0, 4]) == 4.0


@pytest.mark.gpu
@pytest.mark.parametrize("dataset,obs_mode,R,steps", [("45_intersections", "option3", 96, 260),
                                                      ("nine_intersections", "option5", 130, 150),
                                                      ("long_corridor", "option2", 33, 200)])
def test_batched_env_fused_kernel_equals_standalone_kernels(dataset, obs_mode, R, steps, monkeypatch):
    """The batched link kernel (one thread per directed link and replica; actions, observations and reward
    ride along) against the pair-per-thread kernel with the stand-alone environment kernels: same
    observations, rewards and history, bit for bit."""
    outs = []
    for standalone in (False, True):
        if standalone:
            monkeypatch.setenv("PNS_PAIR_THREADS", "1")
        else:
            monkeypatch.delenv("PNS_PAIR_THREADS", raising=False)
        env = BatchedPedNetEnv(dataset, replicas=R, obs_mode=obs_mode, seed=9, device="cuda:0")
        gen = torch.Generator(device="cuda:0")
        gen.manual_seed(3)
        lo = torch.tensor([float(x) for x in env._env_t["act_lo"].cpu()], device="cuda:0")
        hi = torch.tensor([float(x) for x in env._env_t["act_hi"].cpu()], device="cuda:0")
        obs, rew = [env.obs.clone()], []
        for k in range(steps):
            a = (lo + (hi - lo) * torch.rand((R, env.n_act), generator=gen, device="cuda:0")).float()
            o, r, _, _ = env.step(a)
            obs.append(o.clone()); rew.append(r.clone())
        env.engine.check_errors()
        outs.append((torch.stack(obs).cpu(), torch.stack(rew).cpu(), env.cumulative_reward.cpu().clone(),
                     env.engine.hist64[:, : steps + 1].cpu(), env.engine.hist32[:, : steps + 1].cpu(),
                     env.engine.gate.cpu().clone()))
        del env
        torch.cuda.empty_cache()
    monkeypatch.delenv("PNS_PAIR_THREADS", raising=False)
    names = ("obs", "reward", "cumulative reward", "hist64", "hist32", "gate")
    for name, a, b in zip(names, outs[0], outs[1]):
        assert torch.equal(a, b), name
    assert float(outs[0][1].abs().sum()) > 0 or dataset == "long_corridor"


def _rollout_equals_steps(R, K, host, **kw):
    """`rollout` / `rollout_host` (K steps in one native call) against K calls of `step` with the same actions."""
    envs = [BatchedPedNetEnv("nine_intersections", replicas=R, obs_mode="option3", seed=4, **kw) for _ in range(2)]
    dev = envs[0].device
    rs = np.random.RandomState(2)
    actions = torch.from_numpy(rs.uniform(0.0, 4.0, size=(K, R, envs[0].n_act)).astype(np.float32))
    want_obs, want_rew = [], []
    for k in range(K):
        o, r, _, _ = envs[0].step(actions[k].to(dev))
        want_obs.append(o.cpu().clone()); want_rew.append(r.cpu().clone())
    if host:
        ha = actions.pin_memory()
        ho = torch.zeros((K, R, envs[1].n_obs), dtype=torch.float32).pin_memory()
        hr = torch.zeros((K, R), dtype=torch.float32).pin_memory()
        envs[1].rollout_host(ha[: K // 2], ho[: K // 2], hr[: K // 2])          # two calls: the look-ahead state carries over
        envs[1].rollout_host(ha[K // 2:], ho[K // 2:], hr[K // 2:])
        torch.cuda.synchronize()
        got_obs, got_rew = ho, hr
    else:
        o1, r1, _ = envs[1].rollout(actions[: K // 2].to(dev))
        o2, r2, _ = envs[1].rollout(actions[K // 2:].to(dev))
        got_obs, got_rew = torch.cat([o1, o2]).cpu(), torch.cat([r1, r2]).cpu()
    assert torch.equal(torch.stack(want_obs), got_obs) and torch.equal(torch.stack(want_rew), got_rew)
    assert envs[0].sim_step == envs[1].sim_step == K + 1
    assert torch.equal(envs[0].cumulative_reward.cpu(), envs[1].cumulative_reward.cpu())
    assert torch.equal(envs[0].engine.hist64[:, : K + 1].cpu(), envs[1].engine.hist64[:, : K + 1].cpu())
    assert torch.equal(envs[0].engine.hist32[:, : K + 1].cpu(), envs[1].engine.hist32[:, : K + 1].cpu())
    o, r, _, _ = envs[1].step(actions[0].to(dev))                              # and stepping continues from there
    o0, r0, _, _ = envs[0].step(actions[0].to(dev))
    assert torch.equal(o.cpu(), o0.cpu()) and torch.equal(r.cpu(), r0.cpu())


def test_batched_env_rollout_equals_steps_emulated(emu_lib):
    _rollout_equals_steps(3, 30, False, _lib=emu_lib, _emulation=True)


@pytest.mark.gpu
@pytest.mark.parametrize("host", [False, True])
def test_batched_env_rollout_equals_steps_cuda(host):
    _rollout_equals_steps(70, 120, host, device="cuda:0")
