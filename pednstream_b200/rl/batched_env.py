"""`BatchedPedNetEnv`: R independent replicas of one scenario stepped per launch.

The per-replica semantics are those of the reference's `PedNetParallelEnv`
(rl/pz_pednet_env.py:143-254): actions are absolute widths (rate-limited, clipped,
rl/builders.py:264-352), `action_gap` `network_loading` steps per env step (default 1), observations in
the reference's link-major layout (rl/builders.py:68-177), the first agent's reward
(pz_pednet_env.py:548-581), termination after S env steps.  Everything per step -- action
application, the LTM step with on-device Philox draws, observation and reward -- is a CUDA kernel
over the replica-fastest state, so a warp is 32 replicas of the same link or node.

Tensors: `actions [R, n_act] float32` (agents concatenated in `possible_agents` order),
`obs [R, n_obs] float32`, `reward [R] float32`, `done [R] bool`.  Replicas are sharded over
GPUs by the caller (`replica_base` is the global index of local replica 0); no collective is
needed inside a step.  Demand is drawn on the device at reset (`pns_env_draw_demand`): Poisson around the
scenario's gaussian peaks (od_manager.py:145-155), keyed by the global replica index.

`randomize=True`: every replica and episode gets its own perturbed scenario from the
reference's generators (env_loader.py:183-258, 363-424) -- link bottlenecks (k_critical / k_jam /
free-flow speed of 20 % of the corridors), OD weights and demand patterns -- as per-replica parameter
classes, OD weights and demand on the device.  OD *nodes* are not perturbed (they would need one route
plan per replica).  Replica r equals a single-network facade built with
`create_network(dataset, od_flows=, link_params_overrides=, demand_params_overrides=)` from
`scenario(r)`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _native, ops
from ..engine import Engine, _ptr
from ..env_loader import NetworkEnvGenerator
from .builders import DENSITY_NORM, FLOW_NORM, OBS_LAYOUT
from .discovery import AgentManager


def _gater_divisors(obs_mode, k):
    """Per-feature divisor of the reference's fixed-constant normalisation (rl/builders.py:206-238)."""
    d = [1.0] * k
    if obs_mode in ("option1", "option2"):
        d[0] = d[1] = FLOW_NORM
    elif obs_mode in ("option3", "option4"):
        if k < 3:
            raise IndexError("normalize_obs with this obs_mode indexes past the observation vector "
                             "(the reference raises IndexError too)")
        d[0], d[1], d[2] = DENSITY_NORM, FLOW_NORM, FLOW_NORM
    return d


class BatchedPedNetEnv:
    def __init__(self, dataset: str, replicas: int, obs_mode: str = "option3", normalize_obs: bool = False,
                 seed: int = 0, replica_base: int = 0, device=None, data_dir="data", randomize: bool = False,
                 params: dict = None, od_nodes_seed: int = None, action_gap: int = 1, _lib=None,
                 _emulation: bool = False):
        """params: overrides of the scenario's parameter block (e.g. {"assign_flows_type": "optimal"}), applied
        before the template network is built.  od_nodes_seed: perturb the scenario's origin / destination nodes as
        the reference's generate_random_od_nodes(seed) does (env_loader.py:261-361) -- all replicas of this
        environment share that topology; `GroupedPedNetEnv` runs several such groups side by side."""
        if obs_mode not in OBS_LAYOUT:
            raise ValueError(f"obs_mode must be one of {list(OBS_LAYOUT)}, got: {obs_mode}")
        if int(action_gap) < 1:
            raise ValueError("action_gap must be >= 1")
        self.action_gap = int(action_gap)      # simulation steps per decision (rl/pz_pednet_env.py:110-111, 224-253)
        self.dataset, self.R, self.obs_mode, self.seed = dataset, int(replicas), obs_mode, int(seed)
        self.replica_base = int(replica_base)
        state = np.random.get_state()                 # building the template must not disturb the caller's stream
        np.random.seed(self.seed)
        self.generator = NetworkEnvGenerator(data_dir)
        if params:
            self.generator.network_data = self.generator.load_network_data(dataset)
            self.generator.config["params"].update(params)
        self.network = self.generator.create_network(dataset, verbose=False)
        self.od_nodes = self.od_nodes_seed = None
        if od_nodes_seed is not None:
            cfg = self.generator.config
            own = (list(cfg.get("origin_nodes", [])), list(cfg.get("destination_nodes", [])))
            for attempt in range(64):
                # a perturbation may leave no origin -> destination pair at all (e.g. both sets shrink to the same
                # node); the reference's network constructor fails on that (KeyError, path_finder.py:226): next seed
                self.od_nodes_seed = (int(od_nodes_seed) + 1000003 * attempt) % (2 ** 32)
                cfg["origin_nodes"], cfg["destination_nodes"] = list(own[0]), list(own[1])
                self.od_nodes = self.generator.generate_random_od_nodes(self.od_nodes_seed)
                if any(o != d for o in self.od_nodes["origin_nodes"] for d in self.od_nodes["destination_nodes"]):
                    break
            else:
                raise RuntimeError("no usable origin / destination perturbation found")
            od_flows = self.generator.generate_random_od_flows(self.od_nodes_seed)   # weights for the new OD pairs
            np.random.seed(self.seed)
            self.network = self.generator.create_network(dataset, od_flows=od_flows, verbose=False)
        np.random.set_state(state)
        if randomize not in (False, True, "host", "device"):
            raise ValueError("randomize must be False, True / 'host' or 'device'")
        self.randomize = "host" if randomize is True else randomize
        self._scenarios = None
        net = self.network
        self.simulation_steps = S = net.simulation_steps
        self.agent_manager = am = AgentManager(net)
        self.possible_agents = am.get_all_agent_ids()
        unit_time = net.params["unit_time"]
        max_delta = 0.25 * unit_time                   # pz_pednet_env.py:84-85
        min_sep = 1.5                                  # pz_pednet_env.py:86

        # ---- action / observation / reward programs (agents in possible_agents order) ----------
        act_link, act_sep, act_lo, act_hi, act_md, act_w = [], [], [], [], [], []
        obs_link, obs_src, obs_div = [], [], []
        self.action_slices, self.obs_slices = {}, {}
        for aid in self.possible_agents:
            a0, o0 = len(act_link), len(obs_link)
            if am.get_agent_type(aid) == "sep":
                fwd, rev = am.get_separator_links(aid)
                act_link.append(fwd.index); act_sep.append(1)
                act_lo.append(min_sep); act_hi.append(fwd.width - min_sep)
                act_md.append(max_delta); act_w.append(fwd._width)
                for l, src in ((fwd, "inflow"), (fwd, "outflow"), (rev, "inflow"), (rev, "outflow")):
                    obs_link.append(l.index); obs_src.append(_native.OBS_SRC[src])
                if normalize_obs:
                    if obs_mode in ("option3", "option4"):
                        raise IndexError("normalize_obs with this obs_mode indexes past the separator "
                                         "observation (the reference raises IndexError too)")
                    obs_div += [FLOW_NORM] * 4 if obs_mode in ("option1", "option2") else [1.0] * 4
                else:
                    obs_div += [1.0] * 4
            else:
                feats = OBS_LAYOUT[obs_mode]
                div = _gater_divisors(obs_mode, len(feats)) if normalize_obs else [1.0] * len(feats)
                for link in am.get_gater_outgoing_links(aid):
                    act_link.append(link.index); act_sep.append(0)
                    act_lo.append(0.0); act_hi.append(link.width)
                    act_md.append(max_delta); act_w.append(link._width)
                    for f, d in zip(feats, div):
                        obs_link.append(link.index); obs_src.append(_native.OBS_SRC[f]); obs_div.append(d)
            self.action_slices[aid] = slice(a0, len(act_link))
            self.obs_slices[aid] = slice(o0, len(obs_link))
        first = self.possible_agents[0] if self.possible_agents else None
        reward_links = ([l.index for l in am.get_gater_outgoing_links(first)]
                        if first is not None and am.get_agent_type(first) == "gate" else [])
        self.n_act, self.n_obs = len(act_link), len(obs_link)

        # ---- device runtime --------------------------------------------------------------------
        self.engine = eng = Engine(net.plan, replicas=self.R, device=device, rng="philox", seed=self.seed,
                                   lib=_lib, emulation=_emulation)
        eng.io.replica_base = self.replica_base
        dev = eng.device
        self.device = dev
        to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=dt))).to(dev)
        self._env_t = dict(act_link=to(act_link, np.int32), act_sep=to(act_sep, np.int32),
                           act_lo=to(act_lo, np.float64), act_hi=to(act_hi, np.float64),
                           act_max_delta=to(act_md, np.float64), act_total_width=to(act_w, np.float64),
                           obs_link=to(obs_link, np.int32), obs_src=to(obs_src, np.int32),
                           obs_div=to(obs_div, np.float32), reward_link=to(reward_links, np.int32))
        # the same programs indexed by directed link (the batched link kernel applies actions and emits
        # observations itself): REV_* entries are produced by the reverse link as its own inflow / outflow
        L = len(net.links)
        lk_act = np.full(L, -1, dtype=np.int32)
        for a, l in enumerate(act_link):
            if lk_act[l] >= 0:
                raise ValueError(f"link {l} is set by two actions")
            lk_act[l] = a
        by_link = [[] for _ in range(L)]
        rev_of = {_native.OBS_SRC["rev.inflow"]: _native.OBS_SRC["inflow"],
                  _native.OBS_SRC["rev.outflow"]: _native.OBS_SRC["outflow"]}
        for col, (l, src, d) in enumerate(zip(obs_link, obs_src, obs_div)):
            if src in rev_of:
                by_link[l ^ 1].append((col, rev_of[src], d))
            else:
                by_link[l].append((col, src, d))
        lk_obs_ptr = np.concatenate([[0], np.cumsum([len(b) for b in by_link])]).astype(np.int32)
        flat = [x for b in by_link for x in b]
        lk_reward = np.zeros(L, dtype=np.int32)
        for l in reward_links:
            lk_reward[l] = lk_reward[l ^ 1] = 1
        assert int(lk_reward.sum()) == 2 * len(reward_links)
        self._env_t.update(lk_act=to(lk_act, np.int32), lk_obs_ptr=to(lk_obs_ptr, np.int32),
                           lk_obs_col=to([x[0] for x in flat], np.int32),
                           lk_obs_src=to([x[1] for x in flat], np.int32),
                           lk_obs_div=to([x[2] for x in flat], np.float32), lk_reward=to(lk_reward, np.int32),
                           reward_count=torch.zeros((self.R,), dtype=torch.int32, device=dev))
        env = _native.PnsEnv()
        env.n_act, env.n_obs, env.n_reward_links = self.n_act, self.n_obs, len(reward_links)
        for k, t in self._env_t.items():
            setattr(env, k, _ptr(t))
        if not len(flat):                 # keep the per-link pointers non-null: they select the fused path
            env.lk_obs_col = env.lk_obs_src = env.lk_obs_div = _ptr(self._env_t["lk_obs_ptr"])
        self._env = env
        eng._env_struct = env
        self._no_actions = torch.zeros((1,), dtype=torch.float32, device=dev)
        self.obs = torch.zeros((self.R, max(1, self.n_obs)), dtype=torch.float32, device=dev)
        self.reward = torch.zeros((self.R,), dtype=torch.float32, device=dev)
        self.cumulative_reward = torch.zeros((self.R,), dtype=torch.float32, device=dev)
        self._tf, self._supplied = net._static_fractions()
        self._od_w = (np.stack([net.od_manager.od_flows[k] for k in net.plan["od_keys"]], axis=1)
                      if net.od_manager is not None else None)
        self.episode = 0
        self.sim_step = 1
        self.reset()

    # ------------------------------------------------------------------ per-replica scenarios
    def scenario_seed(self, replica: int, episode: int) -> int:
        return (self.seed + 7919 * (self.replica_base + replica) + 104729 * episode + 1) % (2 ** 32)

    def scenario(self, replica: int, episode: int = None) -> dict:
        """The overrides of local replica `replica` in `episode` (default: the current one), as the
        reference's generators produce them for `scenario_seed(replica, episode)`."""
        if self.randomize == "device":
            if episode is not None and episode != self._device_episode:
                raise ValueError("device scenarios are kept for the current episode only")
            return self._device_scenario(replica)
        episode = self.episode - 1 if episode is None else episode
        gen, s = self.generator, self.scenario_seed(replica, episode)
        state = np.random.get_state()
        try:
            out = {"link_params_overrides": gen.generate_random_link_params(s),
                   "od_flows": gen.generate_random_od_flows(s),
                   "demand_params_overrides": gen.generate_random_demand_params(s)}
        finally:
            np.random.set_state(state)
        return out

    def _apply_scenarios(self, episode: int):
        """Builds the class table, the per-replica class indices and OD weights of this episode and hands
        them to the engine; keeps the per-replica demand parameters for `_draw_demand`.  Only the corridors a
        scenario perturbs (20 % of them) are looked at per replica: every other link keeps the class of the
        unperturbed scenario (create_network merges an override into the corridor's own block and nothing else,
        env_loader.py:93-144)."""
        from ..plan import CLASS_DTYPE, class_record
        net, gen, R = self.network, self.generator, self.R
        unit_time = net.params["unit_time"]
        if not hasattr(self, "_scn_base"):
            links = list(net.links.values())
            per_corridor = gen.scenario_link_params(None)
            table, rows = {}, []
            base_idx = np.zeros(len(links), dtype=np.int32)
            corridor_links, link_static = {}, {}
            for l in links:
                u, v = l.start_node.node_id, l.end_node.node_id
                c = (min(u, v), max(u, v))
                kw = per_corridor[c]
                static = (l.length, l._width, l.gamma, l.activity_probability, l.bi_factor, l.speed_noise_std,
                          l.fd_type, bool(l.is_separator))
                key = (static, kw["free_flow_speed"], kw["k_critical"], kw["k_jam"])
                k = table.get(key)
                if k is None:
                    k = table[key] = len(rows)
                    rows.append(class_record(static[0], static[1], key[1], key[2], key[3], *static[2:], unit_time))
                base_idx[l.index] = k
                corridor_links.setdefault(c, []).append(l.index)
                link_static[l.index] = static
            self._scn_base = (per_corridor, base_idx, corridor_links, link_static, dict(table), list(rows))
        per_corridor, base_idx, corridor_links, link_static, table0, rows0 = self._scn_base
        table, rows = dict(table0), list(rows0)
        od_keys = list(net.plan["od_keys"])
        lk_class = np.repeat(base_idx[:, None], R, axis=1)
        od_w = np.zeros((self.simulation_steps + 1, len(od_keys), R)) if od_keys else None
        self._scenarios = []
        for r in range(R):
            sc = self.scenario(r, episode)
            self._scenarios.append(sc["demand_params_overrides"])
            for link_id, over in sc["link_params_overrides"].items():
                u, v = (int(x) for x in link_id.split("_"))
                c = (min(u, v), max(u, v))
                kw = {**per_corridor[c], **over}
                for li in corridor_links[c]:
                    static = link_static[li]
                    key = (static, kw["free_flow_speed"], kw["k_critical"], kw["k_jam"])
                    k = table.get(key)
                    if k is None:
                        k = table[key] = len(rows)
                        rows.append(class_record(static[0], static[1], key[1], key[2], key[3], *static[2:], unit_time))
                    lk_class[li, r] = k
            if od_keys:
                if set(sc["od_flows"]) != set(od_keys):
                    raise NotImplementedError("randomised OD weights need the scenario's full origin x destination set")
                for j, key in enumerate(od_keys):
                    od_w[:, j, r] = sc["od_flows"][key][0]          # constant over the episode (env_loader.py:236-243)
        classes = np.array(rows, dtype=CLASS_DTYPE).reshape(len(rows))
        self.engine.set_replica_scenarios(classes, lk_class, od_w)

    # ------------------------------------------------------------------ per-replica scenarios drawn on the device
    def device_scenario_seed(self, episode: int) -> int:
        return (self.seed * 0xD1B54A32D192ED03 + 0x9E3779B97F4A7C15 * (episode + 1)) % (2 ** 64)

    def _corridors(self):
        """[(u, v)] with u < v in link-pair order: corridor c is links 2c (u -> v) and 2c + 1."""
        keys = list(self.network.links.keys())
        return [keys[2 * c] for c in range(len(keys) // 2)]

    def _apply_scenarios_device(self, episode: int):
        """randomize='device': the scenarios of all replicas drawn by one kernel launch (C-ABI pns_env_randomize; the
        reference generators' distributions with counter-based draws keyed by the global replica index), so an
        episode turnover has no host loop over replicas.  `scenario(r)` restates a replica's draws on the host."""
        net, eng, R = self.network, self.engine, self.R
        plan = net.plan
        L, S, n_od = len(net.links), self.simulation_steps, int(plan["n_od"])
        dev = eng.device
        n_base = len(plan["classes"])
        n_change = int((L // 2) * 0.2)                    # env_loader.py:401
        rows = plan["demand_nodes"]
        if not hasattr(self, "_dev_scn"):
            from ..plan import CLASS_DTYPE
            rec = CLASS_DTYPE.itemsize
            classes = torch.zeros(((n_base + R * n_change) * rec,), dtype=torch.uint8, device=dev)
            base_bytes = np.ascontiguousarray(plan["classes"]).view(np.uint8).reshape(-1)
            classes[: n_base * rec].copy_(torch.from_numpy(base_bytes.copy()))
            is_origin = np.array([1 if n.node_id in net.origin_nodes else 0 for n in rows], dtype=np.int32)
            self._dev_scn = dict(
                classes=classes, lk_class=torch.zeros((L * R,), dtype=torch.int32, device=dev),
                base_class=torch.from_numpy(np.ascontiguousarray(plan["lk_class"], dtype=np.int32)).to(dev),
                od_w=torch.zeros((S + 1, max(1, n_od) * R), dtype=torch.float64, device=dev),
                row_is_origin=torch.from_numpy(is_origin).to(dev),
                base=torch.zeros((max(1, len(rows)) * R,), dtype=torch.float64, device=dev),
                peak=torch.zeros((max(1, len(rows)) * R,), dtype=torch.float64, device=dev),
                pattern=torch.zeros((max(1, len(rows)) * R,), dtype=torch.int32, device=dev))
        d = self._dev_scn
        with eng._guard():
            _native.check(eng.lib, eng.lib.pns_env_randomize(
                C.byref(eng.net), _ptr(d["classes"]), n_base, _ptr(d["base_class"]), n_change, _ptr(d["lk_class"]),
                _ptr(d["od_w"]) if n_od else C.c_void_p(0), len(rows), _ptr(d["row_is_origin"]), _ptr(d["base"]),
                _ptr(d["peak"]), _ptr(d["pattern"]), C.c_uint64(self.device_scenario_seed(episode)),
                self.replica_base, self._stream()), "pns_env_randomize")
        eng.set_replica_scenarios_device(d["classes"], n_base + R * n_change, d["lk_class"], d["od_w"] if n_od else None)
        self._scenarios = None
        self._device_episode = episode

    def _device_scenario(self, replica: int) -> dict:
        """What the kernel drew for one replica in the current episode, read back from the device tables, in the
        shape of the reference generators' results (so `create_network(dataset, **scenario)` builds that replica)."""
        from ..plan import CLASS_DTYPE
        net, R, d = self.network, self.R, self._dev_scn
        n_base = len(net.plan["classes"])
        L, S = len(net.links), self.simulation_steps
        lk = d["lk_class"].view(L, R)[:, replica].cpu().numpy()
        table = d["classes"].cpu().numpy().view(CLASS_DTYPE)
        over = {}
        for c, (u, v) in enumerate(self._corridors()):
            k = int(lk[2 * c])
            if k >= n_base:
                rec = table[k]
                over[f"{u}_{v}"] = {"k_critical": float(rec["kc"]), "k_jam": float(rec["kj"]),
                                    "free_flow_speed": float(rec["vf"])}
        od_keys = list(net.plan["od_keys"])
        od_row = d["od_w"][0].view(-1, R)[:, replica].cpu().numpy() if od_keys else []
        od_flows = {key: np.full(S + 1, float(od_row[j])) for j, key in enumerate(od_keys)}
        names = {v: k for k, v in self._PATTERN_CODE.items()}
        rows = net.plan["demand_nodes"]
        base = d["base"].view(-1, R)[:, replica].cpu().numpy()
        peak = d["peak"].view(-1, R)[:, replica].cpu().numpy()
        pat = d["pattern"].view(-1, R)[:, replica].cpu().numpy()
        demand = {f"origin_{n.node_id}": {"pattern": names[int(pat[k])], "base_lambda": float(base[k]),
                                         "peak_lambda": float(peak[k])}
                  for k, n in enumerate(rows) if pat[k] >= 0}
        return {"link_params_overrides": over, "od_flows": od_flows, "demand_params_overrides": demand}

    # ------------------------------------------------------------------ demand
    _PATTERN_CODE = {"gaussian_peaks": 0, "constant": 1, "sudden_demand": 2}

    def demand_seed(self, episode: int) -> int:
        return (self.seed * 0x9E3779B1 + 0x7F4A7C15 * (episode + 1)) % (2 ** 64)

    def _draw_demand(self, episode: int):
        """Fills the engine's demand table [S+1, rows * R] on the device (C-ABI pns_env_draw_demand): every
        origin of every replica draws Poisson demand around its two-peak rate (od_manager.py:145-155), or
        its constant / sudden-demand variant, keyed by (demand_seed(episode), global replica; step, row) --
        no host loop over replicas, and independent of how the replicas are sharded."""
        net, S, R, eng = self.network, self.simulation_steps, self.R, self.engine
        rows = net.plan["demand_nodes"]
        if not rows:
            return
        n = len(rows)
        t = np.arange(S)
        spread = 2 * (S / 20) ** 2
        bump1 = np.exp(-(t - S / 4) ** 2 / spread)
        bump2 = np.exp(-(t - 3 * S / 4) ** 2 / spread)
        if self.randomize == "device":                    # parameters already on the device (pns_env_randomize)
            d = self._dev_scn
            if not hasattr(self, "_bumps_dev"):
                self._bumps_dev = (torch.from_numpy(bump1).to(eng.device), torch.from_numpy(bump2).to(eng.device))
            self._demand_args = (*self._bumps_dev, d["base"], d["peak"], d["pattern"])
            with eng._guard():
                _native.check(eng.lib, eng.lib.pns_env_draw_demand(
                    S, n, R, self.replica_base, C.c_uint64(self.demand_seed(episode)),
                    *[_ptr(x) for x in self._demand_args], _ptr(eng.demand), self._stream()), "pns_env_draw_demand")
            return
        base = np.zeros((n, R)); peak = np.zeros((n, R)); pattern = np.full((n, R), -1, dtype=np.int32)
        gen = net.demand_generator
        for k, node in enumerate(rows):
            if node.node_id not in net.origin_nodes:
                continue
            cfg = gen._get_demand_config(node.node_id)
            name = net.params.get("demand", {}).get(f"origin_{node.node_id}", {}).get("pattern", "gaussian_peaks")
            if name not in self._PATTERN_CODE:
                raise NotImplementedError(f"batched demand does not support the custom pattern {name!r}")
            base[k], peak[k], pattern[k] = cfg.base_lambda, cfg.peak_lambda, self._PATTERN_CODE[name]
            if self._scenarios is not None:                  # per-replica demand parameters (randomize=True)
                key = f"origin_{node.node_id}"
                for r, sc in enumerate(self._scenarios):
                    p = sc.get(key)
                    if p is not None:
                        base[k, r], peak[k, r] = p["base_lambda"], p["peak_lambda"]
                        pattern[k, r] = self._PATTERN_CODE[str(p["pattern"])]
        dev = eng.device
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self._demand_args = (to(bump1), to(bump2), to(base.reshape(-1)), to(peak.reshape(-1)), to(pattern.reshape(-1)))
        with eng._guard():
            _native.check(eng.lib, eng.lib.pns_env_draw_demand(
                S, n, R, self.replica_base, C.c_uint64(self.demand_seed(episode)),
                *[_ptr(x) for x in self._demand_args], _ptr(eng.demand), self._stream()), "pns_env_draw_demand")

    # ------------------------------------------------------------------ API
    def reset(self):
        """Start a new episode in every replica; returns obs [R, n_obs] at step 1 (all zeros + widths)."""
        eng, net = self.engine, self.network
        eng.io.seed = (self.seed + 0x9E3779B97F4A7C15 * self.episode) % (2 ** 64)
        if self.randomize == "device":
            self._apply_scenarios_device(self.episode)
        elif self.randomize:
            self._apply_scenarios(self.episode)
        eng.initialise(net._store.gate, net._store.sep_np64, self._tf, None, self._od_w, self._supplied)
        self._draw_demand(self.episode)
        self.sim_step = 1
        self.cumulative_reward.zero_()
        self.episode += 1
        self._observe(self.sim_step)
        return self.obs

    def _stream(self):
        return self.engine._stream()

    def _observe(self, row):
        eng = self.engine
        with eng._guard():
            _native.check(eng.lib, eng.lib.pns_env_observe(C.byref(eng.net), C.byref(eng.state), C.byref(self._env),
                                                           int(row), _ptr(self.obs), _ptr(self.reward),
                                                           self._stream()), "pns_env_observe")

    def step(self, actions: torch.Tensor, obs_out: torch.Tensor = None, reward_out: torch.Tensor = None):
        """actions [R, n_act] float32 on the env's device.  Returns (obs, reward, done, info).
        obs_out / reward_out: device tensors to receive this step's observations / rewards instead of the
        env's own buffers (pipelined host loops keep two sets, see `rollout_host`)."""
        eng = self.engine
        obs = self.obs if obs_out is None else obs_out
        reward = self.reward if reward_out is None else reward_out
        if obs.shape != self.obs.shape or obs.dtype != torch.float32 or reward.shape != self.reward.shape:
            raise ValueError("obs_out / reward_out must match the env's obs / reward tensors")
        if self.sim_step > self.simulation_steps:
            raise RuntimeError("episode finished: call reset()")
        if actions is not None and self.n_act:
            if actions.shape != (self.R, self.n_act) or actions.dtype != torch.float32:
                raise ValueError(f"actions must be float32 [{self.R}, {self.n_act}]")
            actions = actions.contiguous()
        if self.action_gap > 1:
            return self._step_with_gap(actions, obs, reward)
        with eng._guard():      # actions, the LTM step, observations + reward: one native call behind one custom op
            eng._begin_steps(self.sim_step, 1)
            has_actions = actions is not None and self.n_act > 0
            ops.env_step(eng.hist64, eng.hist32, eng.runsum, eng.tf_routed, eng.probs, eng.err, eng.gate,
                         actions if has_actions else self._no_actions, obs, reward, self.cumulative_reward,
                         eng.handle, int(self.sim_step), has_actions)
        eng.t_done = self.sim_step
        done = self.sim_step >= self.simulation_steps           # tested before the increment (quirk Q8)
        self.sim_step += 1
        return obs, reward, done, {"step": self.sim_step - 1}

    def _step_with_gap(self, actions, obs, reward):
        """action_gap > 1 (rl/pz_pednet_env.py:224-253): the actions are applied once, then `action_gap` simulation
        steps run; the observation is the last step's, the reward the float32 sum of the steps' rewards in step
        order, and the running total grows by that sum once.  Sub-steps are ordinary native environment steps
        without actions, their rewards are added up by torch."""
        eng = self.engine
        if self.sim_step + self.action_gap - 1 > self.simulation_steps:
            raise RuntimeError("the decision's simulation steps run past the end of the episode")
        if not hasattr(self, "_gap_reward"):
            self._gap_reward = torch.zeros_like(self.reward)
            self._gap_total = torch.zeros_like(self.reward)     # receives the native running total (not used)
        done = False
        with eng._guard():
            eng._begin_steps(self.sim_step, self.action_gap)
            for k in range(self.action_gap):
                has_actions = k == 0 and actions is not None and self.n_act > 0
                ops.env_step(eng.hist64, eng.hist32, eng.runsum, eng.tf_routed, eng.probs, eng.err, eng.gate,
                             actions if has_actions else self._no_actions, obs, reward if k == 0 else self._gap_reward,
                             self._gap_total, eng.handle, int(self.sim_step), has_actions)
                if k:
                    reward.add_(self._gap_reward)
                eng.t_done = self.sim_step
                done = self.sim_step >= self.simulation_steps
                self.sim_step += 1
            self.cumulative_reward.add_(reward)
        return obs, reward, done, {"step": self.sim_step - 1}

    def rollout(self, actions: torch.Tensor, obs_out: torch.Tensor = None, reward_out: torch.Tensor = None):
        """K environment steps with given actions ([K, R, n_act] float32 on the device) in one native call
        (`torch.ops.pednstream.env_rollout` -> pns_env_rollout): returns (obs [K, R, n_obs], reward [K, R], done)."""
        eng = self.engine
        K = int(actions.shape[0])
        if self.action_gap > 1:                # decisions with several simulation steps: a loop over `step`
            obs = obs_out if obs_out is not None else torch.empty((K, self.R, self.obs.shape[1]), dtype=torch.float32, device=self.device)
            rew = reward_out if reward_out is not None else torch.empty((K, self.R), dtype=torch.float32, device=self.device)
            done = False
            for k in range(K):
                _, _, done, _ = self.step(actions[k], obs[k], rew[k])
            return obs, rew, done
        if self.sim_step + K - 1 > self.simulation_steps:
            raise RuntimeError("rollout runs past the end of the episode")
        if tuple(actions.shape[1:]) != (self.R, self.n_act) or actions.dtype != torch.float32:
            raise ValueError(f"actions must be float32 [K, {self.R}, {self.n_act}]")
        actions = actions.contiguous()
        obs = obs_out if obs_out is not None else torch.empty((K, self.R, self.obs.shape[1]), dtype=torch.float32, device=self.device)
        rew = reward_out if reward_out is not None else torch.empty((K, self.R), dtype=torch.float32, device=self.device)
        with eng._guard():
            eng._begin_steps(self.sim_step, K)
            has_actions = self.n_act > 0
            ops.env_rollout(eng.hist64, eng.hist32, eng.runsum, eng.tf_routed, eng.probs, eng.err, eng.gate,
                            actions if has_actions else self._no_actions, obs, rew, self.cumulative_reward,
                            eng.handle, int(self.sim_step), K, has_actions)
        self.sim_step += K
        eng.t_done = self.sim_step - 1
        self.obs.copy_(obs[K - 1]); self.reward.copy_(rew[K - 1])
        return obs, rew, self.sim_step > self.simulation_steps

    def rollout_host(self, host_actions: torch.Tensor, host_obs: torch.Tensor, host_reward: torch.Tensor):
        """K environment steps driven from host memory in one native call (pns_env_rollout, host form): step k takes
        `host_actions[k]` ([K, R, n_act] pinned float32) and delivers its observations and rewards to `host_obs[k]`
        ([K, R, n_obs]) / `host_reward[k]` ([K, R]), both pinned.  Every step has its own host->device and
        device->host copies; they run on two copy streams, double-buffered against the step kernels (actions of step
        k+1 go up and results of step k come down while the other step computes).  Returns after everything is
        enqueued; synchronise the device (or the current stream) before reading the host tensors."""
        K = int(host_actions.shape[0])
        if self.action_gap > 1:
            raise NotImplementedError("rollout_host runs decisions of one simulation step (action_gap = 1); "
                                      "use step() / rollout() with action_gap > 1")
        if not (host_actions.is_pinned() and host_obs.is_pinned() and host_reward.is_pinned()):
            raise ValueError("rollout_host needs pinned host tensors")
        if host_obs.shape[0] < K or host_reward.shape[0] < K:
            raise ValueError("host_obs / host_reward are shorter than host_actions")
        if self.sim_step + K - 1 > self.simulation_steps:
            raise RuntimeError("rollout runs past the end of the episode")
        if (host_actions.dtype != torch.float32 or tuple(host_actions.shape[1:]) != (self.R, self.n_act)
                or not host_actions.is_contiguous() or not host_obs.is_contiguous() or not host_reward.is_contiguous()
                or tuple(host_obs.shape[1:]) != tuple(self.obs.shape) or tuple(host_reward.shape[1:]) != (self.R,)):
            raise ValueError("host tensors must be contiguous float32 [K, R, n_act] / [K, R, n_obs] / [K, R]")
        dev, eng = self.device, self.engine
        if not hasattr(self, "_stage"):
            self._stage = (torch.zeros((2, self.R, max(1, self.n_act)), dtype=torch.float32, device=dev),
                           torch.zeros((2,) + tuple(self.obs.shape), dtype=torch.float32, device=dev),
                           torch.zeros((2, self.R), dtype=torch.float32, device=dev))
        act, obs, rew = self._stage
        with eng._guard():
            eng._begin_steps(self.sim_step, K)
            eng._native_env_rollout(act if self.n_act else None, obs, rew, self.cumulative_reward, self.sim_step, K,
                                    host=(host_actions if self.n_act else None, host_obs, host_reward))
        self.sim_step += K
        eng.t_done = self.sim_step - 1

    def kpis(self, t_last: int = None) -> torch.Tensor:
        """Per-replica episode KPIs from the device history up to row t_last (default: the last simulated
        step): tensor [R, len(_native.KPI_NAMES)]; `pednstream_b200.kpi.kpi_dict` turns rows into the
        reference's result dictionaries (rl/rl_utils.py:770-1512), `parallel.gather_replica_values`
        collects them across GPUs."""
        t_last = self.sim_step - 1 if t_last is None else int(t_last)
        role, any_path = self.network.link_roles()
        return self.engine.kpis(t_last, role, any_path)

    def split_obs(self, obs=None):
        obs = self.obs if obs is None else obs
        return {a: obs[:, s] for a, s in self.obs_slices.items()}

    def launches_per_step(self):
        """Kernel launches of one environment step."""
        routed = 1 if len(self.network.plan["rt_grp_node"]) else 0
        if self.R > 1 and not self.engine.emulation:
            # flows (+actions, + route choice as extra CTAs) | node | update (+observations, reward); only the first
            # step of an episode with actions runs the route choice as its own launch
            return 3
        return 1 + 3 + routed + 1      # actions, flows | [route] | node | update, observe+reward
